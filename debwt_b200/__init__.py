"""deBWT-B200: B200-native BWT construction (de Bruijn branch method) behind a C ABI.

`debwt_b200.api` mirrors the reference's stage interface; the compute lives in
libdebwt_b200.so (debwt_b200/csrc, hand-written sm_100a CUDA).  No CPU fallback.
"""
from .binding import DebwtError  # noqa: F401
