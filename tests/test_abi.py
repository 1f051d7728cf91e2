"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes

import numpy as np
import pytest

from debwt_b200 import api, binding


def test_library_exports_every_declared_symbol():
    L = binding.lib()
    names = binding.declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), n
    assert set(binding._SIGS) == set(names)


def test_library_exports_every_device_stage_symbol():
    L = binding.lib()
    names = binding.declared_symbols(binding.DEV_HEADER)
    assert len(names) >= 26
    for n in names:
        assert hasattr(L, n), n


def test_stats_struct_matches_header_layout():
    # 7 u64 + 12 float + 3 u32 = 56 + 48 + 12 = 116 -> padded to 120
    assert ctypes.sizeof(binding.Stats) == 120


def test_no_cpu_fallback_without_device():
    if binding.lib().debwt_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(binding.DebwtError) as e:
        api.build_bwt(["ACGT" * 20])
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(binding.DebwtError):
        api.k_radix_sort(np.arange(10, dtype=np.uint64))


def test_join_records_layout():
    text, seps = api.join_records(["ACGT" * 10, b"TTTT" * 9])
    assert bytes(text[40:41]) == b"#" and bytes(text[-1:]) == b"$"
    assert list(seps) == [40, 77] and text.size == 78
