"""One build of a config (for ncu launch lists): python tools/one_build.py c2 [reps]"""
import sys

sys.path.insert(0, ".")
from debwt_b200 import api, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
recs = {"c1": synth.config1, "c2": synth.config2, "c2_20M": lambda: synth.config2(20_000_000),
        "c4s": lambda: synth.config4(10_000_000, 10), "c3s": lambda: synth.config3(200_000_000, 4),
        "c4m": lambda: synth.config4(30_000_000, 10), "c3m": lambda: synth.config3(400_000_000, 4)}[name]()
with api.BwtBuilder() as b:
    for _ in range(reps):
        b.set_records(recs)
        b.build()
        b.result()
    print(b.stats())
