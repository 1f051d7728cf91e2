"""Developer probe: every sort configuration on device-generated random keys -- correctness against numpy on a small
array, then ms per sort and per digit pass (16 B/key algorithmic per pass) against the measured copy bandwidth."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from debwt_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000_000
cfgs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8] + list(range(32, 48))
peak = 6549.4
try:
    peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:  # noqa: BLE001
    pass
rng = np.random.default_rng(1)
small = rng.integers(0, 2**64, size=3_000_001, dtype=np.uint64)
want = np.sort(small)
rows = []
for cfg in cfgs:
    try:
        got, _ = api.k_radix_sort(small, cfg=cfg)
        ok = bool((got == want).all())
        ms, msp = api.bench_sort_passes(n, cfg=cfg, iters=3)
        row = {"cfg": cfg, "ok": ok, "ms_sort": ms, "ms_pass": msp, "pass_GBps": 16 * n / msp / 1e6, "pass_frac": 16 * n / msp / 1e6 / peak}
    except Exception as e:  # noqa: BLE001
        row = {"cfg": cfg, "error": str(e)}
    rows.append(row)
    print(json.dumps(row), flush=True)
json.dump({"n": n, "peak": peak, "rows": rows}, open("gpurun_out/sort_sweep.json", "w"), indent=1)
