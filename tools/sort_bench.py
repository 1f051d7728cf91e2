import sys
sys.path.insert(0, ".")
from debwt_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
cfgs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else list(range(9))
for cfg in cfgs:
    ms = api.bench_sort(n, cfg=cfg, iters=5)
    print(f"cfg {cfg}: {ms:.3f} ms  {136*n/ms/1e6:.0f} GB/s(136B/key)  frac_of_6549={136*n/ms/1e6/6549.4:.3f}", flush=True)
