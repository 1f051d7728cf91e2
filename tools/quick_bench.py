"""Developer timing probe (not the bench contract): sort configurations + per-phase build times."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from debwt_b200 import api, synth  # noqa: E402

out = {}
n_sort = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
for cfg in (0, 1, 2, 3, 4, 5):
    try:
        ms = api.bench_sort(n_sort, cfg=cfg, iters=3)
        out[f"sort_cfg{cfg}"] = {"ms": ms, "GBps_136B": 136 * n_sort / ms / 1e6, "Gkeys_s": n_sort / ms / 1e6}
    except Exception as e:  # noqa: BLE001
        out[f"sort_cfg{cfg}"] = {"error": str(e)}
    print(cfg, out[f"sort_cfg{cfg}"], flush=True)

for name, gen in (("c1", lambda: synth.config1()), ("c2_20M", lambda: synth.config2(20_000_000)),
                  ("c2", lambda: synth.config2()), ("c4_10x10M", lambda: synth.config4(10_000_000, 10)),
                  ("c3_200M", lambda: synth.config3(200_000_000, 4))):
    t0 = time.time()
    recs = gen()
    tg = time.time() - t0
    with api.BwtBuilder() as b:
        for rep in range(2):
            t0 = time.time()
            b.set_records(recs)
            b.build()
            res = b.result()
            wall = time.time() - t0
        st = b.stats()
    st["wall_s_e2e"] = wall
    st["gen_s"] = tg
    st["Mbp_s_e2e"] = sum(r.size for r in recs) / wall / 1e6
    out[name] = st
    print(name, json.dumps(st), flush=True)
json.dump(out, open("gpurun_out/quick_bench.json", "w"), indent=1)
