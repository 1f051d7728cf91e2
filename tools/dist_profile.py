import sys, json
sys.path.insert(0, ".")
import numpy as np, torch
from debwt_b200 import api, dist as D, synth
recs = synth.config2()
text, seps = api.join_records(recs)
comm, ops = D.Comm(), D.CudaOps(0)
d_slice = torch.from_numpy(text).cuda()
for it in range(3):
    stats = {"profile": True}
    D.build_sharded(None, seps, comm, ops, stats, n_symbols=text.size, ascii_slice=d_slice, fetch=False)
    torch.cuda.synchronize()
print(json.dumps(stats["phases_ms"], indent=1))
