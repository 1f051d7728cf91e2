"""torchrun --nproc-per-node N tools/dist_profile_nccl.py [workload] : per-phase wall times of the sharded path"""
import json, os, sys
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from debwt_b200 import api, dist as D, synth
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
recs = {"c2": synth.config2, "c4s": lambda: synth.config4(10_000_000, 10), "c3s": lambda: synth.config3(400_000_000, 4)}[name]()
text, seps = api.join_records(recs)
comm, ops = D.Comm(), D.CudaOps(local)
lo, hi = D.my_slice(text.size, comm)
d_slice = torch.from_numpy(text[lo:hi].copy()).cuda()
import time
for it in range(8):
    stats = {"profile": it >= 4}
    comm.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    D.build_sharded(None, seps, comm, ops, stats, n_symbols=text.size, ascii_slice=d_slice, fetch=False)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    if rank == 0:
        print(it, "wall %.2f ms" % dt, json.dumps({k: round(v, 2) for k, v in stats.get("phases_ms", {}).items()}), flush=True)
dist.destroy_process_group()
