"""torchrun --nproc-per-node N tools/c3_sharded.py <n_bases> <n_records> : sharded build of a C3-like genome,
prints timing and the SHA-256 of the packed BWT (compare with tools/c3_validate.py on one GPU)."""
import hashlib, json, os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from debwt_b200 import api, dist as D, synth

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, nrec = int(float(sys.argv[1])), int(sys.argv[2])
t0 = time.time(); recs = synth.config4(n // nrec, nrec) if os.environ.get('DEBWT_CFG') == 'c4' else synth.config3(n, nrec); tg = time.time() - t0
text, seps = api.join_records(recs)
del recs
comm, ops = D.Comm(), D.CudaOps(local)
ops.timed_main_sort = True
lo, hi = D.my_slice(text.size, comm)
d_slice = torch.from_numpy(text[lo:hi].copy()).cuda()
N = int(text.size)
del text
times = []
for it in range(4):
    stats = {"profile": it == 3}        # last iteration: synchronised per-phase wall times (not a timing run)
    comm.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    out = D.build_sharded(None, seps, comm, ops, stats, n_symbols=N, ascii_slice=d_slice, fetch=(it == 2))
    if it == 2: fetched = out
    if it == 3: prof = stats.get("phases_ms")
    torch.cuda.synchronize(); times.append((time.perf_counter() - t0) * 1e3)
if rank == 0:
    w, s, d = fetched
    res = {"n_gpus": comm.size, "n_bases": n, "gen_s": tg, "ms_per_build": times, "Mbp_s": n / min(times[:2]) / 1e3,
           "sha256": hashlib.sha256(w.tobytes()).hexdigest(), "sharp_sha256": hashlib.sha256(s.tobytes()).hexdigest(),
           "dollar": int(d[0]), "keys_local": stats["keys_local"], "n_branch": stats["n_branch"], "n_blue": stats["n_blue"],
           "n_codes": stats["n_codes"], "sort": ops.sort_stats, "rank0_phases_ms": prof}
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
