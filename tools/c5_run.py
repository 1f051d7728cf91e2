"""torchrun --nproc-per-node N tools/c5_run.py [--hap-len L] [--n-hap H] [--rec-per-hap Q] [--builds B] [--walk STEPS]

BASELINE.json configs[4] (C5): a collection of H haplotypes of a C3-like genome of L bases, each cut into Q records, haplotypes
2..H at 0.1 % substitution divergence from the first -- by default 10 x 3.0 Gbp x 24 records = 30 Gbp, 240 records, on 8 GPUs.
Every rank generates its own position slice of T in HBM (the base genome by debwt_b200/synth_gpu.py, seed 20; haplotype h by
synth mutate with seed 30 + h), the sharded build runs B times, and rank 0 checks the result without any CPU oracle (the
reference cannot run at this size, README.md:18): SHA-256 of the packed BWT, base counts of the BWT against the text's, number
of '#' rows, and a sequential LF walk from the '$' row that must spell the last STEPS symbols of T backwards.
With small L the same script runs on 1 GPU as well, which gives the SHA to compare against."""
import argparse
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch
import torch.distributed as dist

from debwt_b200 import api, dist as D, synth_gpu

ap = argparse.ArgumentParser()
ap.add_argument("--hap-len", type=int, default=3_000_000_000)
ap.add_argument("--n-hap", type=int, default=10)
ap.add_argument("--rec-per-hap", type=int, default=24)
ap.add_argument("--builds", type=int, default=3)
ap.add_argument("--walk", type=int, default=5_000_000)
args = ap.parse_args()

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm, ops = D.Comm(), D.CudaOps(local)
ops.timed_main_sort = True
L, H, Q = args.hap_len, args.n_hap, args.rec_per_hap
per = -(-L // Q)
rec_lens = [min(per, L - i) for i in range(0, L, per)]
hap_span = L + len(rec_lens)                      # bases + one separator per record
N = H * hap_span
seps = []
for h in range(H):
    pos = h * hap_span
    for ln in rec_lens:
        pos += ln
        seps.append(pos)
        pos += 1
seps = np.array(seps, dtype=np.uint64)
lo, hi = D.my_slice(N, comm)

t0 = time.time()
base = synth_gpu.genome_like(L, 20, device=local)


def hap_text(h):
    """haplotype h with its separators, as it appears in T ('#' after every record, '$' at the very end of T)"""
    g = base if h == 0 else synth_gpu.mutate(base, 30 + h, 0.001)
    out = torch.empty(hap_span, dtype=torch.uint8, device=g.device)
    src = dst = 0
    for ln in rec_lens:
        out[dst:dst + ln] = g[src:src + ln]
        out[dst + ln] = ord("#")
        src += ln
        dst += ln + 1
    if h == H - 1:
        out[-1] = ord("$")
    return out


d_slice = torch.empty(max(hi - lo, 1), dtype=torch.uint8, device=base.device)
counts = torch.zeros(256, dtype=torch.int64, device=base.device)
for h in range(H):
    a, b = h * hap_span, (h + 1) * hap_span
    if b <= lo or a >= hi:
        continue
    ht = hap_text(h)
    s, e = max(lo, a), min(hi, b)
    d_slice[s - lo:e - lo] = ht[s - a:e - a]
    del ht
if hi > lo:
    for c0 in range(0, hi - lo, 1 << 28):
        counts += torch.bincount(d_slice[c0:min(c0 + (1 << 28), hi - lo)].to(torch.int64), minlength=256)
tail = hap_text(H - 1)[-args.walk:].clone() if rank == 0 else None
del base
torch.cuda.empty_cache()
torch.cuda.synchronize()
t_gen = time.time() - t0
if world > 1:
    dist.all_reduce(counts)
text_counts = [int(counts[ord(c)]) for c in "ACGT"]

times, prof, out, stats = [], None, None, {}
for it in range(args.builds + 1):
    stats = {"profile": it == args.builds}          # last iteration: synchronised per-phase wall times (not a timing run)
    comm.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    out = D.build_sharded(None, seps, comm, ops, stats, n_symbols=N, ascii_slice=d_slice, fetch=False)
    torch.cuda.synchronize(); comm.barrier()
    times.append((time.perf_counter() - t0) * 1e3)
    if it == args.builds:
        prof = stats.get("phases_ms")
mem = torch.cuda.max_memory_allocated() / 2**30
if rank == 0:
    bwt, sharp_all, dollar = out
    h = hashlib.sha256()
    for c0 in range(0, bwt.numel(), 1 << 26):
        h.update(bwt[c0:c0 + (1 << 26)].cpu().numpy().tobytes())
    sharp_np = np.sort(sharp_all.cpu().numpy().view(np.uint64))
    dollar_row = int(dollar.cpu().numpy().view(np.uint64)[0])
    del d_slice
    torch.cuda.empty_cache()                        # the verifier's occ table (N bytes) comes from cudaMalloc
    bad, carr = api.verify_walk_device(bwt.data_ptr(), N, sharp_np, dollar_row, tail.data_ptr(), int(tail.numel()), device=local)
    bwt_counts = [int(carr[i + 1] - carr[i]) for i in range(4)]
    res = {"workload": f"{H} haplotypes x {L} bases x {len(rec_lens)} records at 0.1 % divergence", "n_gpus": world, "n_symbols": N,
           "n_bases": H * L, "n_records": int(seps.size), "gen_s": t_gen, "ms_per_build": times[:-1], "ms_best": min(times[:-1]),
           "Mbp_s": H * L / min(times[:-1]) / 1e3, "bwt_sha256": h.hexdigest(),
           "sharp_sha256": hashlib.sha256(sharp_np.tobytes()).hexdigest(), "n_sharp_rows": int(sharp_np.size), "dollar_row": dollar_row,
           "verify": {"lf_walk_steps": int(tail.numel()), "lf_walk_mismatches": bad, "bwt_base_counts": bwt_counts,
                      "text_base_counts": text_counts, "base_counts_equal": bwt_counts == text_counts,
                      "sharp_rows_ok": int(sharp_np.size) == int(seps.size) - 1},
           "keys_local": stats.get("keys_local"), "n_branch": stats.get("n_branch"), "n_blue": stats.get("n_blue"),
           "n_codes": stats.get("n_codes"), "sort_rank0": ops.sort_stats, "rank0_phases_ms_synchronised": prof,
           "torch_peak_gib_rank0": mem}
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
