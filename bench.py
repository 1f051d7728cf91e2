#!/usr/bin/env python
"""bench.py -- BWT build throughput (Mbp/s) of the deBWT hot path on B200, next to the CPU reference.

Contract (one JSON line on stdout, printed by rank 0):
  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]
  N > 1 is launched by the driver through torch.distributed.run (one rank per GPU).

A "step" is one pass of the whole hot path (2-bit pack -> 32-mer extraction -> LSD radix sort ->
branch k-mer detection -> branch codes -> segmented sort -> BWT emission) over one synthetic genome.
  value : input bases / device time, text already resident in HBM (CUDA events on the library's stream)
  e2e   : same metric through the public C-ABI call sequence with HOST buffers: pinned host text ->
          H2D -> build -> D2H of the packed BWT, every step
  roofline : the dominant kernel (onesweep_kernel, the radix-sort scatter pass): 16 B/key algorithmic
          per launch / mean launch time, against the measured HBM copy bandwidth
  cpu_baseline : the compiled reference (oracle/_ref/deBWT + Jellyfish stand-in) on a bounded sample
--impl reference times the reference's own CPU implementation on the same workload (bounded sample).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bwt_build_throughput"
UNIT = "Mbp/s"


# ------------------------------------------------------------------------------------------------
# workloads (BASELINE.json configs)
# ------------------------------------------------------------------------------------------------
def make_workload(name: str, rank: int = 0):
    from debwt_b200 import synth
    if name == "c1":
        return synth.config1(), "synthetic 4.6 Mbp random ACGT single sequence, k=32"
    if name == "c2":
        return synth.config2(), "synthetic 100 Mbp sequence with planted repeats (5% copies of 10 kbp elements), k=32"
    if name == "c2s":
        return synth.config2(20_000_000), "synthetic 20 Mbp sequence with planted repeats (reduced C2, smoke only)"
    if name == "c4s":
        return synth.config4(10_000_000, 10), "10 synthetic 10 Mbp genomes at 0.1% divergence (reduced C4)"
    if name == "c3s":
        return synth.config3(400_000_000, 4), "synthetic 400 Mbp genome with repeat families (reduced C3)"
    raise SystemExit(f"unknown workload {name}")


def sample_records(records, max_bases: int):
    """Bounded prefix of the workload for the CPU arms."""
    out, left = [], max_bases
    for r in records:
        if left <= 33:
            break
        out.append(r[:left])
        left -= out[-1].size
    return out


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region through NVML (in-process: spawning
    nvidia-smi in a loop perturbs the driver enough to slow the measured kernels down)."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period, self.rows, self.stop, self.thr, self.h = index, period_s, [], threading.Event(), None, None
        self.nv = None

    def __enter__(self):
        if os.environ.get("DEBWT_NO_CLOCKS"):
            return self
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)   # slow call: once, up front
            self.thr = threading.Thread(target=self._run, daemon=True)
            self.thr.start()
        except Exception:  # noqa: BLE001
            self.nv = None
        return self

    def _run(self):
        nv = self.nv
        while not self.stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = self.max_mhz
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, mx, rs))
            except Exception:  # noqa: BLE001
                pass
            self.stop.wait(self.period)

    def __exit__(self, *a):
        self.stop.set()
        if self.thr:
            self.thr.join(timeout=2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(k for k, bit in names.items() if any(r[2] & bit for r in self.rows))
        return {"sm_mhz": statistics.median(r[0] for r in self.rows), "sm_max_mhz": max(r[1] for r in self.rows),
                "reasons": reasons, "samples": len(self.rows)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except (KeyError, ValueError):
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_launch():
    """dram bytes per onesweep launch from the committed ncu summary (profiles/), scaled per key."""
    p = os.path.join(ROOT, "profiles", "onesweep_traffic.json")
    if os.path.isfile(p):
        try:
            return json.load(open(p))
        except ValueError:
            return None
    return None


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ------------------------------------------------------------------------------------------------
def run_reference_sample(records, max_bases: int, threads: int):
    from debwt_b200 import synth
    from oracle import refrun
    if not refrun.available():
        raise RuntimeError("oracle/_ref is not built")
    sample = sample_records(records, max_bases)
    nb = int(sum(r.size for r in sample))
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "sample.fa")
        synth.write_fasta(sample, fa)
        res = refrun.run_reference(fa, threads=threads, timeout=3600, scratch_root=d)
    return nb, res


def reference_arm(args, rank, world):
    """The reference's own CPU implementation (oracle/_ref/deBWT, all host threads) on a bounded sample
    of the same workload, sized so that steps + warmup runs end within a few minutes."""
    if rank != 0:
        return
    records, desc = make_workload(args.workload)
    cores = os.cpu_count() or 1
    total = int(sum(r.size for r in records))
    runs = max(args.steps + args.warmup, 1)
    per_run_s = 170.0 / runs                                  # whole arm within ~3 minutes
    sample = int(min(total, max(2_000_000, (per_run_s - 3.0) * 1.0e6)))   # ~1 Mbp/s + ~3 s fixed cost per run
    if args.ref_sample:
        sample = min(total, args.ref_sample)
    times, standin = [], []
    nb = sample
    for i in range(runs):
        nb, res = run_reference_sample(records, sample, cores)
        if i >= args.warmup:
            times.append(res.wall_s); standin.append(res.standin_s)
    t = sum(times) / len(times)
    v = nb / t / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": desc, "sample": f"first {nb} bases of the workload per step"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": f"oracle/_ref/deBWT -t {cores} -k 32 on the first {nb} bases; Jellyfish replaced by "
                                       f"oracle/jellyfish_standin.c ({sum(standin) / len(standin):.2f} s of the {t:.2f} s per step)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    from debwt_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    records, desc = make_workload(args.workload, rank)
    text_np, seps = api.join_records(records)
    n = int(text_np.size)
    n_bases = n - len(records)
    # pinned host staging + device-resident copy
    h_text = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_text.numpy()[:] = text_np
    d_text = h_text.to("cuda", non_blocking=False)
    n_words = (n + 31) // 32
    h_out = torch.empty(n_words, dtype=torch.int64, pin_memory=True)

    b = api.BwtBuilder(device=local_rank)

    def step_resident():
        b.set_text_device(d_text.data_ptr(), n, seps)
        b.build()
        return b.stats()

    def step_e2e():
        b.set_text_ptr(h_text.data_ptr(), n, seps)
        b.build()
        b.result_into(h_out.data_ptr())
        return b.stats()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        dev_ms, sort_ms, sweep_ms, launches, sweeps, last = 0.0, 0.0, 0.0, 0, 0, None
        for _ in range(args.steps):
            st = step_resident()
            dev_ms += st["ms_total"]; sort_ms += st["ms_sort"]; sweep_ms += st["ms_sort_sweeps"]
            launches += st["total_launches"]; sweeps += st["sort_sweeps"]
            last = st
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = clk.summary()
    ms_dev = dev_ms / args.steps

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps

    if dist is not None:
        tt = torch.tensor([ms_dev, e2e_ms, wall_ms / args.steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_dev, e2e_ms, wall_step = (float(x) for x in tt.tolist())
    else:
        wall_step = wall_ms / args.steps

    if rank == 0:
        peak, peak_src = measured_peak()
        nk = last["n_keys"]
        per_launch_ms = sweep_ms / max(sweeps, 1)
        achieved = 16.0 * nk / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
        phase = 136.0 * nk / ((sort_ms / args.steps) * 1e-3) / 1e9 if sort_ms > 0 else 0.0
        traffic = ncu_traffic_per_launch()
        line = {
            "metric": METRIC, "value": world * n_bases / (ms_dev * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": desc, "n_bases": n_bases, "n_records": len(records), "k": 32,
                       "l2": "inputs larger than L2 (text %d MB, keys %d MB per step)" % (n // 2**20, 8 * nk // 2**20),
                       "parallelism": "replicas" if world > 1 else "single GPU",
                       "timing": "CUDA events on the library stream (debwt_stats.ms_total); wall per step %.3f ms" % wall_step},
            "clocks": clocks,
            "e2e": {"value": world * n_bases / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": n,
                    "d2h_bytes_per_step": 8 * n_words + 8 * len(records), "ms_per_step": e2e_ms},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "onesweep_kernel (radix-sort scatter pass, %d launches/step)" % (sweeps // args.steps),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8000": achieved / 8000.0,     # north_star quotes ~8 TB/s per GPU
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": 16 * nk,
                         "mean_launch_ms": per_launch_ms,
                         "traffic": (traffic or {}).get("dram_bytes_per_key", None) and traffic["dram_bytes_per_key"] * nk,
                         "sort_phase": {"achieved": phase, "frac": phase / peak, "bytes_per_key": 136,
                                        "ms": sort_ms / args.steps}},
            "phases_ms": {k: last[k] for k in last if k.startswith("ms_")},
            "sizes": {k: last[k] for k in ("n_symbols", "n_keys", "n_branch", "n_blue", "n_codes", "n_special")},
            # checksum of the packed BWT the last end-to-end step copied back: the same at every N (outside the timed regions)
            "bwt_sha256": __import__("hashlib").sha256(h_out.numpy().tobytes()).hexdigest(),
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                cores = os.cpu_count() or 1
                nb, res = run_reference_sample(records, args.cpu_sample, cores)
                line["cpu_baseline"] = {"value": nb / res.wall_s / 1e6, "unit": UNIT, "cores": cores, "kind": "reference",
                                        "sample": f"oracle/_ref/deBWT -t {cores} -k 32 on the first {nb} bases of the workload: "
                                                  f"{res.wall_s:.2f} s wall of which {res.standin_s:.2f} s in the Jellyfish stand-in "
                                                  f"(deBWT proper {nb / max(res.proper_s, 1e-9) / 1e6:.2f} Mbp/s)"}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                                        "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    b.close()
    if dist is not None:
        dist.destroy_process_group()


def ours_sharded(args, rank, world, local_rank):
    """N GPUs, one process per GPU: the text is split by position, keys are range-partitioned by sampled
    splitters and exchanged with one NCCL all-to-all (debwt_b200/dist.py).  Strong scaling: the same
    genome on N GPUs."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from debwt_b200 import api, binding, dist as D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    comm = D.Comm()
    ops = D.CudaOps(local_rank)
    ops.timed_main_sort = True
    records, desc = make_workload(args.workload, rank)
    text_np, seps = api.join_records(records)
    n = int(text_np.size)
    n_bases = n - len(records)
    lo, hi = D.my_slice(n, comm)
    h_slice = torch.empty(max(hi - lo, 1), dtype=torch.uint8, pin_memory=True)
    h_slice.numpy()[:hi - lo] = text_np[lo:hi]
    d_slice = h_slice.to("cuda")
    n_words = (n + 31) // 32
    h_out = torch.empty(n_words, dtype=torch.int64, pin_memory=True) if rank == 0 else None
    del text_np

    def barrier():
        comm.barrier()
        torch.cuda.synchronize()

    stats = {}

    def step(resident: bool):
        src = d_slice if resident else h_slice.to("cuda", non_blocking=True)
        out = D.build_sharded(None, seps, comm, ops, stats, n_symbols=n, ascii_slice=src, fetch=False)
        if not resident and rank == 0:
            h_out.copy_(out[0], non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(True)
    barrier()
    l0 = binding.lib().debwt_launch_count()
    sort_ms = sweep_ms = 0.0
    sweeps = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(args.steps):
            step(True)
            sort_ms += ops.sort_stats["ms"]; sweep_ms += ops.sort_stats["ms_sweeps"]; sweeps += ops.sort_stats["sweeps"]
        ev1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    dev_ms = ev0.elapsed_time(ev1) / args.steps
    launches = int(binding.lib().debwt_launch_count() - l0)
    clocks = clk.summary()
    for _ in range(2):
        step(False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(False)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    tt = torch.tensor([dev_ms, e2e_ms, wall_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, wall_ms = (float(x) for x in tt.tolist())
    if rank == 0:
        peak, peak_src = measured_peak()
        nk_loc = ops.sort_stats["n"]
        per_launch_ms = sweep_ms / max(sweeps, 1)
        achieved = 16.0 * nk_loc / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
        phase = 136.0 * nk_loc / ((sort_ms / args.steps) * 1e-3) / 1e9 if sort_ms > 0 else 0.0
        traffic = ncu_traffic_per_launch()
        line = {
            "metric": METRIC, "value": n_bases / (dev_ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": desc, "n_bases": n_bases, "n_records": len(records), "k": 32,
                       "l2": "inputs larger than L2 (keys %d MB per GPU per step)" % (8 * nk_loc // 2**20),
                       "parallelism": "position split x%d, keys range-partitioned by sampled splitters, NCCL all-to-all" % world,
                       "timing": "CUDA events on the torch stream around the step, max over ranks; wall per step %.3f ms" % wall_ms,
                       "keys_per_gpu": stats.get("keys_local"), "nccl_bytes_sent_rank0_total": stats.get("bytes_sent")},
            "clocks": clocks,
            "e2e": {"value": n_bases / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": n, "d2h_bytes_per_step": 8 * n_words,
                    "ms_per_step": e2e_ms},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "onesweep_kernel (radix-sort scatter pass on rank 0's key range, %d launches/step)"
                                                   % (sweeps // max(args.steps, 1)),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "frac_of_nominal_8000": achieved / 8000.0,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": 16 * nk_loc, "mean_launch_ms": per_launch_ms,
                         "traffic": (traffic or {}).get("dram_bytes_per_key", None) and traffic["dram_bytes_per_key"] * nk_loc,
                         "sort_phase": {"achieved": phase, "frac": phase / peak, "bytes_per_key": 136, "ms": sort_ms / args.steps}},
            "sizes": {k: stats.get(k) for k in ("n_symbols", "n_keys", "n_branch", "n_blue", "n_codes")},
            # checksum of the packed BWT the last end-to-end step copied back: the same at every N (outside the timed regions)
            "bwt_sha256": __import__("hashlib").sha256(h_out.numpy().tobytes()).hexdigest(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--cpu-sample", type=int, default=12_000_000, help="bases of the workload the cpu_baseline leg runs")
    ap.add_argument("--ref-sample", type=int, default=0, help="bases per step of the --impl reference arm (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample clocks (to measure the sampler's own cost)")
    ap.add_argument("--sharded", action="store_true", help="use the sharded (multi-GPU) code path even at N=1")
    args = ap.parse_args()
    if args.no_clocks:
        os.environ["DEBWT_NO_CLOCKS"] = "1"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if world > 1 or args.sharded:
        ours_sharded(args, rank, world, local_rank)
    else:
        ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
