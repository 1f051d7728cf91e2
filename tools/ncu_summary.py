"""Summarise an .ncu-rep (read here, no GPU): python tools/ncu_summary.py rep [kernel-substring]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed.sum",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed"]
for r in data:
    if flt not in r[ki]:
        continue
    print("==", r[ki][:90])
    for k in keys:
        if k in hdr:
            print(f"  {k:75s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
    st = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr)
          if "issue_stalled" in h and h.endswith("_per_warp_active.pct") and r[i] not in ("", "n/a")]
    for v, h in sorted(st, reverse=True)[:8]:
        print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__warps_issue_stalled_', ''):60s} {v:.1f}")
    break
