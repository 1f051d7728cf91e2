"""Hot CUDA source lines of an .ncu-rep (read here, no GPU): python tools/ncu_lines.py rep [topN]
Per source line: share of stall samples, of executed warp instructions, of shared-memory wavefronts, L2 sectors."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if "# Samples" in r][0]
hdr = rows[hi]
iL, iS, iN, iE = 0, 1, hdr.index("# Samples"), hdr.index("Instructions Executed")
def col(name):                      # a kernel without shared memory / global accesses has no such column
    return hdr.index(name) if name in hdr else None


iW, iG = col("L1 Wavefronts Shared"), col("L2 Theoretical Sectors Global")


def val(r, i):
    return int(r[i]) if i is not None and r[i] not in ("", "n/a") else 0
fname = ""
lines = []
for r in rows[:hi] + rows[hi + 1:]:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if len(r) == len(hdr) and r[iL] not in ("", "Line No"):
        lines.append((fname, r))
tot = sum(int(r[iN]) for _, r in lines) or 1
tote = sum(int(r[iE]) for _, r in lines) or 1
totw = sum(val(r, iW) for _, r in lines) or 1
totg = sum(val(r, iG) for _, r in lines) or 1
print(f"samples {tot}, warp instructions {tote}, shared wavefronts {totw}, L2 sectors {totg}")
print("  smp%   ins%  smem%    L2%  line")
for f, r in sorted(lines, key=lambda t: -int(t[1][iN]))[:top]:
    print(f"{100 * int(r[iN]) / tot:5.1f}  {100 * int(r[iE]) / tote:5.1f}  {100 * val(r, iW) / totw:5.1f}  {100 * val(r, iG) / totg:5.1f}  "
          f"{f}:{r[iL]}  {r[iS].strip()[:100]}")
