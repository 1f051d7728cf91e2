"""Hot SASS instructions of an .ncu-rep source page: python tools/ncu_hot.py rep [topN]  (read here, no GPU)"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
iS, iN, iE, iA = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Address")
iW = hdr.index("L1 Wavefronts Shared")
tot = sum(int(r[iN]) for r in data)
tote = sum(int(r[iE]) for r in data)
print(f"total samples {tot}, warp instructions {tote}, shared wavefronts {sum(int(r[iW]) for r in data)}")
print("-- by samples")
for idx, r in sorted(enumerate(data), key=lambda t: -int(t[1][iN]))[:top]:
    print(f"{idx:5d} {100 * int(r[iN]) / tot:5.1f}%  exec {int(r[iE]):9d}  wf {int(r[iW]):9d}  {r[iS].strip()[:90]}")
print("-- by executed")
for idx, r in sorted(enumerate(data), key=lambda t: -int(t[1][iE]))[:top // 2]:
    print(f"{idx:5d} {100 * int(r[iE]) / tote:5.1f}%  exec {int(r[iE]):9d}  {r[iS].strip()[:90]}")
