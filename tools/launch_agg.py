"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/launch_agg.py file.csv [-v]"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = collections.OrderedDict(); seq = []
for row in csv.DictReader(lines):
    name = re.sub(r"unnamed>::", "", row["Kernel Name"]).split("(")[0].replace("void ", "")
    t = float(row["Metric Value"]) / 1e3
    a = agg.setdefault(name, [0, 0.0, 0.0]); a[0] += 1; a[1] += t; a[2] = max(a[2], t)
    seq.append((name, t, row["Grid Size"]))
tot = sum(v[1] for v in agg.values())
print(f"total {tot/1e3:.2f} ms over {len(seq)} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:44s} n={v[0]:4d} total={v[1]/1e3:9.2f} ms  {100*v[1]/tot:5.1f} %  max={v[2]/1e3:8.2f} ms")
if "-v" in sys.argv:
    for s in seq: print(f"    {s[0]:44s} {s[1]/1e3:9.3f} ms grid={s[2]}")
