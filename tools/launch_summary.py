"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST `n` launches."""
import csv, sys, re, collections
path = sys.argv[1]; last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum": continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name); name = re.sub(r"^void ", "", name)
    rows.append((name[:60], r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e3))
if last: rows = rows[-last:]
tot = sum(r[3] for r in rows)
agg = collections.OrderedDict()
for n, g, b, us in rows:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += us
print(f"{len(rows)} launches, {tot:.1f} us total (serialised, cold-cache)")
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:10.1f} us {100*us/tot:5.1f}%  x{c:<4d} {n}")
if "-v" in sys.argv:
    for i, (n, g, b, us) in enumerate(rows): print(f"{i:4d} {us:9.1f} us  grid {g:>16s} block {b:>14s}  {n}")
