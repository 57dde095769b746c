"""Per-kernel totals of an ncu launch-list CSV (gpu__time_duration.sum): python tools/launch_summary.py <csv> [steps-in-capture] [--seq]"""
import collections
import csv
import sys

f = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 2
rows = [r for r in csv.reader(open(f)) if len(r) > 5]
hdr = rows[0]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
data = rows[1:]
if "--seq" in sys.argv:
    for r in data[len(data) - len(data) // steps:]:
        print(f"{float(r[vi].replace(',', '')) / 1e3:8.1f}  {r[gi]:>16} {r[bi]:>14}  {r[ki][:100]}")
    sys.exit(0)
tot, cnt = collections.OrderedDict(), collections.Counter()
for r in data:
    n = r[ki][:80]
    tot[n] = tot.get(n, 0.0) + float(r[vi].replace(",", ""))
    cnt[n] += 1
s = sum(tot.values())
for n, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{v / 1e3 / steps:9.1f} us/step {cnt[n] / steps:4.1f}x  {v / s * 100:5.1f}%  {n}")
print(f"total {s / 1e3 / steps:.1f} us/step")
