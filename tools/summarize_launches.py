"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel
(count, total, share of the captured region, mean).  Usage:
    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v
    name = r[ki].replace("<unnamed>::", "").replace("void ", "")[:64]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {tot:.1f} us total "
      "(ncu per-launch times are cold-cache and serialised: compare SHARES)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:66s} n={v[0]:4d} total_us={v[1]:10.1f} share={v[1] / tot * 100:5.1f}% avg_us={v[1] / v[0]:8.1f}")
