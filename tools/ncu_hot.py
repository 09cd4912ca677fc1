"""Top stall lines / opcode histogram from `ncu --page source --csv` output.
    ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > /tmp/src.csv; python tools/ncu_hot.py /tmp/src.csv [N]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H, data = rows[hi], [r for r in rows[hi + 1:] if len(r) > 5]
isrc, iS, iE = H.index("Source"), H.index("Warp Stall Sampling (All Samples)"), H.index("Instructions Executed")
num = lambda s: int(s) if s.strip().isdigit() else 0
tot = sum(num(r[iS]) for r in data) or 1
print("samples", tot, "sass lines", len(data), "warp-instr executed", sum(num(r[iE]) for r in data))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
top = sorted(range(len(data)), key=lambda i: -num(data[i][iS]))[:N]
for i in sorted(top):
    r = data[i]
    print(f"{i:6d} {num(r[iS]):6d} {num(r[iS]) / tot * 100:5.1f}%  exec {num(r[iE]):8d}  {r[isrc][:120]}")
c, e = collections.Counter(), collections.Counter()
for r in data:
    t = r[isrc].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c[op] += num(r[iS])
    e[op] += num(r[iE])
print("stall samples by opcode:", [(k, f"{v / tot * 100:.1f}%") for k, v in c.most_common(16)])
print("executed by opcode:", [(k, v) for k, v in e.most_common(20)])
