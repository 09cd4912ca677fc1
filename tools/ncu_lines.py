"""Executed warp-instructions and stall samples per CUDA source line from
    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:K > f.csv
    python tools/ncu_lines.py f.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, agg = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        iE = hdr.index("Instructions Executed")
        iS = hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or len(r) <= iE:
        continue
    if r[0].strip().isdigit():
        key = (cur_file, int(r[0]), r[1].strip()[:90])
        ex = r[iE].replace(",", "")
        st = r[iS].replace(",", "")
        a = agg.setdefault(key, [0, 0])
        a[0] += int(ex) if ex.isdigit() else 0
        a[1] += int(st) if st.isdigit() else 0
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[1] for v in agg.values()) or 1
print(f"total warp-instr {tot}  stall samples {tots}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:N]:
    print(f"{v[0]:10d} {v[0] / tot * 100:5.1f}%  stalls {v[1] / tots * 100:5.1f}%  {k[0]}:{k[1]}  {k[2]}")
