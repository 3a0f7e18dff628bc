"""Developer tool: aggregate an `ncu --page source --csv` export into runs of SASS instructions with the same execution
count (hot regions of a kernel).   ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_regions.py src.csv [units]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]; data = rows[2:]
ia = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); isamp = hdr.index("# Samples")
tot = sum(int(r[ia]) for r in data); ts = sum(int(r[isamp]) for r in data)
print("SASS instructions", len(data), "executed", tot, "samples", ts, "executed per unit", tot / units)
prev = None; start = 0; acc = 0; samp = 0; regions = []
for i, r in enumerate(data):
    c = int(r[ia])
    if prev is None or abs(c - prev) > 0.02 * max(prev, 1) + 10:
        if prev is not None: regions.append((start, i - 1, prev, acc, samp))
        start = i; acc = 0; samp = 0
    acc += c; samp += int(r[isamp]); prev = c
regions.append((start, len(data) - 1, prev, acc, samp))
for s, e, c, a, sm in regions:
    if a > 0.01 * tot or sm > 0.01 * ts:
        print(f"[{s:5d}-{e:5d}] n={e-s+1:4d} exec/instr={c/units:8.1f} total={100*a/tot:5.1f}% samples={100*sm/ts:5.1f}%  {data[s][isrc].strip()[:50]}")
