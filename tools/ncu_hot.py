"""Developer tool: hottest SASS regions of a kernel in an ncu report (needs --import-source on, -lineinfo).
    python tools/ncu_hot.py gpurun_out/prof_x.ncu-rep <kernel-regex> <nth-launch> [block]"""
import csv, io, subprocess, sys
rep, rx, nth = sys.argv[1], sys.argv[2], sys.argv[3]
B = int(sys.argv[4]) if len(sys.argv) > 4 else 32
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{rx}:{nth}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
segs, cur = [], None
for r in rows:
    if not r: continue
    if r[0] in ("Kernel Name", "Function Name"):
        cur = {"name": r[1][:90], "rows": []}; segs.append(cur); continue
    if r[0] == "Address": cur["h"] = r; continue
    if cur is not None: cur["rows"].append(r)
seen, uniq = set(), []
for s in segs:
    key = (s["name"], len(s["rows"]))
    if key in seen: continue
    seen.add(key); uniq.append(s)
tot = 0
for s in uniq:
    ie = s["h"].index("Instructions Executed")
    s["tot"] = sum(int(r[ie]) for r in s["rows"] if len(r) > ie); tot += s["tot"]
print("total warp instructions", tot)
for s in uniq: print(f"  {s['name']:92s} sass={len(s['rows']):6d} share={s['tot']/tot:.3f}")
for s in uniq:
    if s["tot"] / tot < 0.03: continue
    h = s["h"]; ie, isrc, iss = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    data = [(r[isrc].strip(), int(r[ie]), int(r[iss])) for r in s["rows"]]
    print("==", s["name"])
    i = 0
    # contiguous regions with similar execution count
    while i < len(data):
        j = i
        e0 = data[i][1]
        while j < len(data) and (data[j][1] == e0 or (e0 > 0 and 0.5 < data[j][1] / max(e0, 1) < 2.0)): j += 1
        e = sum(d[1] for d in data[i:j]); sm = sum(d[2] for d in data[i:j])
        if e / tot >= 0.01:
            ops = {}
            for d in data[i:j]:
                t = d[0].split(); op = t[1] if t[0].startswith("@") else t[0]
                ops[op] = ops.get(op, 0) + 1
            top = ", ".join(f"{k}x{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:9])
            print(f"  sass[{i:5d}:{j:5d}] n={j-i:4d} exec/instr={e0:9d} share={e/tot:6.3f} samples={sm:6d} | {top}")
        i = j

# ---- same regions ranked by stall samples (where warps WAIT)
for s in uniq:
    if s["tot"] / tot < 0.03: continue
    h = s["h"]; ie, isrc, iss = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    data = [(r[isrc].strip(), int(r[ie]), int(r[iss])) for r in s["rows"]]
    ts = sum(d[2] for d in data)
    print("== by samples (total %d)" % ts)
    top = sorted(range(len(data)), key=lambda i: -data[i][2])[:28]
    for i in sorted(top):
        print(f"  sass[{i:5d}] samples={data[i][2]:6d} ({data[i][2]/ts:5.3f}) exec={data[i][1]:9d}  {data[i][0][:70]}")
