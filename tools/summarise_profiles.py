"""Turn gpurun_out/ captures into the small tracked summaries under profiles/ (run here, no GPU needed).
    python tools/summarise_profiles.py <tag>
reads  gpurun_out/launches_<tag>.csv, gpurun_out/prof_<tag>*.ncu-rep (one capture per kernel), gpurun_out/bench*_<tag>.json
writes profiles/<tag>_launches.csv (per-kernel device time + share), profiles/<tag>_ncu_full.txt, profiles/<tag>_bench.jsonl,
       profiles/<tag>_sass_hist.txt (opcode histogram of the two hot kernels in the built library) and profiles/traffic.json
"""
import collections, csv, glob, io, os, subprocess, sys

tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
os.makedirs(pr, exist_ok=True)

lc = os.path.join(go, f"launches_{tag}.csv")
if os.path.exists(lc):
    rows = list(csv.reader(open(lc)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) > vi:
            agg.setdefault((r[ki], r[gi], r[bi]), []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    mine = sum(sum(v) for k, v in agg.items() if "rtm3d::" in k[0])
    with open(os.path.join(pr, f"{tag}_launches.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# share_of_rtm3d = share among this repo's kernels (torch's randn input generation excluded)\n")
        f.write("# the end-to-end leg of bench.py gathers from PINNED HOST memory (zero-copy over PCIe): its select_post launches take\n")
        f.write("# milliseconds under ncu and dominate the means; share_by_median = launches x median duration among this repo's kernels\n")
        f.write("kernel,grid,block,launches,mean_us,median_us,share_of_all,share_of_rtm3d,share_by_median\n")
        med = lambda v: sorted(v)[len(v) // 2]
        mine_med = sum(len(v) * med(v) for k, v in agg.items() if "rtm3d::" in k[0])
        for (k, g, b), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            s_m = sum(v) / mine if "rtm3d::" in k and mine else 0.0
            s_d = len(v) * med(v) / mine_med if "rtm3d::" in k and mine_med else 0.0
            f.write(f"\"{k[:110]}\",\"{g}\",\"{b}\",{len(v)},{sum(v) / len(v) / 1e3:.2f},{med(v) / 1e3:.2f},{sum(v) / tot:.4f},{s_m:.4f},{s_d:.4f}\n")
    print(open(os.path.join(pr, f"{tag}_launches.csv")).read())

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "gpu__time_duration.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
reps = sorted(glob.glob(os.path.join(go, f"prof_{tag}*.ncu-rep")))
captures = []          # (header, units, row) per captured kernel launch
for rp in reps:
    out = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rws = list(csv.reader(io.StringIO(out)))
    captures += [(rws[0], rws[1], r) for r in rws[2:]]
if captures:
    want = WANT
    with open(os.path.join(pr, f"{tag}_ncu_full.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on (one capture per kernel; replayed, cold cache)\n")
        for h, units, r in captures:
            for w in want:
                if w in h:
                    f.write(f"{w} = {r[h.index(w)][:120]} {units[h.index(w)]}\n")
            f.write("\n")
    print(open(os.path.join(pr, f"{tag}_ncu_full.txt")).read())

with open(os.path.join(pr, f"{tag}_bench.jsonl"), "w") as f:
    for p in sorted(glob.glob(os.path.join(go, f"bench*_{tag}.json"))):
        for line in open(p):
            if line.strip().startswith("{"):
                f.write(line)

# dram traffic per launch of the two kernels -> profiles/traffic.json (bench.py reports the dominant kernel's as roofline.traffic)
if captures:
    import json
    tr = {}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for h, units, r in captures:
        name = r[h.index("Kernel Name")]
        if "scan_planes" in name or "select_post" in name:
            ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
            kn = "scan_planes_kernel" if "scan_planes" in name else "select_post_kernel"
            tr["cfg4" if kn == "scan_planes_kernel" else "cfg4/select_post"] = {
                "bytes": int(float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]),
                "source": f"profiles/{tag}_ncu_full.txt (ncu --set full, one launch of {kn}, default bench workload cfg4)"}
    if tr:
        path = os.path.join(pr, "traffic.json")
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur.update(tr)
        json.dump(cur, open(path, "w"), indent=1)
        print("traffic", tr)

# opcode histogram of the hot kernels (what the SM actually executes: UBLKCP = cp.async.bulk, SYNCS = mbarrier, REDUX, ...)
so = os.path.join(root, "rtm3d_b200", "librtm3d_decode.so")
if os.path.exists(so):
    import re
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    with open(os.path.join(pr, f"{tag}_sass_hist.txt"), "w") as f:
        f.write("# cuobjdump -sass rtm3d_b200/librtm3d_decode.so: instructions per opcode (static count), fp32 instantiations\n")
        for kern in ("scan_planes_kernelIf", "select_post_kernelIf"):
            m = re.search(r"Function : (\S*" + kern + r"\S*)(.*?)(?=Function : |\Z)", out, re.S)
            if not m:
                continue
            ops = collections.Counter(x.split(".")[0] for x in re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", m.group(2)))
            f.write(f"\n{m.group(1)}: {sum(ops.values())} instructions\n")
            f.write("  " + "  ".join(f"{k} {v}" for k, v in ops.most_common()) + "\n")
    print(open(os.path.join(pr, f"{tag}_sass_hist.txt")).read()[:1500])
