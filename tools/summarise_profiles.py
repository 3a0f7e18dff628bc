"""Turn gpurun_out/ captures into the small tracked summaries under profiles/ (run here, no GPU needed).
    python tools/summarise_profiles.py <tag>
reads  gpurun_out/launches_<tag>.csv, gpurun_out/prof_<tag>.ncu-rep, gpurun_out/bench*_<tag>.json
writes profiles/<tag>_launches.csv (per-kernel device time + share), profiles/<tag>_ncu_full.txt, profiles/<tag>_bench.jsonl
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
        f.write("# median_us: the first launch on a fresh workspace has no remembered thresholds and runs cold (~4x)\n")
        f.write("kernel,grid,block,launches,mean_us,median_us,share_of_all,share_of_rtm3d\n")
        for (k, g, b), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            s_m = sum(v) / mine if "rtm3d::" in k and mine else 0.0
            f.write(f"\"{k[:110]}\",\"{g}\",\"{b}\",{len(v)},{sum(v) / len(v) / 1e3:.2f},{sorted(v)[len(v) // 2] / 1e3:.2f},{sum(v) / tot:.4f},{s_m:.4f}\n")
    print(open(os.path.join(pr, f"{tag}_launches.csv")).read())

rep = os.path.join(go, f"prof_{tag}.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[0]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
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
    with open(os.path.join(pr, f"{tag}_ncu_full.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on (one capture per kernel; replayed, cold cache)\n")
        for r in rows[2:]:
            for w in want:
                if w in h:
                    f.write(f"{w} = {r[h.index(w)][:120]} {rows[1][h.index(w)]}\n")
            f.write("\n")
    print(open(os.path.join(pr, f"{tag}_ncu_full.txt")).read())

with open(os.path.join(pr, f"{tag}_bench.jsonl"), "w") as f:
    for p in sorted(glob.glob(os.path.join(go, f"bench*_{tag}.json"))):
        for line in open(p):
            if line.strip().startswith("{"):
                f.write(line)

# dram traffic of the plane-streaming kernel per launch -> profiles/traffic.json (bench.py reports it as roofline.traffic)
if os.path.exists(rep):
    import json
    tr = {}
    for r in rows[2:]:
        if "decode_planes" in r[h.index("Kernel Name")]:
            rd, wr = r[h.index("dram__bytes_read.sum")], r[h.index("dram__bytes_write.sum")]
            unit_r, unit_w = rows[1][h.index("dram__bytes_read.sum")], rows[1][h.index("dram__bytes_write.sum")]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tr = {"bytes": int(float(rd.replace(",", "")) * scale[unit_r] + float(wr.replace(",", "")) * scale[unit_w]),
                  "source": f"profiles/{tag}_ncu_full.txt (ncu --set full, one launch of decode_planes_kernel, default bench workload cfg4)"}
    if tr:
        path = os.path.join(pr, "traffic.json")
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur["cfg4"] = tr
        json.dump(cur, open(path, "w"), indent=1)
        print("traffic", tr)
