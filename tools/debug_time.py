"""Developer tool: time the plane-streaming kernel with parts of it switched off (results are then WRONG; timing only).
    python tools/debug_time.py [cfg4] [kpt|main|fused] [copy_rows,...] [debug modes,...]
debug modes: decode_planes.cu PlaneGeom::debug (0 = the real kernel); copy_rows: rows per bulk copy (0 = whole chunk)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rtm3d_b200 import HeatmapDecoder, _native
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
which = sys.argv[2] if len(sys.argv) > 2 else "kpt"
rows_list = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
modes = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 6, 5, 2, 1, 8, 12]
w = dict(bench.WORKLOADS[name]); w["kpt"] = w["kpt"] or 9
dev = torch.device("cuda:0")
sets = [bench.make_inputs(torch, w, dev, 1234 + i) for i in range(2)]
lib = _native.lib()
for rows in rows_list:
    lib.rtm3d_debug_set_copy_rows(ctypes.c_int(rows))
    for dbg in modes:
        dec = HeatmapDecoder(0.4, w["K"], 4.0)
        dec.flags |= dbg << 24
        if which == "main":
            run = lambda i: dec.decode_packed(sets[i % 2][0])
        elif which == "kpt":
            run = lambda i: dec.decode_keypoints(sets[i % 2][1], sets[i % 2][0][3])
        else:
            run = lambda i: dec.decode_with_keypoints(sets[i % 2][0], sets[i % 2][1])
        for i in range(4): run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20): run(i)
        e1.record(); torch.cuda.synchronize()
        print(f"{name} {which} copy_rows={rows} debug={dbg}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
lib.rtm3d_debug_set_copy_rows(ctypes.c_int(0))
