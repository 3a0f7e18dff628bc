"""Developer tool: time the plane-streaming kernel with parts of it switched off (results are then WRONG; timing only).
    python tools/debug_time.py [cfg4] [kpt|main]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rtm3d_b200 import HeatmapDecoder
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
which = sys.argv[2] if len(sys.argv) > 2 else "kpt"
w = dict(bench.WORKLOADS[name]); w["kpt"] = w["kpt"] or 9
dev = torch.device("cuda:0")
sets = [bench.make_inputs(torch, w, dev, 1234 + i) for i in range(2)]
for dbg in (0, 9, 10, 6, 5, 2, 1):
    dec = HeatmapDecoder(0.4, w["K"], 4.0)
    dec.flags |= dbg << 24
    run = (lambda i: dec.decode_packed(sets[i % 2][0])) if which == "main" else (lambda i: dec.decode_keypoints(sets[i % 2][1], sets[i % 2][0][3])) if which == "kpt" else (lambda i: dec.decode_with_keypoints(sets[i % 2][0], sets[i % 2][1]))
    for i in range(4): run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): run(i)
    e1.record(); torch.cuda.synchronize()
    print(f"{name} {which} debug={dbg} (9: no main emit, 10: no kpt emit, 6: no emit): {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
