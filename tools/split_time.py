"""Developer tool: time the main-only / fused decode for each strip split.   python tools/split_time.py [cfg4main] [main|fused]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rtm3d_b200 import HeatmapDecoder
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4main"
which = sys.argv[2] if len(sys.argv) > 2 else "main"
if name.startswith("b"):      # ad hoc: b<B>[x<H>x<W>]  e.g. b1, b4, b8x192x640
    parts = name[1:].split("x")
    w = dict(B=int(parts[0]), C=3, H=int(parts[1]) if len(parts) > 1 else 96, W=int(parts[2]) if len(parts) > 2 else 320, K=50, kpt=9, note=name)
else:
    w = dict(bench.WORKLOADS[name])
if which == "fused": w["kpt"] = w["kpt"] or 9
dev = torch.device("cuda:0")
sets = [bench.make_inputs(torch, w, dev, 1234 + i) for i in range(2)]
for split in (0, 1, 2, 4, 8):
    dec = HeatmapDecoder(0.4, w["K"], 4.0, split=split)
    run = (lambda i: dec.decode_packed(sets[i % 2][0])) if which == "main" else (lambda i: dec.decode_with_keypoints(sets[i % 2][0], sets[i % 2][1]))
    for i in range(4): run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): run(i)
    e1.record(); torch.cuda.synchronize()
    print(f"{name} {which} split={split}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
