"""Developer tool: phase timestamps of the select + post kernel (DEV build).   python tools/post_stats.py [cfg4|cfg5|cfg2]"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rtm3d_b200 import HeatmapDecoder, _native
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
w = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
logits, kpt, _ = bench.make_inputs(torch, w, dev, 1234)
dec = HeatmapDecoder(0.4, w["K"], 4.0, reuse_outputs=True)
lib = _native.lib()
lib.rtm3d_debug_set_stats.argtypes = [ctypes.c_void_p]; lib.rtm3d_debug_set_stats.restype = None
for _ in range(3):
    dec.decode_with_keypoints(logits, kpt)
torch.cuda.synchronize()
n_cta = w["B"] * 4
stats = torch.zeros(1024 + n_cta * 8, dtype=torch.int64, device=dev)
lib.rtm3d_debug_set_stats(stats.data_ptr())
dec.decode_with_keypoints(logits, kpt)
torch.cuda.synchronize()
lib.rtm3d_debug_set_stats(None)
t = stats[1024:].view(n_cta, 8).cpu().double()
names = ["start->sync1", "check+sort", "merge", "emit+gather", "sync2", "group", "bbox", "-"]
d = t[:, 1:] - t[:, :-1]
t0 = t[:, 0].min()
print("CTA start us: min 0 median %.1f max %.1f ; end us: min %.1f median %.1f max %.1f" % ((t[:,0].median()-t0)/1e3, (t[:,0].max()-t0)/1e3, (t[:,7].min()-t0)/1e3, (t[:,7].median()-t0)/1e3, (t[:,7].max()-t0)/1e3))
for rank in range(4):
    r = d[rank::4]
    print("rank", rank, "  ".join("%s %.0f" % (n, r[:, i].median()) for i, n in enumerate(names[:7])), " total %.0f" % (t[rank::4, 7] - t[rank::4, 0]).median())

import numpy as np
tot = (t[:, 7] - t[:, 0]).numpy() / 1e3
srt = d[:, 1].numpy() / 1e3
img_end = ((t[:, 7] - t0).view(-1, 4).max(dim=1).values / 1e3).numpy()
for name, v in (("CTA total us", tot), ("check+sort us", srt), ("image end us", img_end)):
    print(name, " ".join("p%d %.1f" % (q, np.percentile(v, q)) for q in (10, 50, 75, 90, 99, 100)))
slow = srt.reshape(-1, 4).max(axis=1) > 1.5 * np.median(srt)
print("images with a slow sort: %.1f %%; their end: median %.1f vs others %.1f" % (100 * slow.mean(), np.median(img_end[slow]) if slow.any() else 0, np.median(img_end[~slow])))
