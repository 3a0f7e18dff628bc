"""Developer tool: timestamps of the plane-streaming kernel (one launch): per CTA start / init / pass-0 / end, and for CTA 0
per chunk: producer issue, A-warp wake / done, B-warp wake / stage release / chunk done.
    python tools/trace_cta.py [cfg4]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from rtm3d_b200 import HeatmapDecoder, _native
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
w = dict(bench.WORKLOADS[name]); w["kpt"] = w["kpt"] or 9
dev = torch.device("cuda:0")
sets = [bench.make_inputs(torch, w, dev, 1234 + i) for i in range(2)]
lib = _native.lib()
lib.rtm3d_debug_set_trace.argtypes = [ctypes.c_void_p]
dec = HeatmapDecoder(0.4, w["K"], 4.0)
for i in range(4): dec.decode_with_keypoints(sets[i % 2][0], sets[i % 2][1])
torch.cuda.synchronize()
NK = 40
tr = torch.zeros(NK * 96 + 4 * 256, dtype=torch.int64, device=dev)
lib.rtm3d_debug_set_trace(tr.data_ptr())
dec.decode_with_keypoints(sets[0][0], sets[0][1])
torch.cuda.synchronize()
lib.rtm3d_debug_set_trace(None)
full = tr.cpu().numpy()
ct = full[NK * 96:].reshape(256, 4).astype(np.float64)
ct = ct[ct[:, 0] > 0]
g0 = ct[:, 0].min()
print("CTAs", len(ct), "kernel span %.1f us" % ((ct[:, 3].max() - g0) / 1e3))
print("init done  : min %.1f max %.1f" % ((ct[:, 1].min() - g0) / 1e3, (ct[:, 1].max() - g0) / 1e3))
print("pass0 done : min %.1f max %.1f mean %.1f" % ((ct[:, 2].min() - g0) / 1e3, (ct[:, 2].max() - g0) / 1e3, (ct[:, 2].mean() - g0) / 1e3))
print("end        : min %.1f max %.1f mean %.1f" % ((ct[:, 3].min() - g0) / 1e3, (ct[:, 3].max() - g0) / 1e3, (ct[:, 3].mean() - g0) / 1e3))
d = (ct[:, 2] - g0) / 1e3
print("pass0 done per CTA (every 8th):", " ".join("%.0f" % x for x in d[::8]))
t = full[:NK * 96].reshape(NK, 96).astype(np.float64)
if t[0, 1] == 0: sys.exit(0)
t0 = t[0, 0]
t = np.where(t > 0, (t - t0) / 1.965e3, np.nan)     # us
print("chunk issue | A wake (4 warps, rel. issue)  | A done (rel. issue)        | B wake min/max  release min/max  chunkdone max (rel. A done max) | next issue - last release")
for q in range(0, 84):
    aw = t[1:5, q] - t[0, q]; ad = t[5:9, q] - t[0, q]
    sc = np.nanmax(t[5:9, q])
    bw = t[10:17, q] - sc; br = t[17:24, q] - sc; bd = t[24:31, q] - sc
    nxt = t[0, q + 4] - np.nanmax(t[17:24, q]) if q + 4 < 96 else np.nan
    print(f"{q:3d} {t[0,q]:7.2f} | " + " ".join(f"{x:5.2f}" for x in aw) + " | " + " ".join(f"{x:5.2f}" for x in ad) +
          f" | {np.nanmin(bw):5.2f} {np.nanmax(bw):5.2f}   {np.nanmin(br):5.2f} {np.nanmax(br):5.2f} (w{int(np.nanargmax(br))})  {np.nanmax(bd):5.2f} | {nxt:5.2f}")
