"""Developer tool: counters and per-CTA start/end times of the scan kernel (rtm3d_debug_set_stats).
   python tools/scan_stats.py [cfg4|cfg5|cfg2] [bf16]"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rtm3d_b200 import HeatmapDecoder, _native

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
dtype = "bf16" if "bf16" in sys.argv else "f32"
w = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
logits, kpt, _ = bench.make_inputs(torch, w, dev, 1234, dtype=dtype)
dbg = int(os.environ.get("SCAN_DEBUG", "0"))
dec = HeatmapDecoder(0.4, w["K"], 4.0, reuse_outputs=True, debug=dbg)
lib = _native.lib()
lib.rtm3d_debug_set_stats.argtypes = [ctypes.c_void_p]
lib.rtm3d_debug_set_stats.restype = None
run = (lambda: dec.decode_with_keypoints(logits, kpt)) if w["kpt"] else (lambda: dec.decode_packed(logits))
for _ in range(3):
    run()
torch.cuda.synchronize()
stats = torch.zeros(64 + 2 * 160, dtype=torch.int64, device=dev)
lib.rtm3d_debug_set_stats(stats.data_ptr())
run()
torch.cuda.synchronize()
lib.rtm3d_debug_set_stats(None)
s = stats.cpu()
names = ["strips", "deepen", "exact", "list_keys"]
print({n: int(s[i]) for i, n in enumerate(names)}, "keys/strip %.1f" % (int(s[3]) / max(1, int(s[0]))))
ph = ["wait_first", "pass1", "select", "pass2", "verify", "flush"]
ns = max(1, int(s[0]))
print("cycles per strip (warp 3 of every CTA): " + "  ".join("%s %.0f" % (n, int(s[5 + i]) / ns) for i, n in enumerate(ph)),
      " | producer: wait %.0f total %.0f" % (int(s[11]) / ns, int(s[12]) / ns))
t = s[64:].view(-1, 2)
t = t[t[:, 0] > 0]
t0 = int(t[:, 0].min())
start = (t[:, 0] - t0).float() / 1e3
end = (t[:, 1] - t0).float() / 1e3
print("CTAs %d  start us: min %.1f max %.1f   end us: min %.1f median %.1f max %.1f" % (len(t), start.min(), start.max(), end.min(), end.median(), end.max()))
