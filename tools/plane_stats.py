"""Developer tool: counters of the plane-streaming kernel for one decode of a BASELINE workload.
    python tools/plane_stats.py [cfg4] [main|kpt] [split] [speculate 0/1]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rtm3d_b200 import HeatmapDecoder, _native
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
which = sys.argv[2] if len(sys.argv) > 2 else "kpt"
split = int(sys.argv[3]) if len(sys.argv) > 3 else 0
spec = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
dbg = int(sys.argv[5]) if len(sys.argv) > 5 else 0
w = dict(bench.WORKLOADS[name]); w["kpt"] = w["kpt"] or 9
dev = torch.device("cuda:0")
logits, kpt = bench.make_inputs(torch, w, dev, 1234)
dec = HeatmapDecoder(0.4, w["K"], 4.0, split=split, speculate=spec)
dec.flags |= dbg << 24
run = (lambda: dec.decode_packed(logits)) if which == "main" else (lambda: dec.decode_keypoints(kpt, logits[3])) if which == "kpt" else (lambda: dec.decode_with_keypoints(logits, kpt))
for _ in range(0 if os.environ.get("COLD") else 3): run()
torch.cuda.synchronize()
st = torch.zeros(64 + 2048, dtype=torch.int64, device=dev)
lib = _native.lib()
lib.rtm3d_debug_set_stats.argtypes = [ctypes.c_void_p]
lib.rtm3d_debug_set_stats(st.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
lib.rtm3d_debug_set_stats(None)
names = ["items", "retried", "wl_entries", "batches", "pushed", "updates", "compactions", "wait_buf_free", "wait_full",
         "wait_scanned", "fin_busy", "fin_wait", "a_total", "prod_wait", "b_busy", "b_total", "fin_boundary", "fin_compact",
         "fin_release", "fin_sort", "fin_publish", "fin_emit", "a_loop", "a_setup"]
v = st.cpu().tolist()
trace = v[64:]
v = v[:32]
items = max(v[0], 1)
print(f"{name} {which} split={split} spec={spec}: {e0.elapsed_time(e1)*1e3:.1f} us (instrumented)")
for n, x in zip(names, v):
    per = x / items
    extra = ""
    if n in ("wait_buf_free", "wait_full", "a_total", "a_loop", "a_setup"): extra = f"  per A-warp-item {x/items/8:.0f} clk"
    if n in ("wait_scanned", "b_busy", "b_total"): extra = f"  per B-warp-item {x/items/7:.0f} clk"
    print(f"  {n:14s} {x:14d}  per item {per:10.1f}{extra}")
