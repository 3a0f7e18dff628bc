"""Developer tool: per-CTA phase timeline of the streaming kernel (argv: cfg cluster)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rtm3d_b200 import HeatmapDecoder, synth, _native
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
cluster = int(sys.argv[2]) if len(sys.argv) > 2 else 0
kind = sys.argv[3] if len(sys.argv) > 3 else "randn"
w = synth.WORKLOADS[name]
B, C, H, W, K = w["B"], w["C"], w["H"], w["W"], w["K"]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1234)
sets = [[torch.randn((B, c, H, W), generator=g, device=dev) for c in (C, 16, 2, 2)] for _ in range(3)]
if kind == "trained":
    for s in sets: s[0] = s[0] * 3 - 6
dec = HeatmapDecoder(0.4, K, 4.0, cluster=cluster)
for s in sets: dec.decode_packed(s)
torch.cuda.synchronize()
tl = torch.zeros(B * 8 * 16, dtype=torch.int64, device=dev)
lib = _native.lib()
lib.rtm3d_debug_set_timeline.argtypes = [ctypes.c_void_p]
lib.rtm3d_debug_set_timeline(tl.data_ptr())
dec.decode_packed(sets[0])
torch.cuda.synchronize()
lib.rtm3d_debug_set_timeline(None)
t = tl.view(-1, 16).cpu()
t = t[t[:, 0] > 0]
n = t.shape[0]
t0 = t[:, 0].min()
names = ["entry", "init", "first_data", "scan_done", "sel_drained", "sel_final", "post_sync", "sorted", "emitted"]
print(f"{name} cluster={cluster} kind={kind} ctas={n}")
for i, nm in enumerate(names):
    col = (t[:, i] - t0).float() / 1e3
    print(f"  {nm:12s} abs us: min {col.min():8.2f} mean {col.mean():8.2f} max {col.max():8.2f}")
for a, b, nm in [(0, 2, "entry->first_data"), (2, 3, "scan"), (3, 4, "scan_done->sel_drained"), (4, 5, "final prune"),
                 (6, 7, "sort"), (7, 8, "emit"), (0, 8, "cta total")]:
    d = (t[:, b] - t[:, a]).float() / 1e3
    print(f"  {nm:24s} us: min {d.min():7.2f} mean {d.mean():7.2f} max {d.max():7.2f}")
for i, nm in [(9, "prunes"), (10, "final_count"), (11, "candidates"), (12, "flush passes warp0")]:
    c = t[:, i].float()
    print(f"  {nm:20s}: min {c.min():.0f} mean {c.mean():.1f} max {c.max():.0f}")
