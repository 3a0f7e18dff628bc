"""Developer helper for ncu: a few decode launches of one BASELINE config (argv: cfg cluster iters)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rtm3d_b200 import HeatmapDecoder, synth
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
cluster = int(sys.argv[2]) if len(sys.argv) > 2 else 0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
w = synth.WORKLOADS[name]
B, C, H, W, K = w["B"], w["C"], w["H"], w["W"], w["K"]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1234)
sets = [[torch.randn((B, c, H, W), generator=g, device=dev) for c in (C, 16, 2, 2)] for _ in range(2)]
dec = HeatmapDecoder(0.4, K, 4.0, cluster=cluster)
for i in range(iters):
    out = dec.decode_packed(sets[i % 2])
torch.cuda.synchronize()
print("counts", out.counts[:4].tolist())
