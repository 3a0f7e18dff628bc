"""Developer timing helper (not the bench): CUDA-event time of decode_packed per BASELINE config, rotating input sets."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rtm3d_b200 import HeatmapDecoder, synth

def run(name, generic, nsets=4, iters=20, kind="randn"):
    w = synth.WORKLOADS[name]
    B, C, H, W, K = w["B"], w["C"], w["H"], w["W"], w["K"]
    dev = torch.device("cuda:0")
    sets = []
    for s in range(nsets):
        g = torch.Generator(device=dev).manual_seed(1234 + s)
        mk = lambda c, sc=1.0: torch.randn((B, c, H, W), generator=g, device=dev) * sc
        hm = mk(C)
        if kind == "trained": hm = hm * 3 - 6
        sets.append([hm, mk(16), mk(2), mk(2)])
    dec = HeatmapDecoder(0.4, K, 4.0, force_generic=True) if generic is True else HeatmapDecoder(0.4, K, 4.0, split=int(generic))
    for s in sets: dec.decode_packed(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        dec.decode_packed(sets[i % nsets])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    alg = B * (C * H * W * 4 + K * 18 * 32 + K * 100 + 4)
    print(json.dumps(dict(cfg=name, kind=kind, generic=generic, us=round(ms * 1e3, 2), img_s=round(B / ms * 1e3),
                          gbs=round(alg / ms / 1e6, 1), frac=round(alg / ms / 1e6 / 6528.4, 3))))

if __name__ == "__main__":
    for name in ("cfg2", "cfg3", "cfg4", "cfg5"):
        for generic in (True, False, 1, 2, 4):
            for kind in ("randn", "trained"):
                run(name, generic, kind=kind)
