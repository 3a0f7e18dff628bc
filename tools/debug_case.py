"""Developer tool: run one small case through a kernel variant and print the first mismatching rows vs the CUDA oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from oracle import decode_ref
from rtm3d_b200 import HeatmapDecoder, synth
B, C, H, W, K = [int(x) for x in sys.argv[1:6]]
kind = sys.argv[6]; cluster = int(sys.argv[7]); seed = int(sys.argv[8]) if len(sys.argv) > 8 else 1000 + H * 7 + W
logits, _ = synth.head_outputs(B, C, H, W, seed=seed, kind=kind)
dev = [t.cuda() for t in logits]
dec = HeatmapDecoder(0.4, K, 4.0, cluster=cluster)
dbg = int(os.environ.get('RTM3D_DEBUG', '0'))
dec.flags |= dbg << 16
reps = int(os.environ.get('REPS', '3'))
nbad = 0
for rep in range(reps):
    p = dec.decode_packed(dev); torch.cuda.synchronize()
    for b in range(B):
        r = decode_ref.decode_image(dev[0][b], dev[1][b], dev[2][b], 0.4, K, 4.0)
        n = int(p.counts[b]); want_n = 0 if r is None else len(r["cls"])
        gf = p.flat[b, :n].cpu().numpy(); gs = p.score[b, :n].cpu().numpy()
        wf = r["flat"].cpu().numpy() if r else np.zeros(0); ws = r["score"].cpu().numpy() if r else np.zeros(0)
        ok = n == want_n and np.array_equal(gf, wf) and np.array_equal(gs, ws)
        nbad += (not ok)
        if not ok and nbad <= 2:
            print(f"rep {rep} image {b}: n={n} want={want_n} ok={ok}")
            m = min(n, want_n)
            bad = [i for i in range(m) if gf[i] != wf[i] or gs[i] != ws[i]][:12]
            for i in bad:
                print(f"   row {i}: got flat {gf[i]} score {gs[i]:.9f} | want flat {wf[i]} score {ws[i]:.9f}")
            print("   missing:", sorted(set(wf.tolist()) - set(gf.tolist()))[:10], "extra:", sorted(set(gf.tolist()) - set(wf.tolist()))[:10])
            dup = len(gf) - len(set(gf.tolist())); print("   duplicates:", dup)

print(f'debug={dbg} cluster={cluster}: {nbad} bad of {reps*B}')
