import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from oracle import decode_ref
from rtm3d_b200 import HeatmapDecoder, synth
DEV="cuda:0"
for (B,kind,K,seed) in [(3,"trained",50,77),(2,"randn",100,78)]:
    logits,_ = synth.head_outputs(B,3,48,80,seed=seed,kind=kind)
    dl=[t.to(DEV) for t in logits]
    for split in (2,4):
        dec=HeatmapDecoder(0.0,K,4.0,split=split)
        p=dec.decode_packed(dl); torch.cuda.synchronize()
        for b in range(B):
            r=decode_ref.decode_image(dl[0][b],dl[1][b],dl[2][b],0.0,K,4.0)
            n=int(p.counts[b]); gf=p.flat[b,:n].cpu().numpy(); wf=r["flat"].cpu().numpy()
            gs=p.score[b,:n].cpu().numpy(); ws=r["score"].cpu().numpy()
            bad=np.nonzero(gf!=wf)[0] if n==len(wf) else None
            print(kind,K,"split",split,"img",b,"n",n,len(wf),"mismatch", None if bad is None else len(bad))
            if bad is not None and len(bad):
                print("  got ",gf[:10],gs[:5]); print("  want",wf[:10],ws[:5])
                missing=set(wf.tolist())-set(gf.tolist()); print("  missing", sorted(missing)[:10], "rows(y)", sorted({(m%(48*80))//80 for m in missing})[:20])
