import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from rtm3d_b200 import fit_packed
from oracle import boxfit_ref as bf
g=np.load("tests/golden/boxfit_golden.npz"); dev=torch.device("cuda:0"); n=len(g["cls"])
fit=fit_packed(torch.as_tensor(g["uv"],device=dev).reshape(1,n,8,2), torch.as_tensor(g["cls"],device=dev).reshape(1,n), None, torch.as_tensor(g["K"],device=dev), g["dim_ref"], list(g["ref_loc"]), want_solution=True)
torch.cuda.synchronize()
fun=fit.fun[0].cpu().numpy(); it=fit.iters[0].cpu().numpy(); x8=fit.x8[0].cpu().numpy()
for i in range(n):
    c1=bf.canonical(x8[i]); c0=bf.canonical(g["x"][i])
    print(i, "ref %.5f mine %.5f it %d  Ry ref %.3f mine %.3f  Z ref %.2f mine %.2f" % (g["fun"][i], fun[i], it[i], c0[0], c1[0], c0[6], c1[6]))
