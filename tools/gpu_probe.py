"""First-contact GPU probe: facts the decode kernels' parity contract rests on.

Writes gpurun_out/probe.json.  Not part of the product or the tests.
"""
import ctypes, json, os, sys
import torch
import torch.nn.functional as F

out = {}
dev = torch.device("cuda:0")
p = torch.cuda.get_device_properties(0)
out["device"] = dict(name=p.name, sms=p.multi_processor_count, l2=p.L2_cache_size, mem=p.total_memory,
                     cc=[p.major, p.minor])
out["host"] = dict(cpus=os.cpu_count(), ref_mounted=os.path.exists("/root/reference/models/model.py"),
                   torch_threads=torch.get_num_threads())
try:
    with open("/proc/cpuinfo") as f:
        out["host"]["cpu_model"] = [l.split(":")[1].strip() for l in f if l.startswith("model name")][0]
except Exception as e:  # noqa
    out["host"]["cpu_model"] = str(e)

lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libprobe.so"))
lib.probe_sigmoid.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]


def mine(x):
    y = torch.empty_like(x)
    rc = lib.probe_sigmoid(x.data_ptr(), y.data_ptr(), x.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    torch.cuda.synchronize()
    return y


def ulp_diff(a, b):
    ai = a.view(torch.int32).to(torch.int64)
    bi = b.view(torch.int32).to(torch.int64)
    return (ai - bi).abs()


g = torch.Generator(device="cpu").manual_seed(7)
sig = {}
for name, x in {
    "randn3": torch.randn(4_000_000, generator=g) * 3,
    "randn12": torch.randn(4_000_000, generator=g) * 12,
    "uniform_-100_100": (torch.rand(4_000_000, generator=g) - 0.5) * 200,
    "tiny": torch.randn(1_000_000, generator=g) * 1e-6,
}.items():
    xc = x.to(dev)
    t_cuda = torch.sigmoid(xc)
    m_cuda = mine(xc)
    t_cpu = torch.sigmoid(x)
    d1 = ulp_diff(t_cuda, m_cuda)
    d2 = ulp_diff(t_cuda.cpu(), t_cpu)
    # in-place variant as the reference uses it
    t_inpl = xc.clone().sigmoid_()
    sig[name] = dict(kernel_vs_torchcuda_ndiff=int((d1 > 0).sum()), kernel_vs_torchcuda_maxulp=int(d1.max()),
                     torchcuda_vs_cpu_ndiff=int((d2 > 0).sum()), torchcuda_vs_cpu_maxulp=int(d2.max()),
                     inplace_equal=bool(torch.equal(t_inpl, t_cuda)))
out["sigmoid"] = sig

# monotonicity of the CUDA sigmoid over dense sorted inputs
mono = {}
for name, lo, hi in [("-20..20", -20.0, 20.0), ("-90..-60", -90.0, -60.0), ("0..2", 0.0, 2.0), ("2..18", 2.0, 18.0)]:
    xs = torch.linspace(lo, hi, 16_000_001, device=dev, dtype=torch.float64).float().unique(sorted=True)
    s = torch.sigmoid(xs)
    viol = int((s[1:] < s[:-1]).sum())
    mono[name] = dict(n=int(xs.numel()), violations=viol)
# every float in a narrow band (consecutive bit patterns)
base = torch.arange(0, 8_000_000, device=dev, dtype=torch.int32)
for name, start in [("bits@1.0", 0x3F800000), ("bits@-1.0", 0xBF800000 - 8_000_000), ("bits@8.0", 0x41000000),
                    ("bits@0.01", 0x3C23D70A), ("bits@-87", 0xC2AE0000 - 8_000_000)]:
    xi = (base + (start if start < 2 ** 31 else start - 2 ** 32)).view(torch.float32)
    xs, _ = torch.sort(xi)
    s = torch.sigmoid(xs)
    mono[name] = dict(n=int(xs.numel()), violations=int((s[1:] < s[:-1]).sum()))
out["sigmoid_monotone"] = mono

# topk tie behaviour on CUDA
tk = {}
for (n, k) in [(92160, 50), (92160, 100), (368640, 100), (30720, 50), (30720, 100)]:
    v = (torch.randint(0, 200, (n,), generator=g).float() / 200).to(dev)
    sc, ix = torch.topk(v, k)
    sc_c, ix_c = sc.cpu(), ix.cpu()
    asc = desc = 0
    for j in range(k - 1):
        if sc_c[j] == sc_c[j + 1]:
            if ix_c[j] < ix_c[j + 1]:
                asc += 1
            else:
                desc += 1
    kth = sc_c[-1]
    tie_idx = torch.nonzero(v.cpu() == kth).flatten()
    n_in = int((sc_c == kth).sum())
    lowest = set(tie_idx[:n_in].tolist())
    got = set(ix_c[sc_c == kth].tolist())
    tk[f"1d_{n}_{k}"] = dict(tie_pairs_asc=asc, tie_pairs_desc=desc, boundary_keeps_lowest=bool(lowest == got),
                             boundary_group=int(tie_idx.numel()), boundary_taken=n_in)
vb = (torch.randint(0, 200, (64, 92160), generator=g).float() / 200).to(dev)
sc, ix = torch.topk(vb, 100, dim=-1)
sc_c, ix_c = sc.cpu(), ix.cpu()
same = sc_c[:, :-1] == sc_c[:, 1:]
asc = int((same & (ix_c[:, :-1] < ix_c[:, 1:])).sum())
desc = int((same & (ix_c[:, :-1] > ix_c[:, 1:])).sum())
tk["batched_64x92160_100"] = dict(tie_pairs_asc=asc, tie_pairs_desc=desc)
# all-equal plateau: which indices does CUDA topk return
pl = torch.ones(92160, device=dev)
sc, ix = torch.topk(pl, 50)
tk["plateau_ones_first50"] = ix.cpu().tolist()[:50]
sc, ix = torch.topk(torch.ones(92160), 50)
tk["plateau_ones_first50_cpu"] = ix.tolist()[:50]
out["topk"] = tk

# nms parity cpu vs cuda on randn
hm = torch.randn(4, 3, 96, 320, generator=g)
def nms(h):
    s = h.clone().sigmoid_()
    m = F.max_pool2d(s, 3, 1, 1)
    return s * (m == s).float()
a = nms(hm)
b = nms(hm.to(dev)).cpu()
out["nms"] = dict(mask_equal=bool(torch.equal(a > 0, b > 0)), score_ndiff=int((a != b).sum()),
                  top100_idx_equal=[bool(torch.equal(torch.topk(a[i].view(-1), 100)[1], torch.topk(b[i].view(-1), 100)[1]))
                                    for i in range(4)])
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/probe.json", "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
