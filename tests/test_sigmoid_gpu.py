"""The library's sigmoid over EVERY fp32 input: bit-identical to torch's CUDA sigmoid (models/model.py:85 runs exactly
that in detect.py) and monotone non-decreasing (the kernels decide the 3x3 peak test of utils/model_utils.py:17-26 on
the LARGEST neighbour logit, which is only valid for a monotone sigmoid)."""
import pytest
import torch

from rtm3d_b200 import _native

pytestmark = pytest.mark.gpu


def test_sigmoid_all_floats_bit_exact_and_monotone():
    lib = _native.lib()
    dev = torch.device("cuda:0")
    n = 1 << 26
    prev_last = None
    y = torch.empty(n, dtype=torch.float32, device=dev)
    # ordered index o in [0, 2^32) -> float bits, ascending in value: negative floats first (most negative = o 0)
    lo_o = (0xFF800000 ^ 0xFFFFFFFF)          # -inf
    hi_o = 0x7F800000 | 0x80000000            # +inf
    for start in range(0, 1 << 32, n):
        o = torch.arange(start, start + n, dtype=torch.int64, device=dev)
        keep = (o >= lo_o) & (o <= hi_o)      # drop the NaN encodings at both ends
        bits = torch.where(o >= (1 << 31), o ^ (1 << 31), o ^ 0xFFFFFFFF)
        x = (bits & 0xFFFFFFFF).to(torch.int64)
        x = torch.where(x >= (1 << 31), x - (1 << 32), x).to(torch.int32).view(torch.float32)
        assert lib.rtm3d_sigmoid_f32(x.data_ptr(), y.data_ptr(), n, torch.cuda.current_stream().cuda_stream) == 0
        ref = torch.sigmoid(x)
        same = (y.view(torch.int32) == ref.view(torch.int32)) | ~keep
        assert bool(same.all()), f"library sigmoid differs from torch.sigmoid in chunk {start:#x}"
        yk = y[keep]
        if yk.numel() == 0:
            continue
        assert bool((yk[1:] >= yk[:-1]).all()), f"sigmoid not monotone inside chunk {start:#x}"
        if prev_last is not None:
            assert float(yk[0]) >= prev_last
        prev_last = float(yk[-1])
    assert prev_last == 1.0


def test_running_threshold_bounds_hold_for_every_bin():
    """Contract of the scan's logit threshold (rtm3d_threshold_table): every float below the bound has a sigmoid strictly
    below the bin's score edge -- checked on the 4096 floats just below each bound, and the bound is not uselessly low."""
    import ctypes
    lib = _native.lib()
    dev = torch.device("cuda:0")
    n = ctypes.c_int(0)
    assert lib.rtm3d_threshold_table(None, None, 0, ctypes.byref(n), None) == 0 and n.value > 1000
    nb = n.value
    T = torch.empty(nb, dtype=torch.float32, device=dev)
    edge = torch.empty(nb, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.rtm3d_threshold_table(T.data_ptr(), edge.data_ptr(), nb, ctypes.byref(n), st) == 0
    torch.cuda.synchronize()
    assert float(T[0]) == float("-inf") and int(edge[0]) == 0
    Tb, eb = T[1:], edge[1:].view(torch.float32)
    assert bool(torch.isfinite(Tb).all()) and bool((Tb[1:] >= Tb[:-1]).all())
    # ordered-integer walk: the m-th float below T
    ti = Tb.view(torch.int32).to(torch.int64)
    ordv = torch.where(ti < 0, -(ti & 0x7FFFFFFF), ti)                       # monotone integer image of the float
    steps = torch.arange(1, 4097, device=dev, dtype=torch.int64).view(1, -1)
    o = ordv.view(-1, 1) - steps
    bits = torch.where(o < 0, (-o) | 0x80000000, o)
    x = torch.where(bits >= (1 << 31), bits - (1 << 32), bits).to(torch.int32).view(torch.float32).contiguous()
    assert bool((x < Tb.view(-1, 1)).all())
    y = torch.empty_like(x)
    assert lib.rtm3d_sigmoid_f32(x.data_ptr(), y.data_ptr(), x.numel(), st) == 0
    torch.cuda.synchronize()
    assert bool((y < eb.view(-1, 1)).all()), "a logit below the bound reaches the bin's score edge"
    # usefulness: the bound is within 2% (+0.01) of the edge's true logit for every bin up to 0.999
    e64 = eb.double()
    true_logit = torch.log(e64 / (1 - e64))
    ok = (e64 > 0.999) | ((true_logit - Tb.double()) <= 0.02 * true_logit.abs() + 0.01)
    assert bool(ok.all())
