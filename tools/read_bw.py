"""Developer tool: read-only HBM bandwidth of this B200 through plain torch reductions (calibration for the roofline)."""
import torch
dev = torch.device("cuda:0")
for mb in (377, 1024, 4096):
    x = torch.randn(mb * 1024 * 1024 // 4, device=dev)
    y = torch.empty_like(x)
    for name, fn in (("sum", lambda: x.sum()), ("max", lambda: x.max()), ("copy", lambda: y.copy_(x))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        bytes_ = x.numel() * 4 * (2 if name == "copy" else 1)
        print(f"{mb} MB {name}: {ms*1e3:.1f} us  {bytes_/ms/1e6:.0f} GB/s")
