"""Practical HBM ceiling at the tensor sizes of the cfg2 step: torch device-to-device copy_ (read + write bytes),
CUDA events over 20 back-to-back copies.  MEASURED_PEAKS.json's 6.5 TB/s is the same measurement on 2 GiB tensors."""
import torch

for mb in (46, 92, 184, 369, 1024, 2048):
    n = mb * 1024 * 1024 // 2
    a = torch.randn(n, device="cuda").bfloat16() if n < 2**30 else torch.zeros(n, device="cuda", dtype=torch.bfloat16)
    b = torch.empty_like(a)
    for _ in range(3):
        b.copy_(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        b.copy_(a)
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 20
    print(f"copy {mb:5d} MB -> {mb:5d} MB: {us:8.1f} us  {2 * mb * 1.048576 / us * 1e3:8.1f} GB/s")
