#!/usr/bin/env python
"""Pure read / pure write / copy bandwidth on this GPU (developer tool; roofline context for write-only kernels)."""
import torch
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
x = torch.empty(1 << 28, dtype=torch.float32, device="cuda")  # 1 GiB
y = torch.empty_like(x)
gb = x.numel() * 4 / 1e9
print(f"fill  (write only): {gb / t(lambda: x.fill_(1.0)):8.1f} GB/s")
print(f"sum   (read only) : {gb / t(lambda: x.sum()):8.1f} GB/s")
print(f"copy  (read+write): {2 * gb / t(lambda: y.copy_(x)):8.1f} GB/s")
