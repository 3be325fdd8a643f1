#!/usr/bin/env python
"""Read-only HBM streaming ceiling on this GPU (context for roofline.frac): times torch's
sum-reduction over a 16 GiB fp32 tensor and a device-to-device copy (read+write bytes)."""
import torch

x = torch.empty(4 << 30, dtype=torch.float32, device="cuda").normal_()
y = torch.empty_like(x[: 2 << 30])
for name, fn, nbytes in (("sum (read-only)", lambda: x.sum(), x.numel() * 4),
                         ("copy (read+write)", lambda: y.copy_(x[: 2 << 30]), 2 * y.numel() * 4)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name:20s} {nbytes / best / 1e6:8.1f} GB/s  ({best:.3f} ms)")
