"""Calibration: what does a plain streaming read / copy reach on this GPU (torch kernels)?"""
import torch, time
dev = torch.device("cuda", 0)
n = 8 << 30
x = torch.empty(n // 8, dtype=torch.int64, device=dev).random_()
y = torch.empty_like(x)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: x.sum()); print(f"sum (read only)      {n / ms / 1e6:7.0f} GB/s read")
ms = t(lambda: torch.max(x)); print(f"max (read only)      {n / ms / 1e6:7.0f} GB/s read")
ms = t(lambda: y.copy_(x)); print(f"copy                 {n / ms / 1e6:7.0f} GB/s read + same written = {2 * n / ms / 1e6:.0f} GB/s total")
ms = t(lambda: y.fill_(1)); print(f"fill (write only)    {n / ms / 1e6:7.0f} GB/s written")
