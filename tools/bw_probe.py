#!/usr/bin/env python
"""HBM bandwidth probe with stock torch kernels: write-only (fill), read-only (sum), copy (read+write), each on buffers
far larger than L2. Gives the denominators for write-heavy vs read-heavy epilogue-bound convs."""
import torch

def t(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3

N = 1 << 30
x = torch.empty(N, dtype=torch.uint8, device="cuda").view(torch.float32)
y = torch.empty(N, dtype=torch.uint8, device="cuda").view(torch.float32)
x.fill_(1.0); y.fill_(2.0)
print(f"fill  (write only) {N / t(lambda: x.fill_(3.0)) / 1e9:8.0f} GB/s")
print(f"sum   (read only)  {N / t(lambda: x.sum()) / 1e9:8.0f} GB/s")
print(f"copy  (1R + 1W)    {2 * N / t(lambda: y.copy_(x)) / 1e9:8.0f} GB/s")
z = torch.empty_like(x)
print(f"add   (2R + 1W)    {3 * N / t(lambda: torch.add(x, y, out=z)) / 1e9:8.0f} GB/s")
xb = x.view(torch.bfloat16); 
print(f"relu bf16 (1R+1W)  {2 * N / t(lambda: torch.relu(xb, )) / 1e9:8.0f} GB/s (allocating)")
