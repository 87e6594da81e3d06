"""Pure-write / copy DRAM bandwidth probe (context for the walk kernel's 1.2 GB of mandatory writes)."""
import torch, time
x = torch.empty(1280 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ms = t(lambda: x.fill_(1.0)); print(f"fill 1.34GB: {ms:.4f} ms  {x.numel()*8/ms/1e9:.0f} GB/s write")
ms = t(lambda: y.copy_(x)); print(f"copy 1.34GB: {ms:.4f} ms  {2*x.numel()*8/ms/1e9:.0f} GB/s read+write")
ms = t(lambda: x.sum()); print(f"sum  1.34GB: {ms:.4f} ms  {x.numel()*8/ms/1e9:.0f} GB/s read")
