"""Store-only bandwidth of the device for buffers of the size of one SEAN operand (67 MB at B=64) -- the ceiling of
the pure-store kernels (actv, K-DYN): torch fill_ / copy_ as neutral references."""
import torch, sys
dev = torch.device("cuda:0")
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for mb in (16.8, 67.1, 268.4, 1073.7):
    n = int(mb * 1e6 / 2)
    x = torch.empty(n, device=dev, dtype=torch.bfloat16)
    y = torch.empty(n, device=dev, dtype=torch.bfloat16)
    us_fill = t(lambda: x.fill_(1.0))
    us_copy = t(lambda: y.copy_(x))
    print("%.1f MB: fill_ %.1f us (%.0f GB/s written)   copy_ %.1f us (%.0f GB/s read+written)" % (
        mb, us_fill, mb / us_fill * 1e3, us_copy, 2 * mb / us_copy * 1e3))
