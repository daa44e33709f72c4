"""Write-only and read-only HBM stream rates of this GPU (torch fill_ / sum over 4 GiB), the ceilings of the training kernels'
write-only (forward, dgrad) and read-mostly (wgrad) streams."""
import torch
dev = torch.device("cuda:0")
x = torch.empty(1 << 30, dtype=torch.float32, device=dev)      # 4 GiB


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n


ms_w = t(lambda: x.fill_(1.0))
ms_r = t(lambda: x.sum())
y = torch.empty_like(x[: 1 << 29])
ms_c = t(lambda: y.copy_(x[: 1 << 29]))
gb = x.numel() * 4 / 1e9
print("write-only (fill_ 4 GiB): %.2f TB/s   read-only (sum 4 GiB): %.2f TB/s   copy 2+2 GiB: %.2f TB/s" % (gb / ms_w, gb / ms_r, gb / ms_c))
