"""Pure-write / pure-read / copy bandwidth of the box (torch kernels, CUDA events): context for the HBM-bound kernels."""
import torch
n = 1 << 29   # 2 GiB of fp32
a = torch.empty(n, device="cuda"); b = torch.empty(n, device="cuda")
def t(f, reps=10):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
gb = n * 4 / 1e9
print("fill  (write only) %.0f GB/s" % (gb / t(lambda: a.fill_(1.0)) * 1e3))
print("sum   (read only)  %.0f GB/s" % (gb / t(lambda: a.sum()) * 1e3))
print("copy  (read+write) %.0f GB/s" % (2 * gb / t(lambda: b.copy_(a)) * 1e3))
print("mul_  (read+write in place) %.0f GB/s" % (2 * gb / t(lambda: a.mul_(1.0001)) * 1e3))
