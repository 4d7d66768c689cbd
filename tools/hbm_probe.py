"""Measures plain HBM write (fill), read (sum) and copy bandwidth with torch, for context next to the
env-rollout kernel's write-dominated stream (24 B written per 4 B read).  Not part of the product."""
import torch

def timed(fn, n=10):
    for _ in range(3):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        torch.cuda.synchronize(); s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best * 1e-3

n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
print("fill  %.0f GB/s" % (4 * n / timed(lambda: a.fill_(1.0)) / 1e9))
print("copy  %.0f GB/s (read+write)" % (8 * n / timed(lambda: b.copy_(a)) / 1e9))
print("sum   %.0f GB/s" % (4 * n / timed(lambda: a.sum()) / 1e9))
