"""H2D alone, D2H alone and both at once on two streams (pinned host memory): is the host link full duplex here?"""
import time, torch
dev = torch.device("cuda", 0)
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device=dev); d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    return dt
for _ in range(2): run(True, True)
a, b, c = run(True, False), run(False, True), run(True, True)
print("H2D alone %.1f GB/s, D2H alone %.1f GB/s, both at once: %.2f ms for 2 x 256 MiB = %.1f GB/s aggregate (serial would be %.2f ms)" % (
    n / a / 1e9, n / b / 1e9, c * 1e3, 2 * n / c / 1e9, (a + b) * 1e3))
