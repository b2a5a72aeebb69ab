"""Measures pinned H2D / D2H bandwidth on the box (context for the e2e number)."""
import torch, time
n = 78643200
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, a, b in (("H2D", d, h), ("D2H", h, d)):
    for _ in range(3): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
    print("%s %.1f MB in %.3f ms -> %.1f GB/s" % (name, n / 1e6, dt * 1e3, n / dt / 1e9))
