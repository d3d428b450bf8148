"""Raw device->host bandwidth into pinned memory (context for the e2e number)."""
import time, torch
n = 2 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("D2H %d MiB in %.1f ms = %.1f GB/s" % (n >> 20, dt * 1e3, n / dt / 1e9))
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("H2D %d MiB in %.1f ms = %.1f GB/s" % (n >> 20, dt * 1e3, n / dt / 1e9))
