"""H2D rate of the box from pinned memory, in the frame sizes bench.py uploads (context for the e2e number)."""
import time, torch
n = 3840 * 2160 * 2
src = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(8)]
dst = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(8)]
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    for i in range(40):
        dst[i % 8].copy_(src[i % 8], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("H2D %.1f GB/s (%.2f ms per 16.6 MB plane)" % (40 * n / dt / 1e9, 1000 * dt / 40))
