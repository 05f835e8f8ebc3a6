"""How fast does a 2.4 MB pinned block reach the device: one DMA, or split over 2 / 3 / 4 streams?  (gpurun only)"""
import time

import torch

N = 300_000
src = torch.empty(N, dtype=torch.float64).pin_memory()
src.uniform_()
dst = torch.empty(N, dtype=torch.float64, device="cuda")
streams = [torch.cuda.Stream() for _ in range(4)]
for parts in (1, 2, 3, 4, 8):
    best = 1e9
    for rep in range(30):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step = (N + parts - 1) // parts
        for k in range(parts):
            with torch.cuda.stream(streams[k % 4]):
                dst[k * step:(k + 1) * step].copy_(src[k * step:(k + 1) * step], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print("parts=%d  best %.1f us  %.1f GB/s" % (parts, best * 1e6, N * 8 / best / 1e9), flush=True)
