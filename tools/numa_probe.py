"""Host<->device copy time as a function of where the submitting thread and the pinned buffer live
(run on the GPU box): for each NUMA node, bind, allocate pinned memory, time 3 x 0.8 MB H2D + sync."""
import glob
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exoplanet_b200 import parallel  # noqa: E402

torch.cuda.init()
print("GPU0 local cpus (sysfs):", sorted(parallel.device_local_cpus(0) or [])[:4], "...", len(parallel.device_local_cpus(0) or []))
allc = os.sched_getaffinity(0)
N = 100000
dst = torch.empty(3 * N, dtype=torch.float64, device="cuda")


def timeit(bufs, n=200):
    for _ in range(10):
        for k, b in enumerate(bufs):
            dst[k * N:(k + 1) * N].copy_(b, non_blocking=True)
        torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        for k, b in enumerate(bufs):
            dst[k * N:(k + 1) * N].copy_(b, non_blocking=True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    return "min %.0f median %.0f p90 %.0f us" % (ts.min(), np.median(ts), np.percentile(ts, 90))


nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
bufs_by_node = {}
for nd in nodes:
    cpus = parallel._parse_cpulist(open(nd + "/cpulist").read()) & allc
    if not cpus:
        continue
    os.sched_setaffinity(0, cpus)
    bufs = [torch.from_numpy(np.random.default_rng(k).normal(size=N)).pin_memory() for k in range(3)]
    bufs_by_node[nd] = bufs
    print(os.path.basename(nd), "cpus", len(cpus), "buffers allocated + submitted from this node:", timeit(bufs))
for nd, bufs in bufs_by_node.items():
    for nd2 in nodes:
        cpus = parallel._parse_cpulist(open(nd2 + "/cpulist").read()) & allc
        if not cpus:
            continue
        os.sched_setaffinity(0, cpus)
        print("buffers of", os.path.basename(nd), "submitted from", os.path.basename(nd2), ":", timeit(bufs))
os.sched_setaffinity(0, allc)
# size sweep on the last buffers: fixed cost vs bandwidth
big = torch.empty(1 << 24, dtype=torch.float64).pin_memory()
dbig = torch.empty(1 << 24, dtype=torch.float64, device="cuda")
for nb in (1 << 10, 1 << 14, 1 << 17, 1 << 20, 1 << 24):
    ts = []
    for _ in range(50):
        t0 = time.perf_counter()
        dbig[:nb].copy_(big[:nb], non_blocking=True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print("H2D %9d B: median %.1f us  -> %.1f GB/s" % (nb * 8, np.median(ts) * 1e6, nb * 8 / np.median(ts) / 1e9))
