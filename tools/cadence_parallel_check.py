"""Cadence-parallel evaluation on N GPUs (torchrun): every rank holds one block of the light curve,
evaluates all chains on it, one NCCL all-reduce sums logp and the gradients; rank 0 compares with the
single-GPU evaluation of the whole series and prints the timings.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/cadence_parallel_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exoplanet_b200 import fused, parallel, synthetic  # noqa: E402

rank, world, local = parallel.world()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
parallel.bind_to_device_numa(local)
dist.init_process_group("nccl", device_id=dev)
N, C, P = 4000000, 4, 2
t = synthetic.cadence_times(N, dt=0.00139)
y = synthetic.noise(N)
p = synthetic.transit_params(C, P, ecc=0.3)


def d(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=dev)


par = fused.pack_params(d(p["period"]), d(p["t0"]), d(p["ecc"]), d(p["omega"]), d(p["b"]), d(p["ror"]), d(p["m_planet"]), d(p["K"]))
star = fused.pack_star(torch.ones(C, dtype=torch.float64, device=dev), torch.ones(C, dtype=torch.float64, device=dev), d(p["u1"]), d(p["u2"]))
mean = d(p["mean"])
lo, hi = parallel.cadence_bounds(N, world, rank)
block = fused.TransitModel(fused.TransitData(t[lo:hi], y[lo:hi], np.full(hi - lo, 5e-4), device=dev), cadence_group=True)
full = fused.TransitModel(fused.TransitData(t, y, np.full(N, 5e-4), device=dev))


def timed(model, n=20):
    for _ in range(3):
        out = model.logp_grad(par, star, mean)
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        out = model.logp_grad(par, star, mean)
    b.record()
    torch.cuda.synchronize()
    return out, parallel.max_over_ranks(a.elapsed_time(b) / n, device=dev)


o_s, ms_s = timed(block)
o_f, ms_f = timed(full)
err = max(float(((a - b).abs().max() / b.abs().max()).item()) for a, b in zip(o_s, o_f))
if rank == 0:
    print("cadence-parallel over %d GPUs: N=%d, %d chains x %d planets: %.3f ms/eval (single GPU, whole series: %.3f ms); "
          "max rel. difference of logp / gradients %.2e" % (world, N, C, P, ms_s, ms_f, err))
    assert err < 1e-9, err
dist.destroy_process_group()
