"""CPU, world_size 2, gloo: the host-side logic of the chain-sharded multi-GPU path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from exoplanet_b200 import parallel


def test_shard_bounds_partition():
    for n in (0, 1, 7, 128, 1024, 1025):
        for ws in (1, 2, 3, 8):
            blocks = [parallel.shard_bounds(n, ws, r) for r in range(ws)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks[:-1], blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 2, 2)
    x = torch.arange(10)
    assert torch.equal(torch.cat([parallel.shard(x, 3, r) for r in range(3)]), x)


def test_cpulist_parsing_and_binding_is_harmless_without_a_gpu():
    assert parallel._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert parallel._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    cpus, prev = parallel.bind_to_device_numa(0)  # no CUDA device here: nothing to bind to, affinity untouched
    assert prev == before and os.sched_getaffinity(0) == before
    assert cpus is None or cpus <= before


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, n_chains, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        full = torch.arange(n_chains * 3, dtype=torch.float64).reshape(n_chains, 3)
        mine = parallel.shard(full, ws, rank)
        # stand-in for the per-chain evaluation each rank does on its own GPU
        local = mine * 2.0 + 1.0
        gathered = parallel.gather_chains(local, n_chains)
        ok = torch.equal(gathered, full * 2.0 + 1.0)
        slow = parallel.max_over_ranks(10.0 + rank)
        r, w, lr = parallel.world()
        out[rank] = (bool(ok), slow, r, w)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_chains", [7, 128])
def test_two_ranks_gather_and_max(n_chains):
    ws = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(ws, _free_port(), n_chains, out), nprocs=ws, join=True)
    assert len(out) == ws
    for rank in range(ws):
        ok, slow, r, w = out[rank]
        assert ok and slow == 11.0 and r == rank and w == ws


def test_cadence_bounds_partition():
    for n in (1, 31, 32, 33, 100000, 1000003):
        for ws in (1, 2, 3, 8):
            blocks = [parallel.cadence_bounds(n, ws, r) for r in range(ws)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks[:-1], blocks[1:]))
            assert all(lo % 32 == 0 for lo, hi in blocks if lo < n)


def _cadence_worker(rank, ws, port, out):
    """cadence-parallel step: every rank evaluates all chains on its block of the light curve (the
    oracle stands in for the GPU kernel here), one packed all-reduce sums logp and the gradients"""
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import xo_oracle as O
    from exoplanet_b200 import synthetic

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        N, C = 5000, 3
        t = synthetic.cadence_times(N)
        y = synthetic.noise(N)
        ivar = np.full(N, 1 / 5e-4 ** 2)
        p = synthetic.transit_params(C, 2, ecc=0.3)
        el = O.elements_from_params(p["period"], p["t0"], p["ecc"], p["omega"], p["b"], p["ror"])
        cl = O.get_cl(p["u1"], p["u2"])
        one, zero = np.ones(1), np.zeros(1)
        lo, hi = parallel.cadence_bounds(N, ws, rank)
        part = O.transit_logp_grad(t[lo:hi], y[lo:hi], ivar[lo:hi], el, cl, p["mean"], zero, one)
        tens = [torch.as_tensor(np.ascontiguousarray(x)) for x in part]
        parallel.allreduce_sum_(tens)
        full = O.transit_logp_grad(t, y, ivar, el, cl, p["mean"], zero, one)
        err = max(float(np.max(np.abs(a.numpy() - b) / (1e-300 + np.max(np.abs(b))))) for a, b in zip(tens, full))
        out[rank] = err
    finally:
        dist.destroy_process_group()


def test_two_ranks_cadence_parallel_sum():
    ws = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_cadence_worker, args=(ws, _free_port(), out), nprocs=ws, join=True)
    assert len(out) == ws and all(out[r] < 1e-12 for r in range(ws)), dict(out)
