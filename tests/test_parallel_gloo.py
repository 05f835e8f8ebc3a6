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
