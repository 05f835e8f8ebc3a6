"""Chain-parallel sharding over the GPUs of one box: rank g owns a contiguous
block of chains, every rank holds the (small) data set, and NO collective sits on the hot
path -- the only traffic is the optional gather of O(#params) results per chain and the
max-over-ranks of the device timings.  One process per GPU via ``torch.distributed``."""
import os

import torch

__all__ = ["world", "shard_bounds", "shard", "gather_chains", "max_over_ranks", "cadence_bounds", "allreduce_sum_",
           "device_local_cpus", "bind_to_device_numa"]


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process if absent)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_bounds(n_chains, world_size, rank):
    """Contiguous, balanced block [lo, hi) of ``n_chains`` for ``rank`` (sizes differ by <= 1)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(int(n_chains), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(x, world_size, rank, dim=0):
    """This rank's block of a per-chain array."""
    lo, hi = shard_bounds(x.shape[dim], world_size, rank)
    return x.narrow(dim, lo, hi - lo) if isinstance(x, torch.Tensor) else x[(slice(None),) * dim + (slice(lo, hi),)]


def gather_chains(local, n_chains, group=None):
    """All ranks' per-chain results back in chain order (uneven blocks allowed).  Off the hot
    path: ~(1 + n_params) doubles per chain."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local
    ws, rk = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(n_chains, ws, r)[1] - shard_bounds(n_chains, ws, r)[0] for r in range(ws)]
    cap = max(sizes)
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def max_over_ranks(value, device=None, group=None):
    """max over ranks of a python float (device timings are reported as the slowest rank's)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t.item()


def cadence_bounds(n_cadences, world_size, rank, align=32):
    """Cadence-parallel sharding ("few chains, huge N"): rank g owns the contiguous block
    [lo, hi) of the time series, block edges on multiples of ``align`` cadences.  log-likelihood and
    gradients are sums over cadences, so every rank evaluates ALL chains on its block and one
    all-reduce of (1 + n_params) doubles per chain (``allreduce_sum_``) completes the step."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    units = -(-int(n_cadences) // align)
    lo, hi = shard_bounds(units, world_size, rank)
    return min(lo * align, n_cadences), min(hi * align, n_cadences)


def allreduce_sum_(tensors, group=None):
    """In-place SUM over ranks of a list of tensors with ONE collective: packed into a single flat
    buffer (latency-bound exchange: bytes are irrelevant on NVLink, the launch count is not).
    No-op outside a process group."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tensors
    flat = torch.cat([x.reshape(-1) for x in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    o = 0
    for x in tensors:
        n = x.numel()
        x.copy_(flat[o:o + n].view_as(x))
        o += n
    return tensors


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def device_local_cpus(device_index):
    """CPUs on the NUMA node the GPU's PCIe root hangs off (sysfs ``local_cpulist``), or None."""
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as fh:
            cpus = _parse_cpulist(fh.read())
        return cpus or None
    except Exception:
        return None


def bind_to_device_numa(device_index):
    """Pin this process (one per GPU) to the CPUs local to its GPU so that pinned staging buffers
    allocated afterwards (first touch) and the submitting thread sit on the GPU's side of the
    socket interconnect: on a 2-socket box a far-side process pays ~2x on every host<->device copy.
    Returns (cpus_bound_or_None, previous_affinity)."""
    prev = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    cpus = device_local_cpus(device_index)
    if prev is None or not cpus:
        return None, prev
    use = cpus & prev
    if not use:
        return None, prev
    try:
        os.sched_setaffinity(0, use)
    except OSError:
        return None, prev
    return use, prev
