"""Ensemble partitioning for multi-GPU runs.

A single simulation does not shard: every step needs the global X projection before the next
wake kick.  What scales is an ensemble of independent simulations (a bunch-current or parameter
scan, BASELINE.json configs[3]): member r runs on rank floor(r * world / members) -- contiguous
blocks, one process per GPU, NO collective on the step path.  The only communication is the
gather of the per-member summaries (a few floats per member and out-step) at the end, over
torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def member_block(members, world, rank):
    """[lo, hi) of the members rank `rank` owns: contiguous blocks whose sizes differ by at most 1"""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(members, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def scan_currents(i_min, i_max, members):
    """bunch currents of a linear current scan (configs[3]: 256 members)"""
    return np.linspace(i_min, i_max, members, dtype=np.float64)


def gather_summaries(local, members, dist=None):
    """Collect per-member summary rows from all ranks on every rank.

    `local` is the (hi - lo, k) array of this rank's block; returns the (members, k) array in
    member order.  `dist` is torch.distributed (initialised) or None for a single process.
    """
    local = np.ascontiguousarray(local, np.float64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        assert local.shape[0] == members
        return local
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    k = local.shape[1]
    lo, hi = member_block(members, world, rank)
    assert local.shape[0] == hi - lo, (local.shape, lo, hi)
    width = -(-members // world)  # blocks are padded to the largest one for all_gather
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    buf = torch.zeros((width, k), dtype=torch.float64, device=dev)
    buf[: hi - lo] = torch.from_numpy(local).to(dev)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = np.zeros((members, k))
    for r, p in enumerate(parts):
        a, b = member_block(members, world, r)
        out[a:b] = p[: b - a].cpu().numpy()
    return out
