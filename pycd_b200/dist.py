"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU
box, gloo in the CPU tests).  The path shards without any data-path collective:

* Ewald array: contiguous row blocks, one all-gather at the end;
* KMC: contiguous trajectory blocks, RNG keyed by the global trajectory id;
* MSD: per-rank species-averaged SD arrays, all-gathered (kB..MB) so that every rank
  computes numpy-identical means / standard errors / slopes.
"""
import numpy as np


def block(rank, world, n):
    """[lo, hi) of item `rank` when n items are split into `world` contiguous blocks."""
    return (rank * n) // world, ((rank + 1) * n) // world


def is_initialized():
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized()
    except ImportError:
        return False


def rank_world():
    if is_initialized():
        import torch.distributed as dist
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allgather_rows(full, group=None):
    """In-place completion of a row-sharded (N, ...) tensor: rank g has filled rows
    block(g, world, N); afterwards every rank holds all rows.  One all_gather when the blocks
    are equal, per-block broadcasts otherwise."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = full.shape[0]
    if n % world == 0:
        lo, hi = block(rank, world, n)
        mine = full[lo:hi].clone()
        dist.all_gather_into_tensor(full.view(-1), mine.view(-1), group=group)
    else:
        for g in range(world):
            lo, hi = block(g, world, n)
            if hi > lo:
                dist.broadcast(full[lo:hi], src=dist.get_global_rank(group, g) if group else g, group=group)
    return full


def gather_trajectory_arrays(local, n_total, device=None, group=None):
    """Concatenates per-rank (n_local, ...) numpy arrays in rank order -> (n_total, ...) on
    every rank (trajectory blocks from `block`)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    local = np.ascontiguousarray(local)
    tail = local.shape[1:]
    out = torch.zeros((n_total,) + tail, dtype=torch.from_numpy(local).dtype, device=device)
    lo, hi = block(rank, world, n_total)
    assert hi - lo == local.shape[0], 'local block does not match the trajectory partition'
    out[lo:hi] = torch.from_numpy(local).to(out.device)
    dist.all_reduce(out, group=group)  # disjoint blocks: the sum is the concatenation
    return out.cpu().numpy()
