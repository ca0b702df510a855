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


def ensure_process_group(local_rank=0):
    """Initialises torch.distributed from the torchrun environment when WORLD_SIZE > 1 and nobody did
    yet: NCCL on the GPU box (one process per GPU), gloo in CPU tests (PYCD_DIST_BACKEND overrides)."""
    import os
    if int(os.environ.get('WORLD_SIZE', '1')) <= 1 or is_initialized():
        return
    import torch
    import torch.distributed as dist
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    backend = os.environ.get('PYCD_DIST_BACKEND') or ('nccl' if torch.cuda.is_available() else 'gloo')
    if backend == 'nccl':
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    else:
        dist.init_process_group(backend)


def barrier():
    if is_initialized():
        import torch.distributed as dist
        dist.barrier()


def _comm_device(device=None):
    import torch
    import torch.distributed as dist
    if device is not None:
        return device
    if dist.get_backend() == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def gather_trajectory_arrays(local, n_total, device=None, group=None):
    """Concatenates per-rank (n_local, ...) numpy arrays in rank order -> (n_total, ...) on
    every rank (trajectory blocks from `block`): one all_gather of blocks padded to the largest."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    local = np.ascontiguousarray(local)
    tail = local.shape[1:]
    lo, hi = block(rank, world, n_total)
    assert hi - lo == local.shape[0], 'local block does not match the trajectory partition'
    width = max(block(g, world, n_total)[1] - block(g, world, n_total)[0] for g in range(world))
    dev = _comm_device(device)
    mine = torch.zeros((width,) + tail, dtype=torch.from_numpy(local).dtype, device=dev)
    mine[:hi - lo] = torch.from_numpy(local).to(dev)
    out = torch.empty((world * width,) + tail, dtype=mine.dtype, device=dev)
    dist.all_gather_into_tensor(out.view(-1), mine.view(-1), group=group)
    out = out.view((world, width) + tail).cpu().numpy()
    return np.concatenate([out[g, :block(g, world, n_total)[1] - block(g, world, n_total)[0]]
                           for g in range(world)], axis=0)


def gather_blocks(local, n_total, local_rank=0):
    """gather_trajectory_arrays for drivers launched under torchrun (creates the process group on
    first use)."""
    ensure_process_group(local_rank)
    return gather_trajectory_arrays(local, n_total)
