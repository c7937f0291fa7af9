"""Multi-GPU plumbing for the cube fit: one process per GPU, disjoint contiguous
pixel blocks, results gathered to rank 0 on the host.  No collective touches the
data path (pixels never interact while fitting, reference main.py:436-472); the
only `torch.distributed` calls are a barrier, the max-over-ranks of timings and
the final host gather -- they work identically on NCCL (GPU box) and gloo (CPU
tests)."""
import os

import numpy as np


def rank_info():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def block_bounds(n_items, world):
    """Contiguous [start, stop) of every rank; sizes differ by at most one."""
    base, extra = divmod(int(n_items), int(world))
    starts = [r * base + min(r, extra) for r in range(world + 1)]
    return [(starts[r], starts[r + 1]) for r in range(world)]


def my_block(n_items, rank, world):
    return block_bounds(n_items, world)[rank]


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for world 1)."""
    rank, local_rank, world = rank_info()
    if world == 1:
        return None
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return dist


def max_over_ranks(value, dist, device=None):
    """Max of a python float over ranks (the timing rule: slowest rank counts)."""
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_blocks(local, n_items, dist, device=None):
    """Gather per-rank result arrays (first axis = this rank's block) into the full
    array on every rank.  `local` is a float64/int64 numpy array."""
    if dist is None:
        return local
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    bounds = block_bounds(n_items, world)
    nmax = max(b - a for a, b in bounds)
    pad = np.zeros((nmax,) + local.shape[1:], dtype=local.dtype)
    pad[: local.shape[0]] = local
    t = torch.from_numpy(pad).to(device or "cpu")
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    parts = [o.cpu().numpy()[: b - a] for o, (a, b) in zip(outs, bounds)]
    return np.concatenate(parts, axis=0)
