"""Multi-GPU plumbing of the projection path.

BED rows are independent queries (reference src/main.rs:7435: one
perform_query per row), and the C4 index fits one B200, so N GPUs run N
processes, each with an index replica and its own share of the rows: there is
no data-path collective. torch.distributed is used only for the barrier and the
max-over-ranks timing that bench.py reports (NCCL on GPUs, gloo in CPU tests).
"""
import os

import numpy as np


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_rows(n_rows, rank, world):
    """Contiguous, balanced [lo, hi) slice of a row batch for `rank` (strong scaling of one BED)."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_seed(seed, rank):
    """Weak scaling: every rank draws its own BED of the configured size."""
    return seed + 1000 * rank


def init(backend):
    import torch.distributed as dist

    if not dist.is_initialized():
        dist.init_process_group(backend)
    return dist


def max_over_ranks(values, device="cpu"):
    """Element-wise max of a small float vector over all ranks (timing reduction)."""
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def gather_row_counts(local_rows_done, device="cpu"):
    """Sum of the rows every rank processed (the numerator of the aggregate ranges/s)."""
    import torch
    import torch.distributed as dist

    t = torch.tensor([float(local_rows_done)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t[0])
