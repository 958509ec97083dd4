"""Multi-GPU plumbing of the projection path (one process per GPU).

Two ways to spread a query batch over N GPUs:

* rows over index replicas — BED rows are independent queries (reference
  src/main.rs:7435: one perform_query per row); when the index fits one B200
  every rank holds a replica and takes its own share of the rows: no data-path
  collective;
* index sharded by target sequence — rank r holds the entries,
  run stream and visited sets of the sequences it owns; a batch is a collective
  call and the library exchanges lifted hits (all-to-all-v) and the next
  frontier (all-gather-v) over NCCL between transitive hops (csrc/comm.cu).

torch.distributed carries only the plumbing: the NCCL unique id, the barrier,
the max-over-ranks timing and the gather of result columns to rank 0 (NCCL on
GPUs, gloo in CPU tests).
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


def nccl_comm(rank, world, device):
    """NCCL endpoint of this rank for the sharded index: rank 0 creates the unique id,
    torch.distributed (already initialised) hands it to the others."""
    import torch.distributed as dist

    import impg_b200 as ix

    box = [ix.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return ix.Comm.nccl(box[0], rank, world, device)


def gather_columns(cols):
    """Every rank's result columns on every rank (list indexed by rank); feed to
    impg_b200.merge_shard_columns."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [cols]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, cols)
    return out


def partition_sharded(shard, comm, pparams):
    """`impg partition -o bed` over a target-sharded index, one process per GPU: every rank steps through the same
    windows (the partitioner is deterministic), answers each with the collective masked walk on its shard, and the
    per-rank BED rows are all-gathered so that every rank feeds the same intervals. Every rank returns the partitions."""
    import impg_b200 as ix

    def answer(window, qp):
        part = shard.query_batch_bed_sharded(comm, window, qp).columns()
        return ix.merge_shard_columns(gather_columns(part))

    return ix.partition_with(shard, pparams, answer)


def owner_histogram(owner, world):
    """Sequences per rank of an owner map."""
    return np.bincount(np.asarray(owner, dtype=np.int64), minlength=world)
