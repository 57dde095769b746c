"""One-process-per-GPU plumbing: how submaps, tuples and database rows shard over the ranks of a
torch.distributed job, and the two exchange steps the path really has — the gradient all-reduce of the training step
(inside lpdnet_b200.optim.Adam.step) and the top-k merge of the database-sharded retrieval.  Eval embedding shards the
batch with no collective at all.

The functions take an optional `merge` callable so the host logic can be exercised on CPU with the gloo backend
(tests/test_parallel_cpu.py passes the numpy oracle's merge); on a GPU the default is the lpd_topk_merge kernel.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops

__all__ = ["world", "shard_range", "shard_tuples", "sharded_retrieval_topk", "allreduce_counters", "merge_nan"]


def world(group=None):
    """(rank, world_size) of the default / given process group; (0, 1) when torch.distributed is not initialised"""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n: int, world_size: int, rank: int):
    """contiguous [lo, hi) shard of n units: the first n % world_size ranks get one extra (never an empty middle rank)"""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_tuples(batch, world_size: int, rank: int):
    """(queries, positives, negatives, other_neg) with a leading tuple axis -> this rank's WHOLE tuples (a tuple is never
    split: the loss couples its members, train_pointnetvlad.py:202-217)"""
    lo, hi = shard_range(batch[0].shape[0], world_size, rank)
    return tuple(t[lo:hi] for t in batch)


def sharded_retrieval_topk(db_shard: torch.Tensor, queries: torch.Tensor, k: int, row_offset: int, group=None,
                           local_topk=None, merge=None):
    """Exact k nearest database rows of every query when the database rows are sharded over the ranks.
    db_shard [n_local, D] = global rows [row_offset, row_offset + n_local); queries [Nq, D] replicated.
    Each rank: local top-k with GLOBAL indices -> all-gather of (dist, idx) [Nq, k] -> deterministic merge
    (ascending distance, ties to the lower global index).  Returns (idx int32 [Nq, k], dist float64 [Nq, k]) on every rank."""
    local_topk = local_topk or (lambda d, q, kk, off: ops.retrieval_search(d, q, kk, idx_offset=off))
    merge = merge or ops.topk_merge
    rank, ws = world(group)
    Nq = queries.shape[0]
    if db_shard.shape[0] > 0:
        idx, dst = local_topk(db_shard, queries, min(k, db_shard.shape[0]), row_offset)
        if idx.shape[1] < k:                                  # shard smaller than k: pad with empty slots
            pad = k - idx.shape[1]
            idx = torch.cat((idx, torch.full((Nq, pad), -1, dtype=idx.dtype, device=idx.device)), 1)
            dst = torch.cat((dst, torch.full((Nq, pad), float("inf"), dtype=dst.dtype, device=dst.device)), 1)
    else:
        idx = torch.full((Nq, k), -1, dtype=torch.int32, device=queries.device)
        dst = torch.full((Nq, k), float("inf"), dtype=torch.float64, device=queries.device)
    if ws == 1:
        return merge(dst.unsqueeze(0), idx.unsqueeze(0))
    all_idx = torch.empty((ws,) + tuple(idx.shape), dtype=idx.dtype, device=idx.device)
    all_dst = torch.empty((ws,) + tuple(dst.shape), dtype=dst.dtype, device=dst.device)
    dist.all_gather(list(all_idx.unbind(0)), idx.contiguous(), group=group)
    dist.all_gather(list(all_dst.unbind(0)), dst.contiguous(), group=group)
    return merge(all_dst, all_idx)


def allreduce_counters(t: torch.Tensor, group=None) -> torch.Tensor:
    """sum of per-rank recall / count / one-percent counters when (database run, query run) PAIRS are sharded (evaluate.py:59-93)"""
    _, ws = world(group)
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def merge_nan(t: torch.Tensor, group=None) -> torch.Tensor:
    """Every entry of `t` is set (not NaN) on at most one rank: returns the union on every rank (NaN where no rank set it).
    Used for the per-query top-1 similarities when the database runs are sharded."""
    _, ws = world(group)
    if ws == 1:
        return t
    valid = ~torch.isnan(t)
    both = torch.stack((torch.where(valid, t, torch.zeros_like(t)), valid.to(t.dtype)))
    dist.all_reduce(both, op=dist.ReduceOp.SUM, group=group)
    return torch.where(both[1] > 0, both[0], torch.full_like(t, float("nan")))
