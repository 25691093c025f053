"""Single-node multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL on GPUs, gloo in CPU tests).

The reference only initialises a process group and exchanges files (vsc/baseline/inference.py:107-158).  Here:
  * stage A / C: videos and candidate pairs are independent units -> contiguous shards per rank, results gathered.
  * stage B: query rows are sharded, references replicated.  The global top-K is defined over ALL queries
    (vsc/index.py:142-165), so the ranks must agree on FAISS's radius: every tightening gathers the ranks' held
    scores (<= 2K floats in total) and takes the (min_results+1)-th best -- the only data-path exchange.
Everything here is host logic over tensors; it runs unchanged on CPU tensors with the gloo backend.
"""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n units for `rank`; sizes differ by at most one, order is preserved."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(items: Sequence, group=None) -> Sequence:
    rank, ws = world(group)
    lo, hi = shard_bounds(len(items), rank, ws)
    return items[lo:hi]


def gather_lists(local: List, group=None) -> List:
    """Concatenate per-rank Python lists in rank order on every rank (small results: boxes, matches, candidates)."""
    rank, ws = world(group)
    if ws == 1:
        return list(local)
    parts: List[Optional[List]] = [None] * ws
    dist.all_gather_object(parts, list(local), group=group)
    return [x for part in parts for x in part]


def all_gather_variable(t: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather 1-D tensors of different lengths (concatenated in rank order)."""
    rank, ws = world(group)
    if ws == 1:
        return t
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    padded = torch.zeros((cap,), dtype=t.dtype, device=t.device)
    padded[:t.numel()] = t
    bufs = [torch.empty_like(padded) for _ in range(ws)]
    dist.all_gather(bufs, padded, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)])


def global_count(local_count: int, device, group=None) -> int:
    rank, ws = world(group)
    if ws == 1:
        return int(local_count)
    c = torch.tensor([int(local_count)], dtype=torch.int64, device=device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    return int(c.item())


def agree_radius(local_scores: torch.Tensor, k: int, keep_max: bool, group=None) -> float:
    """The k-th best score over the union of every rank's held scores (k-th largest for inner product, k-th
    smallest for L2) -- FAISS's new radius when the global total exceeds max_results (k = min_results + 1)."""
    every = all_gather_variable(local_scores, group)
    if keep_max:
        return float(torch.topk(every, k, largest=True, sorted=True).values[-1])
    return float(torch.topk(every, k, largest=False, sorted=True).values[-1])


def max_over_ranks(value: float, device, group=None) -> float:
    rank, ws = world(group)
    if ws == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
