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


def agree_radius(local_scores: torch.Tensor, k: int, keep_max: bool, group=None, fallback: float = None) -> float:
    """The k-th best score over the union of every rank's held scores (k-th largest for inner product, k-th
    smallest for L2) -- FAISS's new radius when the global total exceeds max_results (k = min_results + 1).

    On CUDA tensors: three radix-selection passes whose 2048-bin histograms are all-reduced (8 KB per pass over
    NVLink; csrc/select.cu vsc_select_hist / vsc_select_pick) -- no score leaves its GPU.  CPU tensors (the gloo
    tests of the host logic) gather and sort.  Fewer than k scores in total (possible only after an in-batch prune,
    which drops the k-th best itself): the answer is the best prune threshold of any rank, passed as `fallback`."""
    rank, ws = world(group)
    n_all = global_count(local_scores.numel(), local_scores.device, group)
    if n_all < k:
        if fallback is None:
            raise ValueError(f"agree_radius: only {n_all} scores for k = {k}")
        t = torch.tensor([fallback], dtype=torch.float64, device=local_scores.device)
        if ws > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX if keep_max else dist.ReduceOp.MIN, group=group)
        return float(t.item())
    if local_scores.is_cuda:
        import ctypes
        from . import _lib
        lib = _lib.load()
        scores = local_scores.contiguous()
        state = torch.empty((2048 + 8,), dtype=torch.int32, device=scores.device)
        out = torch.empty((1,), dtype=torch.float32, device=scores.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(scores.device).cuda_stream)
        with torch.cuda.device(scores.device):
            for p in range(3):
                _lib.check(lib.vsc_select_hist(scores.data_ptr(), scores.numel(), int(k), 1 if keep_max else 0, p,
                                               state.data_ptr(), stream), "vsc_select_hist")
                if ws > 1:
                    dist.all_reduce(state[:2048], op=dist.ReduceOp.SUM, group=group)
                _lib.check(lib.vsc_select_pick(1 if keep_max else 0, p, state.data_ptr(), out.data_ptr(), stream),
                           "vsc_select_pick")
        return float(out.item())
    every = all_gather_variable(local_scores, group)
    return float(torch.topk(every, k, largest=keep_max, sorted=True).values[-1])


def max_over_ranks(value: float, device, group=None) -> float:
    rank, ws = world(group)
    if ws == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def all_gather_rows(t: torch.Tensor, group=None) -> List[torch.Tensor]:
    """All-gather 2-D tensors [n_r, d] with different n_r; returns the per-rank tensors in rank order."""
    rank, ws = world(group)
    if ws == 1:
        return [t]
    d = t.shape[1]
    flat = all_gather_variable(t.reshape(-1).contiguous(), group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(counts, n, group=group)
    out, at = [], 0
    for c in counts:
        rows = int(c.item())
        out.append(flat[at:at + rows * d].reshape(rows, d))
        at += rows * d
    return out


def all_gather_video_features(local, n_videos: int, device=None, group=None, on_device: bool = False):
    """Stage A runs video i on rank i % world_size (inference_impl.select_videos, the reference's VideoDataset rule);
    the search stage wants every rank to hold ALL reference descriptors (BASELINE.json configs[4]: "ref descriptors
    NCCL all-gathered over NVLink").  `local`: this rank's VideoFeatures in its own order.  Returns the VideoFeatures of
    all n_videos videos in global video order on every rank: ids and timestamps travel as small objects, the
    descriptor rows in one all-gather of device tensors.  `local` features may be CUDA tensors (inference with
    on_device=True); on_device=True leaves the gathered descriptors on the GPU as row views of the gathered matrices."""
    import numpy as np
    from .index import VideoFeature
    rank, ws = world(group)
    if ws == 1:
        return list(local)
    dim = local[0].feature.shape[1] if local else 0
    dims = [None] * ws
    dist.all_gather_object(dims, dim, group=group)
    dim = max(dims)
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    if local and isinstance(local[0].feature, torch.Tensor):
        dtype = np.float32
        rows_t = torch.cat([v.feature for v in local]).to(dev, torch.float32)
    else:
        dtype = np.result_type(*[v.feature.dtype for v in local]) if local else np.float32
        rows = np.concatenate([v.feature for v in local]) if local else np.zeros((0, dim), dtype)
        rows_t = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float32)).to(dev)
    parts = all_gather_rows(rows_t.contiguous(), group)
    meta = [None] * ws
    dist.all_gather_object(meta, [(v.video_id, v.timestamps, len(v)) for v in local], group=group)
    out = [None] * n_videos
    for r in range(ws):
        feats, at = parts[r] if on_device else parts[r].cpu().numpy(), 0
        for slot, (vid, ts, n) in zip(range(r, n_videos, ws), meta[r]):
            rows_v = feats[at:at + n] if on_device else feats[at:at + n].astype(dtype, copy=False)
            out[slot] = VideoFeature(video_id=vid, timestamps=ts, feature=rows_v)
            at += n
    assert all(v is not None for v in out), "every video must be owned by exactly one rank"
    return out
