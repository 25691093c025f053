"""`vsc.candidates` mirror: video-pair candidates from frame-level search (vsc/candidates.py:14-40).

`CandidateGeneration.query` keeps the reference signature and result (pairs sorted by aggregated score, best first,
ties in first-appearance order).  With the stock MaxScoreAggregation the per-pair maximum is taken on the GPU from
the sorted hit list, so no per-frame Python objects are built; any other ScoreAggregation goes through
VideoIndex.search exactly like the reference.
"""
from abc import ABC, abstractmethod
from typing import List, Optional

import numpy as np

from . import _lib
from .index import PairMatches, VideoFeature, VideoIndex
from .metrics import CandidatePair


class ScoreAggregation(ABC):
    @abstractmethod
    def aggregate(self, match: PairMatches) -> float:
        pass

    def score(self, match: PairMatches) -> CandidatePair:
        return CandidatePair(query_id=match.query_id, ref_id=match.ref_id, score=self.aggregate(match))


class MaxScoreAggregation(ScoreAggregation):
    def aggregate(self, match: PairMatches) -> float:
        return np.max([m.score for m in match.matches])


class CandidateGeneration:
    def __init__(self, references: List[VideoFeature], aggregation: ScoreAggregation):
        self.aggregation = aggregation
        self.index = VideoIndex(references[0].dimensions())
        self.index.add(references)
        self._ref_ids = [r.video_id for r in references]
        self._ref_lengths = np.array([len(r) for r in references], dtype=np.int64)

    def query(self, queries: List[VideoFeature], global_k: int, limit: Optional[int] = None, group=None) -> List[CandidatePair]:
        """`limit` (extension): keep only the best `limit` pairs -- what every caller in the reference does next.
        `group` (extension): query-sharded search over the ranks of a torch.distributed process group."""
        if type(self.aggregation) is MaxScoreAggregation and global_k >= 0:
            return self._query_max_fused(queries, global_k, limit, group)
        matches = self.index.search(queries, global_k=global_k)
        candidates = sorted((self.aggregation.score(m) for m in matches), key=lambda c: c.score, reverse=True)
        return candidates if limit is None else candidates[:limit]

    def _query_max_fused(self, queries, global_k, limit, group=None):
        """Global top-K frame pairs, then the per-video-pair maximum, without leaving the device: csrc/search_ops.cu
        vsc_pair_max (hash of video pairs -> first appearance in the best-first hit list -> ordered compaction)."""
        import ctypes
        torch = _lib.require_cuda()
        row, col, score = self.index.global_topk_device(queries, global_k, group=group)
        n = score.numel()
        if n == 0:
            return []
        dev = score.device
        q_vid = torch.from_numpy(np.repeat(np.arange(len(queries), dtype=np.int32), [len(q) for q in queries])).to(dev)
        r_vid = torch.from_numpy(np.repeat(np.arange(len(self._ref_ids), dtype=np.int32), self._ref_lengths)).to(dev)
        lib = _lib.load()
        cap = n if limit is None else min(n, int(limit))
        out_q = torch.empty((cap,), dtype=torch.int32, device=dev)
        out_r = torch.empty((cap,), dtype=torch.int32, device=dev)
        out_s = torch.empty((cap,), dtype=torch.float32, device=dev)
        n_unique = torch.zeros((1,), dtype=torch.int64, device=dev)
        scratch = torch.empty((int(lib.vsc_pair_max_scratch_bytes(n)),), dtype=torch.uint8, device=dev)
        row, col, score = row.contiguous(), col.contiguous(), score.contiguous()
        with torch.cuda.device(dev):
            rc = lib.vsc_pair_max(row.data_ptr(), col.data_ptr(), score.data_ptr(), n, q_vid.data_ptr(), r_vid.data_ptr(),
                                  len(self._ref_ids), cap, out_q.data_ptr(), out_r.data_ptr(), out_s.data_ptr(),
                                  n_unique.data_ptr(), scratch.data_ptr(),
                                  ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(rc, "vsc_pair_max")
        k = min(cap, int(n_unique.item()))
        qv, rv, sc = out_q[:k].cpu().numpy(), out_r[:k].cpu().numpy(), out_s[:k].cpu().numpy()
        return [CandidatePair(query_id=queries[a].video_id, ref_id=self._ref_ids[b], score=s)
                for a, b, s in zip(qv, rv, sc)]
