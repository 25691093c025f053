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
        torch = _lib.require_cuda()
        feats = np.concatenate([q.feature for q in queries])
        row, col, score = self.index.global_topk_device(feats, global_k, group=group)
        if score.numel() == 0:
            return []
        dev = score.device
        q_len = torch.tensor([len(q) for q in queries], dtype=torch.int64, device=dev)
        r_len = torch.from_numpy(self._ref_lengths).to(dev)
        q_vid = torch.repeat_interleave(torch.arange(len(queries), device=dev), q_len)[row]
        r_vid = torch.repeat_interleave(torch.arange(len(self._ref_ids), device=dev), r_len)[col]
        key = q_vid * len(self._ref_ids) + r_vid
        # hits are sorted best-first: a pair's first appearance carries its maximum, and first-appearance order
        # is exactly the reference's dict order followed by its stable sort
        uniq, inverse = torch.unique(key, return_inverse=True)
        first = torch.full((uniq.numel(),), key.numel(), dtype=torch.int64, device=dev)
        first.scatter_reduce_(0, inverse, torch.arange(key.numel(), device=dev), reduce="amin")
        first = torch.sort(first).values
        if limit is not None:
            first = first[:limit]
        qv, rv, sc = q_vid[first].cpu().numpy(), r_vid[first].cpu().numpy(), score[first].cpu().numpy()
        return [CandidatePair(query_id=queries[a].video_id, ref_id=self._ref_ids[b], score=s)
                for a, b, s in zip(qv, rv, sc)]
