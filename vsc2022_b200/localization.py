"""`vsc.baseline.localization` mirror (localization.py:16-96): align candidate pairs, emit Match rows.

Same classes and signatures as the reference.  `VCSLLocalization.localize_all` keeps everything on the device:
the frame descriptors of the videos involved are uploaded once, each pair's similarity matrix Q.R^T + bias is
written straight into the packed buffer the TN kernels read (pairs of equal shape are multiplied together as one
strided-batched fp32 GEMM -- a plain library GEMM, the fused tensor-core kernel for it is listed in DESIGN.md),
the TN pipeline aligns the whole batch in one call, and the MaxSim box score comes back from the same call.
Only the boxes (a few integers per pair) return to the host, where they are mapped to timestamps.
"""
import abc
from typing import Dict, List

import numpy as np

from . import _lib
from .index import VideoFeature
from .metrics import CandidatePair, Match


class Localization(abc.ABC):
    @abc.abstractmethod
    def localize(self, candidate: CandidatePair) -> List[Match]:
        pass

    def localize_all(self, candidates: List[CandidatePair]) -> List[Match]:
        matches = []
        for c in candidates:
            matches.extend(self.localize(c))
        return matches


class LocalizationWithMetadata(Localization):
    def __init__(self, queries: List[VideoFeature], refs: List[VideoFeature]):
        self.queries = {m.video_id: m for m in queries}
        self.refs = {m.video_id: m for m in refs}

    def similarity(self, candidate: CandidatePair):
        """Host-side similarity of one pair (reference API; the batched device path does not call it)."""
        return np.matmul(self.queries[candidate.query_id].feature, self.refs[candidate.ref_id].feature.T)


class _DeviceVideos:
    """Frame descriptors of a video collection, concatenated on the device, uploaded lazily per video."""

    def __init__(self, videos: Dict[object, VideoFeature], device):
        self.videos, self.device = videos, device
        self.slot: Dict[object, int] = {}
        self.start: List[int] = []
        self.length: List[int] = []
        self.chunks = []
        self.rows = 0
        self._cat = None

    def ensure(self, ids):
        torch = _lib.require_cuda()
        new = [i for i in dict.fromkeys(ids) if i not in self.slot]
        if new:
            host = np.concatenate([np.asarray(self.videos[i].feature, dtype=np.float32) for i in new])
            self.chunks.append(torch.from_numpy(host).to(self.device, non_blocking=True))
            for i in new:
                self.slot[i] = len(self.start)
                self.start.append(self.rows)
                self.length.append(len(self.videos[i]))
                self.rows += len(self.videos[i])
            self._cat = None

    def matrix(self):
        torch = _lib.require_cuda()
        if self._cat is None:
            self._cat = self.chunks[0] if len(self.chunks) == 1 else torch.cat(self.chunks)
            self.chunks = [self._cat]
        return self._cat


class VCSLLocalization(LocalizationWithMetadata):
    GROUP_ROWS = 1 << 22   # gathered descriptor rows per batched GEMM (bounds the temporary to ~8 GB at d=512)

    def __init__(self, queries, refs, model_type, similarity_bias=0.0, **kwargs):
        super().__init__(queries, refs)
        from .vta import build_vta_model
        self.model = build_vta_model(model_type, **kwargs)
        self.similarity_bias = similarity_bias
        self._dq = self._dr = None

    def similarity(self, candidate: CandidatePair):
        """Add an optional similarity bias (some aligners do not tolerate negative values well)."""
        return super().similarity(candidate) + self.similarity_bias

    # ---- device path ----------------------------------------------------------------------------------------
    def _pack_similarities(self, candidates):
        """All similarity matrices, computed on the device, in one packed float32 buffer (16-byte aligned pairs)."""
        torch = _lib.require_cuda()
        dev = self.model._device()
        if self._dq is None:
            self._dq, self._dr = _DeviceVideos(self.queries, dev), _DeviceVideos(self.refs, dev)
        self._dq.ensure([c.query_id for c in candidates])
        self._dr.ensure([c.ref_id for c in candidates])
        Q, R = self._dq.matrix(), self._dr.matrix()
        n = len(candidates)
        qs = np.array([self._dq.slot[c.query_id] for c in candidates])
        rs = np.array([self._dr.slot[c.ref_id] for c in candidates])
        q_start, q_len = np.array(self._dq.start)[qs], np.array(self._dq.length)[qs]
        r_start, r_len = np.array(self._dr.start)[rs], np.array(self._dr.length)[rs]
        sizes = q_len.astype(np.int64) * r_len
        off = np.zeros(n, dtype=np.int64)
        padded = (sizes + 3) & ~np.int64(3)
        if n:
            off[1:] = np.cumsum(padded[:-1])
        sims = torch.empty((int(padded.sum()) + 4,), dtype=torch.float32, device=dev)
        # equal-shaped pairs -> one strided-batched GEMM writing straight into the packed buffer
        order = np.lexsort((r_len, q_len))
        bounds = np.flatnonzero(np.diff(q_len[order]) | np.diff(r_len[order])) + 1
        for grp in np.split(order, bounds):
            lq, lr = int(q_len[grp[0]]), int(r_len[grp[0]])
            if lq == 0 or lr == 0:
                continue
            step = max(1, self.GROUP_ROWS // max(lq + lr, 1))
            for s in range(0, len(grp), step):
                g = grp[s:s + step]
                qi = torch.from_numpy(q_start[g][:, None] + np.arange(lq)[None, :]).to(dev)
                ri = torch.from_numpy(r_start[g][:, None] + np.arange(lr)[None, :]).to(dev)
                prod = torch.bmm(Q[qi], R[ri].transpose(1, 2))
                if self.similarity_bias:
                    prod += self.similarity_bias
                dst = torch.from_numpy(off[g][:, None] + np.arange(lq * lr)[None, :]).to(dev)
                sims[dst.reshape(-1)] = prod.reshape(-1)
        return sims, off, q_len.astype(np.int32), r_len.astype(np.int32)

    def localize_all(self, candidates: List[CandidatePair]) -> List[Match]:
        if not candidates:
            return []
        torch = _lib.require_cuda()
        sims, off, lq, lr = self._pack_similarities(candidates)
        dev = sims.device
        n = len(candidates)
        meta = torch.from_numpy(np.concatenate([lq, lr])).to(dev)
        res = self.model.align_device(sims, torch.from_numpy(off).to(dev), meta[:n], meta[n:], n,
                                      int(lq.max()), int(lr.max()), want_maxsim=True)
        boxes, n_boxes, maxsim, _ = res.to_host()
        matches = []
        for i, c in enumerate(candidates):
            query, ref = self.queries[c.query_id], self.refs[c.ref_id]
            for k in range(n_boxes[i]):
                x1, y1, x2, y2 = (int(v) for v in boxes[i, k])
                m = Match(query_id=c.query_id, ref_id=c.ref_id,
                          query_start=query.get_timestamps(x1)[0], query_end=query.get_timestamps(x2)[1],
                          ref_start=ref.get_timestamps(y1)[0], ref_end=ref.get_timestamps(y2)[1], score=0.0)
                matches.append(m._replace(score=self.score(c, m, (x1, y1, x2, y2), _BoxMax(maxsim[i, k]))))
        return matches

    def localize(self, candidate: CandidatePair) -> List[Match]:
        return self.localize_all([candidate])

    def score(self, candidate: CandidatePair, match: Match, box, similarity) -> float:
        return 1.0


class _BoxMax:
    """Stands in for the similarity matrix in `score(...)`: the only thing the reference's scorers take from it is
    `similarity[x1:x2, y1:y2].max()` of the box at hand, which the TN call already returned."""

    def __init__(self, value):
        self.value = value

    def __getitem__(self, _slices):
        return self

    def max(self):
        return self.value


class VCSLLocalizationMaxSim(VCSLLocalization):
    def score(self, candidate: CandidatePair, match: Match, box, similarity) -> float:
        x1, y1, x2, y2 = box
        return similarity[x1:x2, y1:y2].max() - self.similarity_bias


class VCSLLocalizationCandidateScore(VCSLLocalization):
    def score(self, candidate: CandidatePair, match: Match, box, similarity) -> float:
        return candidate.score
