"""`vsc.baseline.localization` mirror (localization.py:16-96): align candidate pairs, emit Match rows.

Same classes and signatures as the reference.  `VCSLLocalization.localize_all` is one engine call per batch:

* the frame descriptors of the videos involved are uploaded once (a collection that consists of row views of one big
  array -- what `storage.load_features` returns -- goes up in a single copy; float16 descriptors are widened on the
  device) and turned into the K-major fp16 split panels of the tensor-core GEMM (three partial products unless every
  value fits the hi part, see gemm.py);
* `vcsl_tn_batch_from_features` multiplies every pair Q_p . R_p^T + bias on tcgen05 tensor cores and runs the temporal
  network on the result; for the usual shapes the Lq x Lr matrices never leave tensor memory
  (csrc/pair_gemm.cu).  They are written out only when a scorer reads them (MaxSim) -- and even then stay on the device:
  the box maxima come back with the boxes;
* only the boxes (a few integers per pair) return to the host, where they are mapped to timestamps with array
  operations (no per-pair Python work for pairs without a match).
"""
import abc
import gc
from typing import Dict, List

import numpy as np

from . import _lib, gemm
from .device_features import features_matrix, is_device_tensor, root_of as _root_of
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
    """Frame descriptors (and timestamps) of a video collection on the device, uploaded lazily.

    Rows live in one float32 matrix; `start[id]` / `length[id]` locate a video.  Videos that are row views of a
    shared base array are uploaded with that array in one copy; anything else is concatenated on the host first."""

    MAX_WASTE = 8   # upload a shared base whole unless it is this many times larger than what the batch needs

    def __init__(self, videos: Dict[object, VideoFeature], device):
        self.videos, self.device = videos, device
        self.start: Dict[object, int] = {}
        self.length: Dict[object, int] = {}
        self.segments = []          # device float32 [rows_i, d]
        self.rows = 0
        self.version = 0            # bumps whenever rows are added
        self._roots = {}            # id(root) -> (root, first device row)
        self._cat = None
        self._ts = [[], []]         # per segment: host arrays of frame start / end timestamps
        self._ts_cat = None
        self._root_cache = {}       # id(root) -> facts about that base array (keeps the array alive: ids stay unique)
        self.h2d_bytes = 0

    def _upload(self, host):
        """Append rows: a host array (copied up) or a device matrix (adopted as it is)."""
        torch = _lib.require_cuda()
        if is_device_tensor(host):
            d = host.to(self.device, torch.float32)
        else:
            if host.dtype not in (np.float32, np.float16):
                host = host.astype(np.float32)
            t = torch.from_numpy(host)
            self.h2d_bytes += t.numel() * t.element_size()
            d = t.to(self.device, non_blocking=True)
            if d.dtype != torch.float32:
                d = d.float()       # --store_fp16 descriptors (inference_impl.py:230-231): widened on the device
        first = self.rows
        self.segments.append(d)
        self.rows += d.shape[0]
        self._ts[0].append(np.zeros(d.shape[0]))
        self._ts[1].append(np.zeros(d.shape[0]))
        self._cat = self._ts_cat = None
        self.version += 1
        return first, len(self.segments) - 1

    def _stamp(self, seg: int, lo: int, ts: np.ndarray):
        n = ts.shape[0]
        if ts.ndim == 1:          # VideoMetadata.get_timestamps: (t, t) for instants, (t[0], t[1]) for intervals
            self._ts[0][seg][lo:lo + n] = ts
            self._ts[1][seg][lo:lo + n] = ts
        else:
            self._ts[0][seg][lo:lo + n] = ts[:, 0]
            self._ts[1][seg][lo:lo + n] = ts[:, 1]

    def _register(self, vid, first_row: int, seg: int, seg_first: int):
        v = self.videos[vid]
        self.start[vid], self.length[vid] = first_row, len(v)
        self._stamp(seg, first_row - seg_first, np.asarray(v.timestamps))
        self._ts_cat = None

    def prefetch(self, vid):
        """Start the upload of the base array `vid`'s descriptors are a view of (asynchronous from pinned memory), so
        that the copy runs while the host walks the rest of the collection."""
        if vid in self.start or is_device_tensor(self.videos[vid].feature):
            return
        where = _root_of(self.videos[vid].feature, self._root_cache)
        if where is not None and id(where[0]) not in self._roots and where[0].shape[0] <= self.MAX_WASTE * (1 << 20):
            first, seg = self._upload(where[0])
            self._roots[id(where[0])] = (where[0], first, seg)

    def ensure(self, ids):
        new = [i for i in dict.fromkeys(ids) if i not in self.start]
        if not new:
            return
        loose = []
        by_root = {}
        on_device = [i for i in new if is_device_tensor(self.videos[i].feature)]
        if on_device:   # descriptors that never left the GPU (score_normalize(on_device=True)): no copy over PCIe
            first, seg = self._upload(features_matrix([self.videos[i] for i in on_device], self.device))
            at = first
            for i in on_device:
                self._register(i, at, seg, first)
                at += len(self.videos[i])
            new = [i for i in new if not is_device_tensor(self.videos[i].feature)]
        for i in new:
            f = self.videos[i].feature
            if f.ndim != 2:
                raise ValueError("descriptors must be 2-D (frames x dimensions)")
            where = _root_of(f, self._root_cache)
            if where is None:
                loose.append(i)
            else:
                by_root.setdefault(id(where[0]), (where[0], []))[1].append((i, where[1]))
        for key, (root, members) in by_root.items():
            if key not in self._roots:
                need = sum(len(self.videos[i]) for i, _ in members)
                if root.shape[0] > self.MAX_WASTE * max(need, 1) and root.nbytes > (1 << 28):
                    loose.extend(i for i, _ in members)
                    continue
                first, seg = self._upload(root)
                self._roots[key] = (root, first, seg)
            _, first, seg = self._roots[key]
            # timestamps: when they are row views of one array laid out like the descriptors (storage.load_features),
            # one bulk copy; otherwise video by video
            ts_where = [_root_of(self.videos[i].timestamps, self._root_cache) for i, _ in members]
            ts_root = ts_where[0][0] if ts_where[0] is not None else None
            bulk = ts_root is not None and ts_root.shape[0] == root.shape[0] and all(
                w is not None and w[0] is ts_root and w[1] == row for w, (_, row) in zip(ts_where, members))
            if bulk:
                if ("ts", key) not in self._roots:
                    self._stamp(seg, 0, ts_root)
                    self._roots[("ts", key)] = ts_root
                self.start.update((i, first + row) for i, row in members)
                self.length.update((i, len(self.videos[i])) for i, _ in members)
                self._ts_cat = None
            else:
                for i, row in members:
                    self._register(i, first + row, seg, first)
        if loose:
            host = np.concatenate([np.asarray(self.videos[i].feature, dtype=np.float32) for i in loose])
            first, seg = self._upload(host)
            at = first
            for i in loose:
                self._register(i, at, seg, first)
                at += len(self.videos[i])

    def matrix(self):
        torch = _lib.require_cuda()
        if self._cat is None:
            dims = {s.shape[1] for s in self.segments}
            if len(dims) != 1:
                raise ValueError(f"descriptors of different dimensions in one collection: {sorted(dims)}")
            self._cat = self.segments[0] if len(self.segments) == 1 else torch.cat(self.segments)
        return self._cat

    def timestamps(self):
        if self._ts_cat is None:
            self._ts_cat = (np.concatenate(self._ts[0]), np.concatenate(self._ts[1]))
        return self._ts_cat


class VCSLLocalization(LocalizationWithMetadata):
    # what score() reads from its `similarity` argument: "none", "boxmax" (only similarity[x1:x2, y1:y2].max()) or
    # "full" (anything: the matrices are brought to the host; subclasses with their own score() get this)
    similarity_use = "none"

    def __init__(self, queries, refs, model_type, similarity_bias=0.0, **kwargs):
        super().__init__(queries, refs)
        from .vta import build_vta_model
        self.model = build_vta_model(model_type, **kwargs)
        self.similarity_bias = similarity_bias
        self._dq = self._dr = None
        self._panels = None

    def similarity(self, candidate: CandidatePair):
        """Add an optional similarity bias (some aligners do not tolerate negative values well)."""
        return super().similarity(candidate) + self.similarity_bias

    # ---- device path ----------------------------------------------------------------------------------------
    def _stores(self):
        if self._dq is None:
            dev = self.model._device()
            self._dq, self._dr = _DeviceVideos(self.queries, dev), _DeviceVideos(self.refs, dev)
        return self._dq, self._dr

    def _operands(self):
        """bf16 panels of all uploaded query / reference rows; the split is chosen for both sides together."""
        dq, dr = self._stores()
        key = (dq.version, dr.version)
        if self._panels is None or self._panels[0] != key:
            Q, R = dq.matrix(), dr.matrix()
            if Q.shape[1] != R.shape[1]:
                raise ValueError(f"query descriptors have {Q.shape[1]} dimensions, reference descriptors {R.shape[1]}")
            oq, orr = gemm.prepare_pair(Q, R)
            self._panels = (key, oq, orr, gemm.Pairing(oq, orr, precise=True))
        return self._panels[1], self._panels[2], self._panels[3]

    def _similarity_use(self):
        known = (VCSLLocalization.score, VCSLLocalizationMaxSim.score, VCSLLocalizationCandidateScore.score)
        return self.similarity_use if type(self).score in known else "full"

    def localize_all(self, candidates: List[CandidatePair]) -> List[Match]:
        if not candidates:
            return []
        torch = _lib.require_cuda()
        from .vta import tn_batch_from_features
        dq, dr = self._stores()
        dev = dq.device
        q_ids = [c.query_id for c in candidates]
        r_ids = [c.ref_id for c in candidates]
        dq.prefetch(q_ids[0])
        dr.prefetch(r_ids[0])
        dq.ensure(q_ids)
        dr.ensure(r_ids)
        oq, orr, pairing = self._operands()
        n = len(candidates)
        meta = np.empty((4, n), dtype=np.int32)      # q_start, lq, r_start, lr
        meta[0] = [dq.start[i] for i in q_ids]
        meta[1] = [dq.length[i] for i in q_ids]
        meta[2] = [dr.start[i] for i in r_ids]
        meta[3] = [dr.length[i] for i in r_ids]
        d_meta = torch.from_numpy(meta).to(dev, non_blocking=True)
        use = self._similarity_use()
        sims = d_off = off = None
        if use == "full":
            sizes = meta[1].astype(np.int64) * meta[3]
            padded = (sizes + 3) & ~np.int64(3)
            off = np.zeros(n, dtype=np.int64)
            off[1:] = np.cumsum(padded[:-1])
            sims = torch.empty((int(padded.sum()) + 4,), dtype=torch.float32, device=dev)
            d_off = torch.from_numpy(off).to(dev, non_blocking=True)
        res = tn_batch_from_features(
            oq.panel, orr.panel, pairing.k, d_meta[0], d_meta[1], d_meta[2], d_meta[3], n, int(meta[1].max()),
            int(meta[3].max()), int(meta[3].min()), float(self.similarity_bias), self.model.params,
            want_maxsim=(use == "boxmax"), sims_out=sims, d_off=d_off, force_exact_order=self.model.force_exact_order,
            fmt=pairing)
        self.model.last_result = res
        boxes, n_boxes, maxsim, _ = res.to_host()

        # ---- boxes -> Match rows (localization.py:61-78), vectorised over all boxes of the batch
        pair_of = np.repeat(np.arange(n), n_boxes)
        if len(pair_of) == 0:
            return []
        slot = np.arange(len(pair_of)) - np.repeat(np.cumsum(n_boxes) - n_boxes, n_boxes)
        bx = boxes[pair_of, slot].astype(np.int64)
        tq0, tq1 = dq.timestamps()
        tr0, tr1 = dr.timestamps()
        qs, rs = meta[0][pair_of].astype(np.int64), meta[2][pair_of].astype(np.int64)
        q_start, q_end = tq0[qs + bx[:, 0]].tolist(), tq1[qs + bx[:, 2]].tolist()
        r_start, r_end = tr0[rs + bx[:, 1]].tolist(), tr1[rs + bx[:, 3]].tolist()
        pairs_l = pair_of.tolist()
        qid, rid = [q_ids[p] for p in pairs_l], [r_ids[p] for p in pairs_l]
        scorer = type(self).score
        if scorer is VCSLLocalizationMaxSim.score:        # similarity[x1:x2, y1:y2].max() - bias, float32 like numpy's
            scores = list(maxsim[pair_of, slot] - self.similarity_bias)
        elif scorer is VCSLLocalizationCandidateScore.score:
            scores = [candidates[p].score for p in pairs_l]
        elif scorer is VCSLLocalization.score:
            scores = [1.0] * len(qid)
        else:                                              # a subclass with its own score(): the reference's calling convention
            host_sims = sims.cpu().numpy()
            scores = []
            for j, p in enumerate(pairs_l):
                lq_p, lr_p = int(meta[1][p]), int(meta[3][p])
                m = Match(query_id=qid[j], ref_id=rid[j], query_start=q_start[j], query_end=q_end[j],
                          ref_start=r_start[j], ref_end=r_end[j], score=0.0)
                similarity = host_sims[off[p]:off[p] + lq_p * lr_p].reshape(lq_p, lr_p)
                scores.append(self.score(candidates[p], m, tuple(int(v) for v in bx[j]), similarity))
        # Match is a NamedTuple: field order (query_id, ref_id, score, query_start, query_end, ref_start, ref_end).  Tens of
        # thousands of new container objects would trigger the cyclic collector again and again (6x slower): off for
        # the duration of the list construction.
        new, was_on = tuple.__new__, gc.isenabled()
        gc.disable()
        try:
            return [new(Match, row) for row in zip(qid, rid, scores, q_start, q_end, r_start, r_end)]
        finally:
            if was_on:
                gc.enable()

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
    similarity_use = "boxmax"

    def score(self, candidate: CandidatePair, match: Match, box, similarity) -> float:
        x1, y1, x2, y2 = box
        return similarity[x1:x2, y1:y2].max() - self.similarity_bias


class VCSLLocalizationCandidateScore(VCSLLocalization):
    def score(self, candidate: CandidatePair, match: Match, box, similarity) -> float:
        return candidate.score
