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
import collections
import ctypes
import gc
import time
from typing import Dict, List

import numpy as np

from . import _hostglue, _lib, gemm
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
    BLOCK = 512     # rows per residency block of a lazily mirrored base array (1 MB at 512-d: PCIe-efficient copies)
    BRIDGE = 4      # blocks nobody asked for that a copy may take along to join two runs (few, large copies)

    def __init__(self, videos: Dict[object, VideoFeature], device, shared_copier=None):
        self.videos, self.device = videos, device
        self.shared_copier = shared_copier      # a list shared with the sibling store: holds the common copy stream once made
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
        # A collection that is row views of ONE base array (storage.load_features) is mirrored lazily: the device matrix
        # is allocated whole, rows go up in blocks when a batch first needs them, and the GEMM panels grow with them --
        # so a large localize_all can multiply its first pairs while the descriptors of the later ones are still on
        # their way.  {root, dev, first, resident (per block), fresh [(row0, row1) uploaded, not yet prepared]}
        self.lazy = None
        self._operand = None        # (key, gemm.Operand)

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

    def _mirror(self, root):
        """Device matrix for all rows of `root`, nothing copied yet."""
        torch = _lib.require_cuda()
        # zero-filled: rows that have not been uploaded yet read as zeros, so the panel conversion may cover a span that
        # contains some (zeros change neither the scale nor the flags) instead of one launch per uploaded run
        d = torch.zeros(root.shape, dtype=torch.float32, device=self.device)
        first = self.rows
        self.segments.append(d)
        self.rows += d.shape[0]
        self._ts[0].append(np.zeros(d.shape[0]))
        self._ts[1].append(np.zeros(d.shape[0]))
        self._cat = self._ts_cat = None
        self.version += 1
        blocks = (root.shape[0] + self.BLOCK - 1) // self.BLOCK
        # the blocks go up on a stream of their own, so that the copy of the next batch's rows overlaps the kernels of
        # this one; `events` = copies the compute stream has to wait for before it converts the `fresh` rows
        # (one copy stream for the query and the reference store of a localization object: their copies then cross PCIe in
        # the order they were issued -- chunk by chunk -- instead of two streams taking turns on the one copy engine)
        copier = self.shared_copier[0] if self.shared_copier else torch.cuda.Stream(device=self.device)
        if self.shared_copier is not None and not self.shared_copier:
            self.shared_copier.append(copier)
        copier.wait_stream(torch.cuda.current_stream(self.device))   # the block may have been another tensor's a moment ago
        self.lazy = {"root": root, "dev": d, "first": first, "resident": np.zeros(blocks, dtype=bool), "fresh": [],
                     "copier": copier, "events": []}
        return first, len(self.segments) - 1

    def _touch(self, lo: np.ndarray, hi: np.ndarray):
        """Rows [lo_i, hi_i) of the lazily mirrored base array must be on the device: copy the missing blocks
        (consecutive blocks in one asynchronous copy)."""
        torch = _lib.require_cuda()
        lz = self.lazy
        root, dev, resident = lz["root"], lz["dev"], lz["resident"]
        keep = hi > lo
        lo, hi = lo[keep], hi[keep]
        if len(lo) == 0:
            return
        edge = np.zeros(len(resident) + 1, dtype=np.int64)
        np.add.at(edge, lo // self.BLOCK, 1)
        np.add.at(edge, (hi - 1) // self.BLOCK + 1, -1)
        todo = (np.cumsum(edge[:-1]) > 0) & ~resident
        idx = np.flatnonzero(todo)
        if len(idx) == 0:
            return
        # few large copies instead of many small ones: a gap of up to BRIDGE blocks between two needed blocks goes along
        # when none of it is resident yet (every row still crosses PCIe at most once)
        gaps = np.flatnonzero((np.diff(idx) > 1) & (np.diff(idx) <= self.BRIDGE + 1))
        for g in gaps:
            if not resident[idx[g] + 1:idx[g + 1]].any():
                todo[idx[g] + 1:idx[g + 1]] = True
        idx = np.flatnonzero(todo)
        cuts = np.flatnonzero(np.diff(idx) > 1) + 1
        first = idx[np.concatenate([[0], cuts])] * self.BLOCK                                     # runs of consecutive blocks
        last = np.minimum((idx[np.concatenate([cuts, [len(idx)]]) - 1] + 1) * self.BLOCK, root.shape[0])
        ranges = np.ascontiguousarray(np.stack([first, last - first], axis=1), dtype=np.int64)
        with torch.cuda.stream(lz["copier"]):
            if root.dtype == np.float32:    # all runs in one engine call (a Python-level copy per run costs more than the copy)
                with torch.cuda.device(self.device):
                    rc = _lib.load().vsc_upload_rows(dev.data_ptr(), root.ctypes.data, root.strides[0], ranges.ctypes.data,
                                                     len(ranges), ctypes.c_void_p(lz["copier"].cuda_stream))
                _lib.check(rc, "vsc_upload_rows")
                self.h2d_bytes += int(ranges[:, 1].sum()) * root.strides[0]
            else:                           # --store_fp16 descriptors: half the bytes over PCIe, widened on the device
                for r0, n in ranges.tolist():
                    src = torch.from_numpy(root[r0:r0 + n])
                    self.h2d_bytes += src.numel() * src.element_size()
                    dev[r0:r0 + n].copy_(src.to(self.device, non_blocking=True))
            lz["fresh"].extend((int(r0), int(r0 + n)) for r0, n in ranges.tolist())
            done = torch.cuda.Event()
            done.record(lz["copier"])
        lz["events"].append(done)
        resident |= todo

    def _settle(self):
        """Another segment is about to join: the lazily mirrored array goes up whole and becomes an ordinary segment."""
        if self.lazy is not None:
            self._touch(np.array([0]), np.array([self.lazy["root"].shape[0]]))
            self._await_copies()
            self.lazy = None
            self._operand = None

    def __del__(self):
        # block copies nobody waited for (a batch that was prepared but never multiplied) must finish before the
        # allocator hands the mirror's memory to a tensor of the compute stream
        try:
            if self.lazy is not None and self.lazy["events"]:
                self._await_copies()
        except Exception:
            pass

    def _await_copies(self):
        """The compute stream waits for the block copies issued so far."""
        torch = _lib.require_cuda()
        stream = torch.cuda.current_stream(self.device)
        for ev in self.lazy["events"]:
            stream.wait_event(ev)
        self.lazy["events"] = []

    def dim(self) -> int:
        return self.segments[0].shape[1]

    def operand(self, side: int):
        """The fp16 split panels of the collection (gemm.Operand); with a lazy mirror only resident rows are valid."""
        if self.lazy is not None:
            lz = self.lazy
            if self._operand is None or self._operand[0] != ("lazy", id(lz)):
                self._operand = (("lazy", id(lz)), gemm.GrowingOperand(self.rows, lz["dev"].shape[1], side, self.device))
            op = self._operand[1]
            self._await_copies()
            fresh = lz["fresh"]                         # (a lazy mirror is the store's only segment: first == 0)
            if len(fresh) > 8:                          # many runs: one launch over the span they cover (see _mirror)
                fresh = [(min(r0 for r0, _ in fresh), max(r1 for _, r1 in fresh))]
            for r0, r1 in fresh:
                op.prepare_rows(lz["dev"], r0, r1 - r0)
            lz["fresh"] = []
            return op
        key = ("full", self.version)
        if self._operand is None or self._operand[0] != key:
            self._operand = (key, gemm.prepare(self.matrix(), side))
        return self._operand[1]

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
        if vid in self.start or is_device_tensor(self.videos[vid].feature) or not self.segments:
            return      # (an empty store mirrors its first base array lazily: ensure() decides)
        where = _root_of(self.videos[vid].feature, self._root_cache)
        if where is not None and id(where[0]) not in self._roots and where[0].shape[0] <= self.MAX_WASTE * (1 << 20):
            first, seg = self._upload(where[0])
            self._roots[id(where[0])] = (where[0], first, seg)

    def _ensure_views_of_one_array(self, new) -> bool:
        """Fast path of ensure(): every new video's descriptors are whole rows of ONE base array that this store mirrors
        lazily (or is empty and about to), timestamps likewise rows of one array laid out the same way -- what
        storage.load_features returns.  One tight loop per video; nothing is changed unless everything fits."""
        videos, nd = self.videos, np.ndarray
        v0 = videos[new[0]]
        root, troot = getattr(v0.feature, "base", None), getattr(v0.timestamps, "base", None)
        if root.__class__ is not nd or troot.__class__ is not nd or isinstance(root.base, nd) or isinstance(troot.base, nd):
            return False
        if (root.ndim != 2 or troot.ndim not in (1, 2) or troot.shape[0] != root.shape[0] or troot.dtype != np.float64
                or root.dtype not in (np.float32, np.float16) or not root.flags.c_contiguous or not troot.flags.c_contiguous):
            return False
        key = id(root)
        known = self._roots.get(key)
        if known is None and self.segments:
            return False
        if known is not None and (self.lazy is None or self.lazy["root"] is not root or self._roots.get(("ts", key)) is not troot):
            return False
        glue = _hostglue.load()
        if glue is not None:        # the per-video loop below, in C (csrc/hostglue.c)
            rows, lens = np.empty(len(new), dtype=np.int64), np.empty(len(new), dtype=np.int64)
            if glue.vsc_scan_views(videos, new, root, troot, nd, rows, lens) != len(new):
                return False
            return self._adopt_views(new, root, troot, key, known, rows, lens)
        rptr, rstride, tptr, tstride = root.ctypes.data, root.strides[0], troot.ctypes.data, troot.strides[0]
        fshape, fstrides, tshape, tstrides = root.shape[1:], root.strides, troot.shape[1:], troot.strides
        rows, lens = [], []
        for i in new:
            v = videos[i]
            f, t = v.feature, v.timestamps
            if f.__class__ is not nd or t.__class__ is not nd or f.base is not root or t.base is not troot:
                return False
            n = f.shape[0]
            if (n == 0 or f.shape[1:] != fshape or f.strides != fstrides or t.shape[0] != n or t.shape[1:] != tshape
                    or t.strides != tstrides):
                return False
            row, rem = divmod(f.__array_interface__["data"][0] - rptr, rstride)
            if rem or t.__array_interface__["data"][0] - tptr != row * tstride:
                return False
            rows.append(row)
            lens.append(n)
        return self._adopt_views(new, root, troot, key, known, np.array(rows, dtype=np.int64), np.array(lens, dtype=np.int64))

    def _adopt_views(self, new, root, troot, key, known, rows, lens) -> bool:
        if known is None:
            if root.shape[0] > self.MAX_WASTE * max(int(lens.sum()), 1) and root.nbytes > (1 << 28):
                return False
            first, seg = self._mirror(root)
            self._roots[key] = (root, first, seg)
            self._ts[0][seg], self._ts[1][seg] = (troot, troot) if troot.ndim == 1 else (troot[:, 0], troot[:, 1])
            self._roots[("ts", key)] = troot
        else:
            first = known[1]
        self._touch(rows, rows + lens)
        self.start.update(zip(new, (rows + first).tolist()))
        self.length.update(zip(new, lens.tolist()))
        self._ts_cat = None
        return True

    def ensure(self, ids):
        new = [i for i in dict.fromkeys(ids) if i not in self.start]
        if not new or self._ensure_views_of_one_array(new):
            return
        loose = []
        by_root = {}
        videos, nd = self.videos, np.ndarray
        on_device = [] if all(videos[i].feature.__class__ is nd for i in new) else \
            [i for i in new if is_device_tensor(videos[i].feature)]
        if on_device:   # descriptors that never left the GPU (score_normalize(on_device=True)): no copy over PCIe
            self._settle()
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
        lazy_ok = not self.segments and len(by_root) == 1 and not loose
        for key, (root, members) in by_root.items():
            if key not in self._roots:
                need = sum(len(self.videos[i]) for i, _ in members)
                if root.shape[0] > self.MAX_WASTE * max(need, 1) and root.nbytes > (1 << 28):
                    loose.extend(i for i, _ in members)
                    continue
                if lazy_ok and root.ndim == 2 and root.dtype in (np.float32, np.float16):
                    first, seg = self._mirror(root)
                else:
                    self._settle()
                    first, seg = self._upload(root)
                self._roots[key] = (root, first, seg)
            _, first, seg = self._roots[key]
            if self.lazy is not None and self.lazy["root"] is root:
                rows = np.fromiter((row for _, row in members), dtype=np.int64, count=len(members))
                lens = np.fromiter((len(self.videos[i]) for i, _ in members), dtype=np.int64, count=len(members))
                self._touch(rows, rows + lens)
            # timestamps: when they are row views of one array laid out like the descriptors (storage.load_features),
            # one bulk copy; otherwise video by video
            ts_where = [_root_of(self.videos[i].timestamps, self._root_cache) for i, _ in members]
            ts_root = ts_where[0][0] if ts_where[0] is not None else None
            bulk = ts_root is not None and ts_root.shape[0] == root.shape[0] and all(
                w is not None and w[0] is ts_root and w[1] == row for w, (_, row) in zip(ts_where, members))
            if bulk:
                if ("ts", key) not in self._roots:
                    if ts_root.dtype == np.float64 and self._ts[0][seg].shape[0] == ts_root.shape[0]:
                        # the segment is this base array: its timestamp array serves as it is (read only), no copy
                        self._ts[0][seg], self._ts[1][seg] = (ts_root, ts_root) if ts_root.ndim == 1 else (ts_root[:, 0], ts_root[:, 1])
                    else:
                        self._stamp(seg, 0, ts_root)
                    self._roots[("ts", key)] = ts_root
                self.start.update((i, first + row) for i, row in members)
                self.length.update((i, len(self.videos[i])) for i, _ in members)
                self._ts_cat = None
            else:
                for i, row in members:
                    self._register(i, first + row, seg, first)
        if loose:
            self._settle()
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
            one = len(self._ts[0]) == 1     # a single segment: its arrays as they are, no copy per batch
            self._ts_cat = (self._ts[0][0], self._ts[1][0]) if one else (np.concatenate(self._ts[0]), np.concatenate(self._ts[1]))
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
        self.d2h_bytes = 0
        self._no_sync = False
        self.profile = None          # set to a dict to collect the host-side phase times of localize_all (tools/probe_e2e.py)

    def similarity(self, candidate: CandidatePair):
        """Add an optional similarity bias (some aligners do not tolerate negative values well)."""
        return super().similarity(candidate) + self.similarity_bias

    # ---- device path ----------------------------------------------------------------------------------------
    def _stores(self):
        if self._dq is None:
            dev = self.model._device()
            copier = []
            self._dq, self._dr = _DeviceVideos(self.queries, dev, copier), _DeviceVideos(self.refs, dev, copier)
        return self._dq, self._dr

    def _operands(self):
        """fp16 split panels of the uploaded query / reference rows (prepared piece by piece while a lazily mirrored
        collection is still going up); the split is chosen for both sides together."""
        dq, dr = self._stores()
        if dq.dim() != dr.dim():
            raise ValueError(f"query descriptors have {dq.dim()} dimensions, reference descriptors {dr.dim()}")
        oq, orr = dq.operand(gemm.SIDE_A), dr.operand(gemm.SIDE_B)
        return oq, orr, gemm.Pairing(oq, orr, precise=True, split=True if self._no_sync else None)

    def _similarity_use(self):
        known = (VCSLLocalization.score, VCSLLocalizationMaxSim.score, VCSLLocalizationCandidateScore.score)
        return self.similarity_use if type(self).score in known else "full"

    # Batches of at least 2 * CHUNK pairs are aligned CHUNK pairs at a time: the descriptors of chunk c+1 cross PCIe and
    # its pairs are multiplied while the host turns the boxes of chunk c into Match rows.
    CHUNK = 1024

    def localize_all(self, candidates: List[CandidatePair]) -> List[Match]:
        if not candidates:
            return []
        self.d2h_bytes = 0
        n = len(candidates)
        if self.profile is not None:
            self.profile["t_start"] = time.perf_counter()
        self._no_sync = n >= 2 * self.CHUNK   # chunked: all three partial products without waiting for the operands' flags
        if n < 2 * self.CHUNK:
            matches = self._rows(self._launch(candidates))
        else:
            # Launch chunks ahead while the oldest outstanding one is still on the device, build its Match rows as soon
            # as its boxes have landed: the host never idles while descriptors are crossing PCIe.
            chunks = [candidates[at:at + self.CHUNK] for at in range(0, n, self.CHUNK)]
            depth = 2 if self._similarity_use() == "full" else len(chunks)    # whole matrices per job: keep few alive
            matches, pending, c = [], collections.deque(), 0
            while c < len(chunks) or pending:
                if pending and (c >= len(chunks) or len(pending) >= depth or pending[0][5].ready()):
                    matches.extend(self._rows(pending.popleft()))
                else:
                    pending.append(self._launch(chunks[c]))
                    c += 1
        for store in self._stores():   # a later piece outside the fp16 range of the first piece's scale: start over, whole
            op = store._operand[1] if store._operand is not None else None
            if isinstance(op, gemm.GrowingOperand) and op.overflowed():
                for st in self._stores():
                    st._settle()
                return self.localize_all(candidates)
        return matches

    def _ensure(self, candidates: List[CandidatePair]):
        dq, dr = self._stores()
        q_ids = [c.query_id for c in candidates]
        r_ids = [c.ref_id for c in candidates]
        dq.prefetch(q_ids[0])
        dr.prefetch(r_ids[0])
        dq.ensure(q_ids)
        dr.ensure(r_ids)
        return q_ids, r_ids

    def _launch(self, candidates: List[CandidatePair]):
        """Upload what is missing, enqueue the per-pair GEMM + temporal network and the copy of the result."""
        torch = _lib.require_cuda()
        from .vta import tn_batch_from_features
        prof = self.profile
        t0 = time.perf_counter() if prof is not None else 0.0
        dq, dr = self._stores()
        dev = dq.device
        q_ids, r_ids = self._ensure(candidates)
        if prof is not None:
            t1 = time.perf_counter(); prof["ensure"] = prof.get("ensure", 0.0) + t1 - t0; t0 = t1
        oq, orr, pairing = self._operands()
        if prof is not None:
            t1 = time.perf_counter(); prof["operands"] = prof.get("operands", 0.0) + t1 - t0; t0 = t1
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
        self.d2h_bytes += res.buf.numel() * 4
        job = candidates, q_ids, r_ids, meta, res, res.to_host_async(), sims, off, pairing
        if prof is not None:
            prof["launch"] = prof.get("launch", 0.0) + time.perf_counter() - t0
        return job

    def _rows(self, job) -> List[Match]:
        """boxes -> Match rows (localization.py:61-78), vectorised over all boxes of the batch."""
        candidates, q_ids, r_ids, meta, res, wait, sims, off, _pairing = job
        dq, dr = self._stores()
        n = len(candidates)
        prof = self.profile
        t0 = time.perf_counter() if prof is not None else 0.0
        boxes, n_boxes, maxsim, _ = wait()
        if prof is not None:
            t1 = time.perf_counter(); prof["wait_for_results"] = prof.get("wait_for_results", 0.0) + t1 - t0; t0 = t1
            prof.setdefault("results_at_ms", []).append(round(1e3 * (t1 - prof.get("t_start", t1)), 2))
        pair_of = np.repeat(np.arange(n), n_boxes)
        if len(pair_of) == 0:
            return []
        slot = np.arange(len(pair_of)) - np.repeat(np.cumsum(n_boxes) - n_boxes, n_boxes)
        bx = boxes[pair_of, slot].astype(np.int64)
        tq0, tq1 = dq.timestamps()
        tr0, tr1 = dr.timestamps()
        qs, rs = meta[0][pair_of].astype(np.int64), meta[2][pair_of].astype(np.int64)
        scorer = type(self).score
        glue = _hostglue.load()
        if glue is not None and scorer in (VCSLLocalizationMaxSim.score, VCSLLocalizationCandidateScore.score, VCSLLocalization.score):
            # the row construction below, in C (csrc/hostglue.c): same objects, same types
            if scorer is VCSLLocalizationMaxSim.score:
                scores = maxsim[pair_of, slot] - self.similarity_bias          # float32: iterating yields numpy.float32
            elif scorer is VCSLLocalizationCandidateScore.score:
                all_scores = [c.score for c in candidates]
                scores = [all_scores[p] for p in pair_of.tolist()]
            else:
                scores = [1.0] * len(pair_of)
            was_on = gc.isenabled()
            gc.disable()
            try:
                return glue.vsc_match_rows(Match, q_ids, r_ids, np.ascontiguousarray(pair_of, dtype=np.int64), scores,
                                           tq0[qs + bx[:, 0]], tq1[qs + bx[:, 2]], tr0[rs + bx[:, 1]], tr1[rs + bx[:, 3]])
            finally:
                if was_on:
                    gc.enable()
                if prof is not None:
                    prof["match_rows"] = prof.get("match_rows", 0.0) + time.perf_counter() - t0
        q_start, q_end = tq0[qs + bx[:, 0]].tolist(), tq1[qs + bx[:, 2]].tolist()
        r_start, r_end = tr0[rs + bx[:, 1]].tolist(), tr1[rs + bx[:, 3]].tolist()
        pairs_l = pair_of.tolist()
        qid = np.fromiter(q_ids, dtype=object, count=n)[pair_of].tolist()     # ids pass through untouched (int, str, ...)
        rid = np.fromiter(r_ids, dtype=object, count=n)[pair_of].tolist()
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
            if prof is not None:
                prof["match_rows"] = prof.get("match_rows", 0.0) + time.perf_counter() - t0

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
