"""`vcsl.vta` mirror: video temporal alignment models, TN on the GPU.

Drop-in for the surface the reference uses (vsc/baseline/localization.py:44-46,58):

    model = build_vta_model("TN", concurrency=16, tn_max_step=5, min_length=4)
    results = model.forward_sim([(key, sim_matrix), ...])   # [(key, [[q0, r0, q1, r1], ...]), ...]

Same defaults as VCSL's TN (tn_max_step=10, tn_top_k=5, max_path=10, min_sim=0.2,
min_length=5, max_iou=0.3); results come back in input order with the key
echoed.  `concurrency` is accepted and ignored (one kernel launch aligns the whole
batch).  Similarities are handled as float32 (what vsc produces); other dtypes
are converted.  Only "TN" is provided; VCSL's other aligners (DTW, DP, HV, SPD)
are not on the vsc2022 path.
"""
import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib


class TnResult:
    """Device-resident result of one TN batch (boxes, counts, MaxSim scores) in ONE int32 buffer, so that a single
    device-to-host copy brings everything back."""

    def __init__(self, buf, n_pairs, box_cap, has_maxsim):
        self.buf, self.n_pairs, self.box_cap, self.has_maxsim = buf, n_pairs, box_cap, has_maxsim
        n, cap = max(n_pairs, 1), box_cap
        self._o_boxes, self._o_nb, self._o_ms, self._o_st = 0, n * cap * 4, n * cap * 4 + n, n * cap * 5 + n

    # device views
    @property
    def boxes(self):
        return self.buf[self._o_boxes:self._o_nb].view(-1, self.box_cap, 4)[:self.n_pairs]

    @property
    def n_boxes(self):
        return self.buf[self._o_nb:self._o_ms][:self.n_pairs]

    @property
    def maxsim(self):
        if not self.has_maxsim:
            return None
        torch = _lib.require_cuda()
        return self.buf[self._o_ms:self._o_st].view(torch.float32).view(-1, self.box_cap)[:self.n_pairs]

    @property
    def status(self):
        return self.buf[self._o_st:][:self.n_pairs]

    def to_host_async(self):
        """Start the device-to-host copy into pinned memory; returns wait() -> the tuple to_host() gives."""
        torch = _lib.require_cuda()
        host = torch.empty(self.buf.shape, dtype=self.buf.dtype, pin_memory=True)
        host.copy_(self.buf, non_blocking=True)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(self.buf.device))

        def wait():
            done.synchronize()
            return self._split(host.numpy())
        wait.ready = done.query     # non-blocking: has the copy landed?
        return wait

    def to_host(self):
        return self._split(self.buf.cpu().numpy())

    def _split(self, h):
        n, cap = self.n_pairs, self.box_cap
        bx = h[self._o_boxes:self._o_nb].reshape(-1, cap, 4)[:n]
        nb = h[self._o_nb:self._o_ms][:n]
        ms = h[self._o_ms:self._o_st].view(np.float32).reshape(-1, cap)[:n] if self.has_maxsim else None
        st = h[self._o_st:][:n]
        return bx, nb, ms, st


def _result_buffer(n_pairs: int, cap: int, dev):
    torch = _lib.require_cuda()
    n = max(n_pairs, 1)
    return torch.zeros((n * cap * 5 + 2 * n,), dtype=torch.int32, device=dev)


def tn_params(tn_max_step=10, tn_top_k=5, max_path=10, min_sim=0.2, min_length=5, max_iou=0.3) -> "_lib.TnParams":
    return _lib.TnParams(int(tn_max_step), int(tn_top_k), int(max_path), float(min_sim),
                         float(min_length), float(max_iou))


def tn_batch_device(d_sims, d_off, d_lq, d_lr, n_pairs: int, max_lq: int, max_lr: int,
                    params: "_lib.TnParams", want_maxsim: bool = True, force_exact_order: bool = False,
                    stream=None) -> TnResult:
    """Align `n_pairs` similarity matrices that are already resident in device memory.

    d_sims: float32 CUDA tensor holding all matrices; d_off (int64), d_lq, d_lr (int32) CUDA
    tensors give each pair's element offset and shape.  Asynchronous on `stream`.
    """
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = d_sims.device
    cap = params.max_path + 1
    res = TnResult(_result_buffer(n_pairs, cap, dev), n_pairs, cap, want_maxsim)
    base = res.buf.data_ptr()
    s = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        rc = lib.vcsl_tn_batch(
            d_sims.data_ptr(), d_off.data_ptr(), d_lq.data_ptr(), d_lr.data_ptr(), n_pairs,
            int(max_lq), int(max_lr), ctypes.byref(params), base + 4 * res._o_boxes, base + 4 * res._o_nb,
            base + 4 * res._o_ms if want_maxsim else None, base + 4 * res._o_st,
            1 if force_exact_order else 0, ctypes.c_void_p(s.cuda_stream))
    _lib.check(rc, "vcsl_tn_batch")
    return res


def tn_batch_from_features(q_panel, r_panel, k: int, d_q_start, d_lq, d_r_start, d_lr, n_pairs: int, max_lq: int,
                           max_lr: int, min_lr: int, bias: float, params: "_lib.TnParams", want_maxsim: bool = False,
                           sims_out=None, d_off=None, force_exact_order: bool = False, stream=None, fmt=None) -> TnResult:
    """Align `n_pairs` pairs whose similarity matrices are Q_p . R_p^T + bias of rows of two descriptor panels
    (CUDA tensors from gemm.prepare; `fmt`: their gemm.Pairing, None for plain bf16 panels of row stride k): the batch
    form of localization.py:57-58.  Asynchronous."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = q_panel.device
    cap = params.max_path + 1
    res = TnResult(_result_buffer(n_pairs, cap, dev), n_pairs, cap, want_maxsim)
    base = res.buf.data_ptr()
    s = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        off = fmt.col_off if fmt is not None else 0
        rc = lib.vcsl_tn_batch_from_features(
            q_panel.data_ptr() + off, q_panel.shape[0], r_panel.data_ptr() + off, r_panel.shape[0], int(k),
            d_q_start.data_ptr(), d_lq.data_ptr(), d_r_start.data_ptr(), d_lr.data_ptr(), n_pairs,
            int(max_lq), int(max_lr), int(min_lr), float(bias), ctypes.byref(params),
            sims_out.data_ptr() if sims_out is not None else None, d_off.data_ptr() if d_off is not None else None,
            base + 4 * res._o_boxes, base + 4 * res._o_nb, base + 4 * res._o_ms if want_maxsim else None,
            base + 4 * res._o_st, 1 if force_exact_order else 0, fmt.ref() if fmt is not None else None,
            ctypes.c_void_p(s.cuda_stream))
    _lib.check(rc, "vcsl_tn_batch_from_features")
    return res


def pair_similarity(q_panel, r_panel, k: int, d_q_start, d_lq, d_r_start, d_lr, n_pairs: int, max_lq: int, max_lr: int,
                    bias: float, sims_out, d_off, stream=None, fmt=None):
    """sims_out[d_off[p] + i * lr[p] + j] = Q[q_start[p] + i] . R[r_start[p] + j] + bias (localization.py:33-36,49-54)."""
    torch = _lib.require_cuda()
    dev = q_panel.device
    s = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        off = fmt.col_off if fmt is not None else 0
        rc = _lib.load().vsc_pair_similarity(
            q_panel.data_ptr() + off, q_panel.shape[0], r_panel.data_ptr() + off, r_panel.shape[0], int(k), d_q_start.data_ptr(),
            d_lq.data_ptr(), d_r_start.data_ptr(), d_lr.data_ptr(), n_pairs, int(max_lq), int(max_lr), float(bias),
            sims_out.data_ptr(), d_off.data_ptr(), fmt.ref() if fmt is not None else None,
            ctypes.c_void_p(s.cuda_stream))
    _lib.check(rc, "vsc_pair_similarity")
    return sims_out


def pack_sims(sims: Sequence[np.ndarray]):
    """Flatten host matrices into one pinned float32 buffer + offsets/shape arrays."""
    torch = _lib.require_cuda()
    n = len(sims)
    lq = np.fromiter((s.shape[0] for s in sims), dtype=np.int32, count=n)
    lr = np.fromiter((s.shape[1] for s in sims), dtype=np.int32, count=n)
    sizes = lq.astype(np.int64) * lr.astype(np.int64)
    padded = (sizes + 3) & ~np.int64(3)   # every matrix starts 16-byte aligned (TMA bulk copies)
    off = np.zeros(n, dtype=np.int64)
    if n:
        off[1:] = np.cumsum(padded[:-1])
    total = int(padded.sum())
    flat = torch.empty((max(total, 1) + 4,), dtype=torch.float32, pin_memory=True)
    view = flat.numpy()
    for s, o, sz in zip(sims, off, sizes):
        if s.ndim != 2:
            raise ValueError("similarity matrices must be 2-D")
        view[o:o + sz] = np.asarray(s, dtype=np.float32).reshape(-1)
    return flat, off, lq, lr


class TN:
    """Temporal-network aligner (GPU).  See module docstring."""

    def __init__(self, concurrency: int = 4, version: str = "v1", request_max: int = 20,
                 tn_max_step: int = 10, tn_top_k: int = 5, max_path: int = 10, min_sim: float = 0.2,
                 min_length: int = 5, max_iou: float = 0.3, device: Optional[str] = None, **unused):
        self.concurrency = concurrency  # accepted for API compatibility; unused
        self.params = tn_params(tn_max_step, tn_top_k, max_path, min_sim, min_length, max_iou)
        self.device = device
        self.force_exact_order = False
        self.last_result: Optional[TnResult] = None

    def _device(self):
        torch = _lib.require_cuda()
        return torch.device(self.device) if self.device else torch.device("cuda", torch.cuda.current_device())

    def align_device(self, d_sims, d_off, d_lq, d_lr, n_pairs, max_lq, max_lr, want_maxsim=True) -> TnResult:
        return tn_batch_device(d_sims, d_off, d_lq, d_lr, n_pairs, max_lq, max_lr, self.params,
                               want_maxsim, self.force_exact_order)

    def forward_packed(self, host_sims, off: np.ndarray, lq: np.ndarray, lr: np.ndarray, want_maxsim=False):
        """Host-buffer entry point: `host_sims` is ONE (ideally pinned) float32 torch tensor holding every
        matrix, matrix i at element offset off[i] (multiple of 4 for the fast path) with shape lq[i] x lr[i].
        Copies in, aligns, copies the boxes out.  Returns (boxes[n, max_path+1, 4], n_boxes[n], maxsim|None)."""
        torch = _lib.require_cuda()
        dev = self._device()
        n = len(off)
        d_sims = host_sims.to(dev, non_blocking=True)
        d_off = torch.from_numpy(np.ascontiguousarray(off, dtype=np.int64)).to(dev, non_blocking=True)
        meta = torch.from_numpy(np.concatenate([lq, lr]).astype(np.int32)).to(dev, non_blocking=True)
        res = self.align_device(d_sims, d_off, meta[:n], meta[n:], n, int(lq.max()) if n else 0,
                                int(lr.max()) if n else 0, want_maxsim=want_maxsim)
        self.last_result = res
        boxes, n_boxes, maxsim, _ = res.to_host()
        return boxes, n_boxes, maxsim

    def forward_sim(self, data: Sequence[Tuple[str, np.ndarray]]) -> List[Tuple[str, List[List[int]]]]:
        torch = _lib.require_cuda()
        data = list(data)
        if not data:
            return []
        dev = self._device()
        flat, off, lq, lr = pack_sims([s for _, s in data])
        d_sims = flat.to(dev, non_blocking=True)
        meta = torch.from_numpy(np.concatenate([lq, lr])).to(dev, non_blocking=True)
        d_off = torch.from_numpy(off).to(dev, non_blocking=True)
        n = len(data)
        res = self.align_device(d_sims, d_off, meta[:n], meta[n:], n, int(lq.max()), int(lr.max()))
        self.last_result = res
        boxes, n_boxes, _, _ = res.to_host()
        return [(key, boxes[i, :n_boxes[i]].tolist()) for i, (key, _) in enumerate(data)]


def build_vta_model(method: str = "DTW", concurrency: int = 4, **config) -> TN:
    if method != "TN":
        raise NotImplementedError(
            f"vsc2022_b200 provides the 'TN' aligner used by the vsc2022 baseline; got {method!r}")
    return TN(concurrency=concurrency, **config)
