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
    """Device-resident result of one TN batch (boxes, counts, MaxSim scores)."""

    def __init__(self, boxes, n_boxes, maxsim, status, box_cap):
        self.boxes, self.n_boxes, self.maxsim, self.status, self.box_cap = boxes, n_boxes, maxsim, status, box_cap

    def to_host(self):
        nb = self.n_boxes.cpu().numpy()
        bx = self.boxes.cpu().numpy().reshape(len(nb), self.box_cap, 4)
        ms = self.maxsim.cpu().numpy().reshape(len(nb), self.box_cap) if self.maxsim is not None else None
        st = self.status.cpu().numpy() if self.status is not None else None
        return bx, nb, ms, st


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
    boxes = torch.empty((max(n_pairs, 1), cap, 4), dtype=torch.int32, device=dev)
    n_boxes = torch.zeros((max(n_pairs, 1),), dtype=torch.int32, device=dev)
    maxsim = torch.zeros((max(n_pairs, 1), cap), dtype=torch.float32, device=dev) if want_maxsim else None
    status = torch.zeros((max(n_pairs, 1),), dtype=torch.int32, device=dev)
    s = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        rc = lib.vcsl_tn_batch(
            d_sims.data_ptr(), d_off.data_ptr(), d_lq.data_ptr(), d_lr.data_ptr(), n_pairs,
            int(max_lq), int(max_lr), ctypes.byref(params), boxes.data_ptr(), n_boxes.data_ptr(),
            maxsim.data_ptr() if want_maxsim else None, status.data_ptr(),
            1 if force_exact_order else 0, ctypes.c_void_p(s.cuda_stream))
    _lib.check(rc, "vcsl_tn_batch")
    return TnResult(boxes[:n_pairs], n_boxes[:n_pairs], maxsim[:n_pairs] if want_maxsim else None,
                    status[:n_pairs], cap)


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
