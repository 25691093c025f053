"""Host wrappers of the tensor-core descriptor GEMM (csrc/gemm_tc.cu, csrc/prep.cu).

Operands are prepared once (fp32 descriptors -> K-major bf16 panels) and reused across launches.
`precise=True` uses the 3-term bf16 split (fp32-class products, K' = 3K) whenever the values are not
bf16-representable; bf16-representable descriptors take the single-pass path automatically.
"""
import ctypes
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib

MODE_HI, MODE_SPLIT_A, MODE_SPLIT_B = 0, 1, 2


def _stream_ptr(torch, dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


@dataclass
class Operand:
    """A prepared GEMM operand: bf16 panel [rows][k] on the device."""
    panel: "object"      # torch bf16 tensor
    rows: int
    k: int               # padded inner dimension actually multiplied (kpad or 3*kpad)
    split: bool


def pad_k(d: int) -> int:
    return (d + 63) // 64 * 64


def prepare(x, mode: int = MODE_HI, want_flag: bool = False):
    """x: float32 CUDA tensor [n, d] (row stride may exceed d).  Returns (Operand, lo_flag tensor|None)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    n, d = x.shape
    kpad = pad_k(d)
    k = kpad if mode == MODE_HI else 3 * kpad
    out = torch.empty((max(n, 1), k), dtype=torch.bfloat16, device=x.device)
    flag = torch.zeros((1,), dtype=torch.int32, device=x.device) if want_flag else None
    with torch.cuda.device(x.device):
        rc = lib.vsc_prepare_operand(x.data_ptr(), n, d, x.stride(0), kpad, mode, out.data_ptr(),
                                     flag.data_ptr() if want_flag else None, _stream_ptr(torch, x.device))
    _lib.check(rc, "vsc_prepare_operand")
    return Operand(out, n, k, mode != MODE_HI), flag


def prepare_pair(a, b, precise: bool = True):
    """Prepare query-side `a` and reference-side `b`; split only if some value is not bf16-representable.
    Real descriptors never are, so the split panels are produced first (one pass, the flag comes with it) and only
    bf16-representable inputs pay a second pass for the short panels."""
    if not precise:
        return prepare(a, MODE_HI)[0], prepare(b, MODE_HI)[0]
    oa, fa = prepare(a, MODE_SPLIT_A, want_flag=True)
    ob, fb = prepare(b, MODE_SPLIT_B, want_flag=True)
    if not (int(fa.item()) | int(fb.item())):
        oa, _ = prepare(a, MODE_HI)
        ob, _ = prepare(b, MODE_HI)
    return oa, ob


def prepare_like(x, role: int, split: bool) -> Operand:
    """Prepare `x` for the given role (MODE_SPLIT_A / MODE_SPLIT_B) matching an existing partner."""
    return prepare(x, role if split else MODE_HI)[0]


def row_sqnorm(x):
    torch = _lib.require_cuda()
    out = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().vsc_row_sqnorm(x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), out.data_ptr(),
                                        _stream_ptr(torch, x.device))
    _lib.check(rc, "vsc_row_sqnorm")
    return out


def gemm_store(a: Operand, b: Operand):
    torch = _lib.require_cuda()
    assert a.k == b.k
    c = torch.empty((a.rows, b.rows), dtype=torch.float32, device=a.panel.device)
    with torch.cuda.device(c.device):
        rc = _lib.load().vsc_gemm_store(a.panel.data_ptr(), a.rows, b.panel.data_ptr(), b.rows, a.k, c.data_ptr(),
                                        b.rows, _stream_ptr(torch, c.device))
    _lib.check(rc, "vsc_gemm_store")
    return c


def gemm_rowmax(a: Operand, b: Operand):
    torch = _lib.require_cuda()
    assert a.k == b.k
    out = torch.empty((max(a.rows, 1),), dtype=torch.float32, device=a.panel.device)
    with torch.cuda.device(out.device):
        rc = _lib.load().vsc_gemm_rowmax(a.panel.data_ptr(), a.rows, b.panel.data_ptr(), b.rows, a.k,
                                         out.data_ptr(), _stream_ptr(torch, out.device))
    _lib.check(rc, "vsc_gemm_rowmax")
    return out[:a.rows]


def gemm_rowargmax(a: Operand, b: Operand):
    """(best score, its column) per row of a.b^T; lowest column on exact ties."""
    torch = _lib.require_cuda()
    assert a.k == b.k
    dev = a.panel.device
    score = torch.empty((max(a.rows, 1),), dtype=torch.float32, device=dev)
    col = torch.empty((max(a.rows, 1),), dtype=torch.int64, device=dev)
    scratch = torch.empty((max(a.rows, 1),), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().vsc_gemm_rowargmax(a.panel.data_ptr(), a.rows, b.panel.data_ptr(), b.rows, a.k,
                                            score.data_ptr(), col.data_ptr(), scratch.data_ptr(),
                                            _stream_ptr(torch, dev))
    _lib.check(rc, "vsc_gemm_rowargmax")
    return score[:a.rows], col[:a.rows]


class HitBuffer:
    """Device buffer of (score, row, col) survivors + the two counters the emit epilogue maintains."""

    def __init__(self, capacity: int, device):
        torch = _lib.require_cuda()
        self.capacity = int(capacity)
        self.score = torch.empty((self.capacity,), dtype=torch.float32, device=device)
        self.row = torch.empty((self.capacity,), dtype=torch.int32, device=device)
        self.col = torch.empty((self.capacity,), dtype=torch.int32, device=device)
        self.counters = torch.zeros((2,), dtype=torch.int64, device=device)  # [stored (claimed), counted]

    def read_counters(self):
        c = self.counters.cpu().numpy()
        return int(c[0]), int(c[1])


def gemm_emit(a: Operand, b: Operand, hits: HitBuffer, count_thr: float, emit_thr: float, metric_l2: bool = False,
              a_norm=None, b_norm=None, row_offset: int = 0, col_offset: int = 0, rows: Optional[slice] = None):
    """Append the scores of a (or a[rows]) x b beyond the thresholds to `hits` (asynchronous)."""
    torch = _lib.require_cuda()
    assert a.k == b.k
    a_panel, m, an = a.panel, a.rows, a_norm
    if rows is not None:
        a_panel = a.panel[rows]
        m = a_panel.shape[0]
        an = a_norm[rows] if a_norm is not None else None
    with torch.cuda.device(hits.score.device):
        rc = _lib.load().vsc_gemm_emit(
            a_panel.data_ptr(), m, b.panel.data_ptr(), b.rows, a.k,
            an.data_ptr() if an is not None else None, b_norm.data_ptr() if b_norm is not None else None,
            1 if metric_l2 else 0, float(count_thr), float(emit_thr), int(row_offset), int(col_offset),
            hits.score.data_ptr(), hits.row.data_ptr(), hits.col.data_ptr(), hits.capacity,
            hits.counters.data_ptr(), _stream_ptr(torch, hits.score.device))
    _lib.check(rc, "vsc_gemm_emit")
