"""Host wrappers of the tensor-core descriptor GEMM (csrc/gemm_tc.cu, csrc/prep.cu).

Operands are prepared once (fp32 descriptors -> K-major fp16 split panels, include/vsc_b200.h vsc_gemm_format) and
reused across launches.  A pair of operands multiplies over k = 3*kpad (hi.hi + hi.lo + lo.hi: every product exact in
the fp32 accumulator, ~3e-7 absolute on unit-norm rows -- the level of an fp32 sgemm) unless every value of both
has at most 11 significant bits (grid / test data), in which case the hi parts alone (k = kpad, same panels) are
exact.  `precise=False` always takes the single pass.
"""
import ctypes
from typing import Optional

import numpy as np

from . import _lib

SIDE_A, SIDE_B = 0, 1     # query side / reference side of the split (vsc_prepare_operand_f16)


def _stream_ptr(torch, dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class Operand:
    """A prepared GEMM operand: fp16 panel [rows, 3*kpad] on the device + its power-of-two scale."""

    def __init__(self, panel, rows: int, kpad: int, inv_scale, lo_flag, side: int):
        self.panel, self.rows, self.kpad, self.inv_scale, self.lo_flag, self.side = panel, rows, kpad, inv_scale, lo_flag, side
        self._needs_split: Optional[bool] = None
        self._scales = {}

    @property
    def ld(self) -> int:
        return 3 * self.kpad

    def needs_split(self) -> bool:
        """True unless every value is exactly representable by the hi part (one 4-byte read back, cached)."""
        if self._needs_split is None:
            self._needs_split = bool(int(self.lo_flag.item()))
        return self._needs_split

    def ptr(self, split: bool) -> int:
        """Device address of the columns a GEMM reads: the whole panel, or only its last third (the hi parts)."""
        return self.panel.data_ptr() + (0 if split else 2 * self.kpad * 2)

    def rows_slice(self, r0: int, r1: int) -> "Operand":
        sub = Operand(self.panel[r0:r1], r1 - r0, self.kpad, self.inv_scale, self.lo_flag, self.side)
        sub._needs_split, sub._scales = self._needs_split, self._scales
        return sub


class GrowingOperand(Operand):
    """An operand converted piece by piece (vsc_prepare_operand_f16 for the first piece, which fixes the power-of-two
    scale, vsc_prepare_operand_f16_more for the rest): the panel covers `rows` rows from the start, rows that were never
    prepared hold garbage and must not be multiplied."""

    def __init__(self, rows: int, d: int, side: int, device):
        torch = _lib.require_cuda()
        kpad = pad_k(d)
        self.meta = torch.zeros((4,), dtype=torch.int32, device=device)
        panel = torch.empty((max(rows, 1), 3 * kpad), dtype=torch.float16, device=device)
        super().__init__(panel, rows, kpad, self.meta[0:1].view(torch.float32), self.meta[1:2], side)
        self.d, self.started = d, False

    def prepare_rows(self, x, r0: int, n: Optional[int] = None):
        """Convert rows [r0, r0 + n) of the collection (asynchronous).  x: the float32 CUDA matrix [rows, d] the panel
        mirrors (n given: only its address and row stride are used, no tensor slicing per call), or just the rows to
        convert."""
        torch = _lib.require_cuda()
        if n is None:
            n, src = x.shape[0], x.data_ptr()
        else:
            src = x.data_ptr() + r0 * x.stride(0) * 4
        assert x.is_cuda and x.dtype == torch.float32 and x.shape[1] == self.d and 0 <= r0 and r0 + n <= max(self.rows, 1)
        if n == 0:
            return
        lib = _lib.load()
        fn = lib.vsc_prepare_operand_f16_more if self.started else lib.vsc_prepare_operand_f16
        base = self.meta.data_ptr()
        with torch.cuda.device(x.device):
            rc = fn(src, n, self.d, x.stride(0), self.kpad, self.side, self.panel.data_ptr() + r0 * self.ld * 2,
                    base, base + 4, base + 8, _stream_ptr(torch, x.device))
        _lib.check(rc, "vsc_prepare_operand_f16(_more)")
        self.started = True

    def prepare_ranges(self, x, ranges):
        """Convert the row ranges [(r0, r1), ...] of the float32 CUDA matrix `x` the panel mirrors, in ONE launch
        (vsc_prepare_operand_f16_rows over a row list built here)."""
        torch = _lib.require_cuda()
        if not ranges:
            return
        if len(ranges) <= 64:       # a launch per range is cheaper than staging a row list (a pinned allocation + a copy)
            for r0, r1 in ranges:
                self.prepare_rows(x, r0, r1 - r0)
            return
        r = np.asarray(ranges, dtype=np.int64)
        lens = r[:, 1] - r[:, 0]
        total = int(lens.sum())
        if total == 0:
            return
        # row list: for every range its rows in order (start of the range + position inside it)
        starts = np.repeat(r[:, 0] - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens)
        rows = starts + np.arange(total, dtype=np.int64)
        assert x.is_cuda and x.dtype == torch.float32 and x.shape[1] == self.d and int(r[:, 1].max()) <= max(self.rows, 1)
        staged = torch.empty((total,), dtype=torch.int32, pin_memory=True)   # pinned: the copy must not block the host, which
        staged.numpy()[:] = rows                                                # is running ahead of the descriptor uploads
        d_rows = staged.to(x.device, non_blocking=True)
        base = self.meta.data_ptr()
        with torch.cuda.device(x.device):
            rc = _lib.load().vsc_prepare_operand_f16_rows(x.data_ptr(), d_rows.data_ptr(), total, self.d, x.stride(0), self.kpad,
                                                          self.side, self.panel.data_ptr(), base, base + 4, base + 8,
                                                          1 if self.started else 0, _stream_ptr(torch, x.device))
        _lib.check(rc, "vsc_prepare_operand_f16_rows")
        self.started = True

    def needs_split(self) -> bool:
        if not self._needs_split:          # the flag only ever goes up: once set, no more read-backs
            self._needs_split = bool(int(self.lo_flag.item()) & 1)
        return self._needs_split

    def overflowed(self) -> bool:
        """True when a later piece held a value outside the fp16 range under the first piece's scale (one read back)."""
        return bool(int(self.lo_flag.item()) & 2)


class Pairing:
    """How two operands multiply: inner dimension, the format struct of the C ABI (kept alive with its scale)."""

    def __init__(self, a: Operand, b: Operand, precise: bool = True, split: Optional[bool] = None):
        """`split`: None = decide from the operands' flags (a 4-byte read back the first time: waits for their
        preparation); True = all three partial products without asking (always exact, 3x the work when one would do)."""
        assert a.kpad == b.kpad, "operands of different dimensions"
        assert a.side == SIDE_A and b.side == SIDE_B, "first operand must be prepared as SIDE_A, second as SIDE_B"
        self.split = (precise and (a.needs_split() or b.needs_split())) if split is None else bool(split)
        self.k = 3 * a.kpad if self.split else a.kpad
        key = id(b.inv_scale)
        if key not in a._scales:
            a._scales[key] = (a.inv_scale * b.inv_scale, b.inv_scale)    # keeps b's tensor alive: ids stay unique
        self.scale = a._scales[key][0]
        self.fmt = _lib.GemmFormat(1, a.ld, b.ld, self.scale.data_ptr())
        self.col_off = 0 if self.split else 2 * a.kpad * 2      # bytes into a panel row where the multiplied columns start

    def ref(self):
        return ctypes.byref(self.fmt)


def pad_k(d: int) -> int:
    return (d + 63) // 64 * 64


def prepare(x, side: int) -> Operand:
    """x: float32 CUDA tensor [n, d] (row stride may exceed d)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and (x.shape[0] == 0 or x.stride(1) == 1)
    n, d = x.shape
    kpad = pad_k(d)
    out = torch.empty((max(n, 1), 3 * kpad), dtype=torch.float16, device=x.device)
    meta = torch.zeros((4,), dtype=torch.int32, device=x.device)     # [inv_scale (f32 bits), lo flag, absmax scratch, -]
    inv_scale = meta[0:1].view(torch.float32)
    with torch.cuda.device(x.device):
        rc = lib.vsc_prepare_operand_f16(x.data_ptr(), n, d, x.stride(0) if n else d, kpad, side, out.data_ptr(),
                                         inv_scale.data_ptr(), meta[1:2].data_ptr(), meta[2:3].data_ptr(),
                                         _stream_ptr(torch, x.device))
    _lib.check(rc, "vsc_prepare_operand_f16")
    return Operand(out, n, kpad, inv_scale, meta[1:2], side)


def prepare_pair(a, b, precise: bool = True):
    """Prepare query-side `a` and reference-side `b`.  Returns (Operand, Operand); Pairing(oa, ob, precise) decides the
    inner dimension."""
    return prepare(a, SIDE_A), prepare(b, SIDE_B)


def row_sqnorm(x):
    torch = _lib.require_cuda()
    out = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().vsc_row_sqnorm(x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), out.data_ptr(),
                                        _stream_ptr(torch, x.device))
    _lib.check(rc, "vsc_row_sqnorm")
    return out


def gemm_store(a: Operand, b: Operand, precise: bool = True):
    torch = _lib.require_cuda()
    p = Pairing(a, b, precise)
    c = torch.empty((a.rows, b.rows), dtype=torch.float32, device=a.panel.device)
    with torch.cuda.device(c.device):
        rc = _lib.load().vsc_gemm_store(a.ptr(p.split), a.rows, b.ptr(p.split), b.rows, p.k, c.data_ptr(),
                                        b.rows, p.ref(), _stream_ptr(torch, c.device))
    _lib.check(rc, "vsc_gemm_store")
    return c


def gemm_rowmax(a: Operand, b: Operand, precise: bool = True):
    torch = _lib.require_cuda()
    p = Pairing(a, b, precise)
    out = torch.empty((max(a.rows, 1),), dtype=torch.float32, device=a.panel.device)
    with torch.cuda.device(out.device):
        rc = _lib.load().vsc_gemm_rowmax(a.ptr(p.split), a.rows, b.ptr(p.split), b.rows, p.k,
                                         out.data_ptr(), p.ref(), _stream_ptr(torch, out.device))
    _lib.check(rc, "vsc_gemm_rowmax")
    return out[:a.rows]


def gemm_rowargmax(a: Operand, b: Operand, precise: bool = True):
    """(best score, its column) per row of a.b^T; lowest column on exact ties."""
    torch = _lib.require_cuda()
    p = Pairing(a, b, precise)
    dev = a.panel.device
    score = torch.empty((max(a.rows, 1),), dtype=torch.float32, device=dev)
    col = torch.empty((max(a.rows, 1),), dtype=torch.int64, device=dev)
    scratch = torch.empty((max(a.rows, 1),), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().vsc_gemm_rowargmax(a.ptr(p.split), a.rows, b.ptr(p.split), b.rows, p.k,
                                            score.data_ptr(), col.data_ptr(), scratch.data_ptr(), p.ref(),
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
              a_norm=None, b_norm=None, row_offset: int = 0, col_offset: int = 0, rows: Optional[slice] = None,
              pairing: Optional[Pairing] = None):
    """Append the scores of a (or a[rows]) x b beyond the thresholds to `hits` (asynchronous)."""
    torch = _lib.require_cuda()
    p = pairing if pairing is not None else Pairing(a, b)
    a_panel, m, an = a.panel, a.rows, a_norm
    if rows is not None:
        a_panel = a.panel[rows]
        m = a_panel.shape[0]
        an = a_norm[rows] if a_norm is not None else None
    with torch.cuda.device(hits.score.device):
        rc = _lib.load().vsc_gemm_emit(
            a_panel.data_ptr() + p.col_off, m, b.ptr(p.split), b.rows, p.k,
            an.data_ptr() if an is not None else None, b_norm.data_ptr() if b_norm is not None else None,
            1 if metric_l2 else 0, float(count_thr), float(emit_thr), int(row_offset), int(col_offset),
            hits.score.data_ptr(), hits.row.data_ptr(), hits.col.data_ptr(), hits.capacity,
            hits.counters.data_ptr(), p.ref(), _stream_ptr(torch, hits.score.device))
    _lib.check(rc, "vsc_gemm_emit")


def gemm_emit_rows(a: Operand, b: Operand, hits: HitBuffer, row_thr, pairing: Pairing):
    """Append every inner product of a x b beyond its ROW's threshold (row_thr: float32 CUDA [a.rows]) to `hits`."""
    torch = _lib.require_cuda()
    with torch.cuda.device(hits.score.device):
        rc = _lib.load().vsc_gemm_emit_rows(
            a.panel.data_ptr() + pairing.col_off, a.rows, b.ptr(pairing.split), b.rows, pairing.k, row_thr.data_ptr(),
            hits.score.data_ptr(), hits.row.data_ptr(), hits.col.data_ptr(), hits.capacity, hits.counters.data_ptr(),
            pairing.ref(), _stream_ptr(torch, hits.score.device))
    _lib.check(rc, "vsc_gemm_emit_rows")


# |single-product score - exact score| <= SINGLE_PASS_EPS * |a| * |b|: both hi parts carry a relative rounding error of at
# most 2^-11 (fp16, round to nearest), so every product is off by at most (2^-10 + 2^-22) |a_k b_k|, Cauchy-Schwarz bounds
# the sum; the rest (fp32 accumulation in the tensor core, measured <= 2e-6 on unit rows) is covered by the slack.
SINGLE_PASS_EPS = 2.0 ** -10 * 1.02 + 1e-5


def rowmax_filtered(xq, xb, a: Operand, b: Operand):
    """max_j <xq_i, xb_j> per row in float32 from TWO single-product GEMMs instead of one three-product GEMM: pass 1 gives the
    approximate row maxima, pass 2 lists, per row, the columns within twice the error bound of it (a superset of the true
    arg max: ~1.3 columns per row on Gaussian descriptors), vsc_rowmax_rescore takes their exact float32 inner products
    from the original matrices.  Returns None when the candidate buffer overflowed (caller falls back)."""
    torch = _lib.require_cuda()
    single = Pairing(a, b, precise=False)
    approx = gemm_rowmax(a, b, precise=False)
    bound = SINGLE_PASS_EPS * torch.sqrt(row_sqnorm(xq)) * torch.sqrt(row_sqnorm(xb).max())
    thr = (approx - 2.0 * bound).contiguous()
    from .index import emit_pad
    hits = HitBuffer(8 * a.rows + emit_pad(), xq.device)
    gemm_emit_rows(a, b, hits, thr, single)
    stored, _ = hits.read_counters()
    if stored > hits.capacity:
        return None
    out = torch.empty((max(a.rows, 1),), dtype=torch.float32, device=xq.device)
    keys = torch.empty((max(a.rows, 1),), dtype=torch.int32, device=xq.device)
    with torch.cuda.device(xq.device):
        rc = _lib.load().vsc_rowmax_rescore(xq.data_ptr(), xq.shape[0], xq.stride(0), xb.data_ptr(), xb.shape[0], xb.stride(0),
                                            xq.shape[1], hits.score.data_ptr(), hits.row.data_ptr(), hits.col.data_ptr(),
                                            min(stored, hits.capacity), keys.data_ptr(), out.data_ptr(),
                                            _stream_ptr(torch, xq.device))
    _lib.check(rc, "vsc_rowmax_rescore")
    return out[:a.rows]
