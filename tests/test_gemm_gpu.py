"""GPU parity of the tcgen05 descriptor GEMM and its fused epilogues.

Inputs live on a coarse grid (multiples of 1/16 in [-4, 4]) so that every product and every partial sum
is exact in float32: the result is then independent of summation order and must equal the float32
reference BIT FOR BIT.  Arbitrary float32 inputs go through the fp16 split (hi.lo + lo.hi + hi.hi, 22 significant bits
per value, every product exact in the fp32 accumulator) and are compared with a float64 reference.  Tolerances stated
here, unit-norm 511-d rows: median |error| <= 2e-7; 2.5e-6 on IDENTICAL rows, where every product has the same sign and the
truncating accumulator of the tensor core loses ~4e-8 per K=16 step (a float32 sgemm: 1e-8 median, 6e-7 on identical rows).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def grid(rng, shape):
    return (rng.integers(-64, 65, size=shape) / 16.0).astype(np.float32)


def to_dev(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.parametrize("m,n,d", [(128, 256, 64), (300, 700, 512), (1, 5, 64), (129, 257, 130), (1000, 2049, 512),
                                   (64, 64, 512), (2, 3, 3)])
def test_store_exact_on_grid(m, n, d):
    from vsc2022_b200 import gemm
    rng = np.random.default_rng(m * 7 + n)
    a, b = grid(rng, (m, d)), grid(rng, (n, d))
    oa, ob = gemm.prepare_pair(to_dev(a), to_dev(b))
    assert not gemm.Pairing(oa, ob).split, "grid values fit the hi part: single pass expected"
    c = gemm.gemm_store(oa, ob).cpu().numpy()
    assert np.array_equal(c, a @ b.T)


def test_split_precision_on_arbitrary_fp32():
    from vsc2022_b200 import gemm
    rng = np.random.default_rng(3)
    a = rng.normal(size=(257, 511)).astype(np.float32)
    b = rng.normal(size=(513, 511)).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    b[:40] = a[100:140]                       # identical rows: every product positive, errors cannot cancel
    oa, ob = gemm.prepare_pair(to_dev(a), to_dev(b))
    p = gemm.Pairing(oa, ob)
    assert p.split and p.k == 3 * 512
    c = gemm.gemm_store(oa, ob).cpu().numpy().astype(np.float64)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    err = np.abs(c - ref)
    print("fp16 split: max |err|", float(err.max()), "on identical rows", float(err[100:140, :40].max()),
          "float32 matmul:", float(np.abs((a @ b.T).astype(np.float64) - ref).max()))
    assert err.max() <= 2.5e-6 and np.median(err) <= 2e-7, (float(err.max()), float(np.median(err)))
    # a differently scaled operand (raw, un-normalised descriptors): the error scales with the magnitudes
    a2, b2 = a * np.float32(37.5), b * np.float32(0.0123)
    oa, ob = gemm.prepare_pair(to_dev(a2), to_dev(b2))
    c2 = gemm.gemm_store(oa, ob).cpu().numpy().astype(np.float64)
    ref2 = a2.astype(np.float64) @ b2.astype(np.float64).T
    assert np.abs(c2 - ref2).max() <= 3.5e-6 * 37.5 * 0.0123, float(np.abs(c2 - ref2).max())


def test_rowmax_exact_on_grid():
    from vsc2022_b200 import gemm
    rng = np.random.default_rng(5)
    a, b = grid(rng, (777, 512)), grid(rng, (3001, 512))
    oa, ob = gemm.prepare_pair(to_dev(a), to_dev(b))
    got = gemm.gemm_rowmax(oa, ob).cpu().numpy()
    assert np.array_equal(got, (a @ b.T).max(axis=1))


@pytest.mark.parametrize("metric_l2", [False, True])
def test_emit_counts_and_entries(metric_l2):
    import torch
    from vsc2022_b200 import gemm
    rng = np.random.default_rng(9)
    a, b = grid(rng, (500, 128)), grid(rng, (1300, 128))
    da, db = to_dev(a), to_dev(b)
    oa, ob = gemm.prepare_pair(da, db)
    ip = a @ b.T
    if metric_l2:
        an, bn = (a * a).sum(1, dtype=np.float32), (b * b).sum(1, dtype=np.float32)
        s = (an[:, None] + bn[None, :] - np.float32(2.0) * ip).astype(np.float32)
        count_thr, emit_thr = np.float32(np.quantile(s, 0.02)), np.float32(np.quantile(s, 0.01))
        want = s < emit_thr
        n_count = int((s < count_thr).sum())
    else:
        s = ip
        count_thr, emit_thr = np.float32(np.quantile(s, 0.98)), np.float32(np.quantile(s, 0.99))
        want = s > emit_thr
        n_count = int((s > count_thr).sum())
    # output slots are claimed per warp in blocks of 256 whose unused tail is filled with a never-accepted score
    hits = gemm.HitBuffer(int(want.sum()) + 1184 * 256 + 100, da.device)
    gemm.gemm_emit(oa, ob, hits, float(count_thr), float(emit_thr), metric_l2=metric_l2,
                   a_norm=gemm.row_sqnorm(da) if metric_l2 else None,
                   b_norm=gemm.row_sqnorm(db) if metric_l2 else None, row_offset=10, col_offset=20)
    stored, counted = hits.read_counters()
    assert counted == n_count and int(want.sum()) <= stored <= hits.capacity
    sc = hits.score[:stored].cpu().numpy()
    real = np.isfinite(sc)                      # fillers are -inf (inner product) / +inf (L2)
    assert int(real.sum()) == int(want.sum())
    assert np.all(sc[~real] == (np.inf if metric_l2 else -np.inf))
    rows = hits.row[:stored].cpu().numpy()[real] - 10
    cols = hits.col[:stored].cpu().numpy()[real] - 20
    sc = sc[real]
    got = np.zeros_like(want)
    got[rows, cols] = True
    assert np.array_equal(got, want)
    assert np.array_equal(sc, s[rows, cols])


def test_emit_capacity_overflow_is_counted_not_written():
    from vsc2022_b200 import gemm
    rng = np.random.default_rng(11)
    a, b = grid(rng, (256, 64)), grid(rng, (512, 64))
    oa, ob = gemm.prepare_pair(to_dev(a), to_dev(b))
    hits = gemm.HitBuffer(1000, oa.panel.device)
    gemm.gemm_emit(oa, ob, hits, -1e10, -1e10)
    stored, counted = hits.read_counters()
    assert stored >= 256 * 512 and counted == 256 * 512   # claimed slots (incl. block fillers) / true hits


def test_growing_operand_pieces_equal_whole():
    """fp16 split panels converted piece by piece (first piece fixes the scale: vsc_prepare_operand_f16, then _more and the
    row-list form _rows) are bit-identical to the panel of one whole-matrix conversion when the first piece holds max|x|."""
    import torch
    from vsc2022_b200 import gemm
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    x = torch.randn((1500, 96), generator=g, device="cuda")
    x[0, 0] = 9.0                                                   # the largest value sits in the first piece
    whole = gemm.prepare(x, gemm.SIDE_B)
    grow = gemm.GrowingOperand(1500, 96, gemm.SIDE_B, x.device)
    grow.prepare_rows(x, 0, 100)                                    # vsc_prepare_operand_f16
    grow.prepare_rows(x[100:260], 100)                              # _more, rows given directly
    ranges = [(260 + 12 * i, 260 + 12 * i + 7) for i in range(80)]  # > 64 runs: one launch over a row list
    grow.prepare_ranges(x, ranges)
    grow.prepare_ranges(x, [(1220, 1500)])
    torch.cuda.synchronize()
    assert float(grow.inv_scale.item()) == float(whole.inv_scale.item())
    done = torch.zeros(1500, dtype=torch.bool, device="cuda")
    done[:260] = True
    done[1220:] = True
    for r0, r1 in ranges:
        done[r0:r1] = True
    assert torch.equal(grow.panel[done], whole.panel[done])
    assert grow.needs_split() and not grow.overflowed()
    late = gemm.GrowingOperand(64, 96, gemm.SIDE_A, x.device)
    late.prepare_rows(x[:32] * 1e-3, 0)
    late.prepare_rows(x[32:64] * 1e3, 32)                           # far outside the first piece's fp16 range
    assert late.overflowed()


def test_filtered_rowmax_is_float32_exact():
    """index.max_similarity on float32 descriptors: two single-product passes + exact re-score (gemm.rowmax_filtered).
    Against the float64 row maximum: float32-dot accuracy, on unit rows, on rows of very different norms, and with
    near-tied columns (several candidates per row); a row of identical columns overflows the candidate buffer and falls back."""
    import torch
    from vsc2022_b200 import gemm
    from vsc2022_b200.index import METRIC_INNER_PRODUCT, FlatIndex
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    unit = lambda n, d: torch.nn.functional.normalize(torch.randn((n, d), generator=g, device="cuda"), dim=1)
    q, b = unit(3000, 511), unit(20000, 511)
    b[5000:5040] = b[100] + 1e-4 * torch.randn((40, 511), generator=g, device="cuda")   # near ties around column 100
    q[7] = b[100]
    for scale_q, scale_b in ((1.0, 1.0), (37.0, 0.02)):
        xq, xb = q * scale_q, b * scale_b
        idx = FlatIndex(511, METRIC_INNER_PRODUCT)
        idx.add_device(xb, copy=False)
        got = idx.max_similarity(xq)
        want = (xq.double() @ xb.double().T).max(dim=1).values
        err = (got.double() - want).abs().max().item()
        assert err <= 4e-7 * scale_q * scale_b, err
        idx.filtered_rowmax = False
        split = idx.max_similarity(xq)
        assert (split.double() - want).abs().max().item() <= 4e-6 * scale_q * scale_b   # identical rows: the accumulator truncates
    # the filter really ran, found few candidates, and saw the planted near-ties
    oa, ob = gemm.prepare_pair(q, b)
    assert gemm.Pairing(oa, ob).split
    assert gemm.rowmax_filtered(q, b, oa, ob) is not None
    # every column identical: every column is a candidate -> buffer overflow -> fallback to the three-product GEMM
    same = unit(1, 511).repeat(4000, 1).contiguous()
    oa, ob = gemm.prepare_pair(q[:256].contiguous(), same)
    assert gemm.rowmax_filtered(q[:256].contiguous(), same, oa, ob) is None
    idx = FlatIndex(511, METRIC_INNER_PRODUCT)
    idx.add_device(same, copy=False)
    got = idx.max_similarity(q[:256].contiguous())
    want = (q[:256].double() @ same.double().T).max(dim=1).values
    assert (got.double() - want).abs().max().item() <= 2e-6
