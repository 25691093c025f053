"""GPU parity of stage C taken straight from descriptors (csrc/pair_gemm.cu through the C ABI):
per-pair Q.R^T + bias on tensor cores (localization.py:33-36,49-54) and the temporal network fed from tensor memory.

Bar: on grid descriptors (every product and partial sum exact in float32) the similarity matrices equal numpy's
bit for bit and the boxes equal the oracle's; on Gaussian descriptors the matrices are within the fp16-split bound and
the boxes equal the oracle run ON THE MATRICES THE SAME CALL WROTE (the row top-K out of tensor memory must agree with
the stored matrix exactly).
"""
import numpy as np
import pytest

from oracle import tn_fast

pytestmark = pytest.mark.gpu

VSC = dict(tn_max_step=5, tn_top_k=5, max_path=10, min_sim=0.2, min_length=4, max_iou=0.3)


def grid_feats(rng, n, dim, levels=16, scale=16.0):
    return (rng.integers(-levels, levels + 1, size=(n, dim)) / scale).astype(np.float32)


def make_videos(rng, lens, dim, gaussian=False, **kw):
    if gaussian:
        out = []
        for n in lens:
            x = rng.normal(size=(n, dim)).astype(np.float32)
            out.append((x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32))
        return out
    return [grid_feats(rng, n, dim, **kw) for n in lens]


def plant(rng, q, r, n):
    n = min(n, len(q), len(r))
    if n < 6:
        return
    a, b = int(rng.integers(0, len(q) - n + 1)), int(rng.integers(0, len(r) - n + 1))
    q[a:a + n] = r[b:b + n]


def run(q, r, pairs, bias, cfg, want_sims, want_maxsim, force_exact=False):
    import torch
    from vsc2022_b200 import gemm, vta
    dev = torch.device("cuda")
    Q, R = torch.from_numpy(np.concatenate(q)).to(dev), torch.from_numpy(np.concatenate(r)).to(dev)
    oq, orr = gemm.prepare_pair(Q, R)
    pairing = gemm.Pairing(oq, orr)
    qs, rs = np.cumsum([0] + [len(x) for x in q]), np.cumsum([0] + [len(x) for x in r])
    meta = np.array([[qs[i] for i, _ in pairs], [len(q[i]) for i, _ in pairs],
                     [rs[j] for _, j in pairs], [len(r[j]) for _, j in pairs]], dtype=np.int32)
    d_meta = torch.from_numpy(meta).to(dev)
    n = len(pairs)
    sims = d_off = off = None
    if want_sims:
        sizes = meta[1].astype(np.int64) * meta[3]
        padded = (sizes + 3) & ~np.int64(3)
        off = np.zeros(n, dtype=np.int64)
        off[1:] = np.cumsum(padded[:-1])
        sims = torch.full((int(padded.sum()) + 4,), np.nan, dtype=torch.float32, device=dev)
        d_off = torch.from_numpy(off).to(dev)
    res = vta.tn_batch_from_features(oq.panel, orr.panel, pairing.k, d_meta[0], d_meta[1], d_meta[2], d_meta[3], n,
                                     int(meta[1].max()), int(meta[3].max()), int(meta[3].min()), bias,
                                     vta.tn_params(**cfg), want_maxsim=want_maxsim, sims_out=sims, d_off=d_off,
                                     force_exact_order=force_exact, fmt=pairing)
    boxes, n_boxes, maxsim, status = res.to_host()
    got = [boxes[i, :n_boxes[i]].tolist() for i in range(n)]
    mats = None
    if want_sims:
        h = sims.cpu().numpy()
        mats = [h[off[p]:off[p] + meta[1][p] * meta[3][p]].reshape(meta[1][p], meta[3][p]) for p in range(n)]
    return got, maxsim, status, mats, pairing.split


def check_case(q, r, pairs, bias, cfg=VSC, exact_products=True):
    full = {**dict(tn_max_step=10, tn_top_k=5, max_path=10, min_sim=0.2, min_length=5, max_iou=0.3), **cfg}
    got, maxsim, status, mats, split = run(q, r, pairs, bias, cfg, want_sims=True, want_maxsim=True)
    ref = [np.matmul(q[i], r[j].T) + np.float32(bias) for i, j in pairs]
    for m, w in zip(mats, ref):
        if exact_products:
            assert np.array_equal(m, w)
        else:
            np.testing.assert_allclose(m, w, atol=3e-6, rtol=0)   # fp16 split; the bound is the truncating accumulator on identical rows
    want = tn_fast.tn_batch(mats, **full)          # the oracle on the matrices this very call produced
    bad = [p for p in range(len(pairs)) if got[p] != want[p]]
    assert not bad, (bad[:5], [(got[p], want[p], mats[p].shape) for p in bad[:2]])
    for p, m in enumerate(mats):
        for k, (x1, y1, x2, y2) in enumerate(got[p]):
            assert maxsim[p, k] == m[x1:x2, y1:y2].max()
    # without the matrices (nothing but descriptors in, boxes out) and through the exact-order kernel
    got2, _, status2, _, _ = run(q, r, pairs, bias, cfg, want_sims=False, want_maxsim=False)
    assert got2 == got
    got3, _, status3, _, _ = run(q, r, pairs, bias, cfg, want_sims=False, want_maxsim=False, force_exact=True)
    assert got3 == got and (status3 == 1).all()
    return got, status2, split


def test_uniform_300x300_grid():
    rng = np.random.default_rng(1)
    q, r = make_videos(rng, [300] * 6, 64), make_videos(rng, [300] * 8, 64)
    pairs = [(i, j) for i in range(6) for j in range(8)]
    for i, j in pairs[::3]:
        plant(rng, q[i], r[j], int(rng.integers(20, 80)))
    got, status, split = check_case(q, r, pairs, 0.5)
    assert not split
    assert sum(len(b) for b in got) >= 5
    assert (status == 0).mean() > 0.8, np.bincount(status, minlength=3)


def test_ragged_shapes_grid():
    rng = np.random.default_rng(2)
    lq = [1, 7, 33, 64, 127, 128, 129, 200, 256, 257, 300, 45]
    lr = [5, 8, 31, 32, 33, 96, 128, 129, 160, 161, 255, 256, 300, 319, 320, 321, 480, 512]
    q, r = make_videos(rng, lq, 64), make_videos(rng, lr, 64)
    pairs = [(int(rng.integers(len(lq))), int(rng.integers(len(lr)))) for _ in range(150)]
    pairs += [(i, j) for i in (0, 5, 9) for j in range(len(lr))]
    for i, j in pairs[::4]:
        plant(rng, q[i], r[j], int(rng.integers(6, 60)))
    check_case(q, r, pairs, 0.5)
    check_case(q, r, pairs, 0.0, cfg=dict())   # VCSL defaults: 64-bit masks


def test_small_matrices_40x40():
    rng = np.random.default_rng(3)
    q, r = make_videos(rng, [40] * 50, 128), make_videos(rng, [40] * 60, 128)
    pairs = [(int(rng.integers(50)), int(rng.integers(60))) for _ in range(500)]
    for i, j in pairs[::2]:
        plant(rng, q[i], r[j], int(rng.integers(8, 30)))
    got, _, _ = check_case(q, r, pairs, 0.5)
    assert sum(len(b) for b in got) >= 20


def test_heavy_ties():
    rng = np.random.default_rng(4)
    q = make_videos(rng, [150, 300, 60], 8, levels=1, scale=1.0)      # entries in {-1, 0, 1}: ties everywhere
    r = make_videos(rng, [300, 200, 17, 140], 8, levels=1, scale=1.0)
    r.append(np.zeros((250, 8), np.float32))                           # constant rows
    pairs = [(i, j) for i in range(3) for j in range(5)]
    check_case(q, r, pairs, 0.5)
    check_case(q, r, pairs, 0.25, cfg=dict(tn_max_step=3, tn_top_k=3, min_length=2))


def test_shapes_outside_the_direct_path():
    rng = np.random.default_rng(5)
    q = make_videos(rng, [50, 300, 12], 64)
    r = make_videos(rng, [3, 700, 64, 1, 513], 64)                     # lr < top_k, lr > 512
    pairs = [(i, j) for i in range(3) for j in range(5)]
    plant(rng, q[1], r[1], 60)
    check_case(q, r, pairs, 0.5)


def test_gaussian_descriptors_split_path():
    rng = np.random.default_rng(6)
    q, r = make_videos(rng, [300, 120, 300, 77], 512, gaussian=True), make_videos(rng, [300, 300, 96, 210], 512, gaussian=True)
    pairs = [(i, j) for i in range(4) for j in range(4)]
    for i, j in pairs[::2]:
        plant(rng, q[i], r[j], int(rng.integers(20, 80)))
    got, _, split = check_case(q, r, pairs, 0.5, exact_products=False)
    assert split
    assert sum(len(b) for b in got) >= 4


def test_pair_similarity_entry_point():
    import torch
    from vsc2022_b200 import gemm, vta
    rng = np.random.default_rng(7)
    q, r = make_videos(rng, [130, 1, 300], 192), make_videos(rng, [900, 2, 257], 192)
    dev = torch.device("cuda")
    oq, orr = gemm.prepare_pair(torch.from_numpy(np.concatenate(q)).to(dev), torch.from_numpy(np.concatenate(r)).to(dev))
    pairing = gemm.Pairing(oq, orr)
    pairs = [(i, j) for i in range(3) for j in range(3)]
    qs, rs = np.cumsum([0] + [len(x) for x in q]), np.cumsum([0] + [len(x) for x in r])
    meta = np.array([[qs[i] for i, _ in pairs], [len(q[i]) for i, _ in pairs],
                     [rs[j] for _, j in pairs], [len(r[j]) for _, j in pairs]], dtype=np.int32)
    sizes = meta[1].astype(np.int64) * meta[3]
    off = np.zeros(len(pairs), dtype=np.int64)
    off[1:] = np.cumsum(sizes[:-1])                                    # unpadded: unaligned rows and starts
    sims = torch.full((int(sizes.sum()),), np.nan, dtype=torch.float32, device=dev)
    d_meta = torch.from_numpy(meta).to(dev)
    vta.pair_similarity(oq.panel, orr.panel, pairing.k, d_meta[0], d_meta[1], d_meta[2], d_meta[3], len(pairs),
                        int(meta[1].max()), int(meta[3].max()), -0.25, sims, torch.from_numpy(off).to(dev), fmt=pairing)
    h = sims.cpu().numpy()
    assert not np.isnan(h).any()
    for p, (i, j) in enumerate(pairs):
        got = h[off[p]:off[p] + sizes[p]].reshape(meta[1][p], meta[3][p])
        assert np.array_equal(got, np.matmul(q[i], r[j].T) + np.float32(-0.25))
