"""GPU parity of the TN alignment kernel (through the C ABI) against the oracle.

Bar: bit-exact boxes (integer index work); MaxSim box scores bit-exact (a max over
float32 inputs has no rounding).
"""
import numpy as np
import pytest

from oracle import synth, tn_fast, tn_networkx

pytestmark = pytest.mark.gpu

VSC_CFG = dict(tn_max_step=5, min_length=4)


def run_gpu(sims, force_exact=False, **cfg):
    from vsc2022_b200 import vta
    model = vta.build_vta_model("TN", concurrency=16, **cfg)
    model.force_exact_order = force_exact
    out = model.forward_sim([(f"k{i}", s) for i, s in enumerate(sims)])
    assert [k for k, _ in out] == [f"k{i}" for i in range(len(sims))]
    boxes, n_boxes, maxsim, status = model.last_result.to_host()
    return [b for _, b in out], maxsim, status


@pytest.mark.parametrize("force_exact", [False, True])
def test_golden_boxes(golden_tn, force_exact):
    n = int(golden_tn["n"])
    sims = [golden_tn[f"sims_{i}"] for i in range(n)]
    for tag, cfg in (("vsc", VSC_CFG), ("default", {})):
        got, _, _ = run_gpu(sims, force_exact=force_exact, **cfg)
        for i in range(n):
            want = golden_tn[f"boxes_{tag}_{i}"].tolist()
            assert got[i] == want, (tag, i, sims[i].shape)


def _random_cases(seed, count, max_len):
    rng = np.random.default_rng(seed)
    cases = []
    for it in range(count):
        lq, lr = int(rng.integers(1, max_len)), int(rng.integers(1, max_len))
        if it % 3:
            lr = max(8, lr & ~3)                # 16-byte aligned rows: eligible for the fast pipeline
        quant = [0.0, 8.0, 4.0, 64.0][it % 4]   # coarse grids force exact ties everywhere
        cases.append(synth.sim_matrix(rng, lq, lr, bias=[0.5, 0.0][it % 2], quant=quant))
    return cases


@pytest.mark.parametrize("cfg", [
    dict(tn_max_step=5, tn_top_k=5, min_length=4),
    dict(),                                             # VCSL defaults: 45 slots -> 64-bit masks
    dict(tn_max_step=3, tn_top_k=3, min_length=2),
    dict(tn_max_step=7, tn_top_k=8, min_length=3, max_path=4, min_sim=0.45, max_iou=0.1),
    dict(tn_max_step=13, tn_top_k=5, min_length=5),
])
@pytest.mark.parametrize("force_exact", [False, True])
def test_random_vs_oracle(cfg, force_exact):
    sims = _random_cases(seed=11, count=240, max_len=90)
    want = tn_fast.tn_batch(sims, **{**dict(tn_max_step=10, tn_top_k=5, max_path=10, min_sim=0.2,
                                            min_length=5, max_iou=0.3), **cfg})
    got, _, status = run_gpu(sims, force_exact=force_exact, **cfg)
    bad = [i for i in range(len(sims)) if got[i] != want[i]]
    assert not bad, (bad[:5], [(got[i], want[i]) for i in bad[:2]])
    if force_exact:
        assert (status == 1).all()
    else:
        print("status histogram (0 fast pipeline, 2 general, 1 exact):", np.bincount(status, minlength=3))


def test_tie_heavy_uses_exact_kernel_and_matches_networkx():
    rng = np.random.default_rng(3)
    sims = [synth.sim_matrix(rng, 40, 40, quant=4.0) for _ in range(40)]
    got, _, status = run_gpu(sims, **VSC_CFG)
    for s, g in zip(sims, got):
        assert g == tn_networkx.tn(s, **VSC_CFG)
    # informational: how many pairs needed the exact-order kernel
    print("pairs routed to exact-order kernel:", int(status.sum()), "of", len(sims))


def test_maxsim_scores():
    rng = np.random.default_rng(5)
    sims = [synth.sim_matrix(rng, int(rng.integers(20, 120)), int(rng.integers(20, 120))) for _ in range(64)]
    got, maxsim, _ = run_gpu(sims, **VSC_CFG)
    seen = 0
    for i, s in enumerate(sims):
        for k, (x1, y1, x2, y2) in enumerate(got[i]):
            assert maxsim[i, k] == s[x1:x2, y1:y2].max()
            seen += 1
    assert seen > 10


def test_full_size_pairs_300x300():
    rng = np.random.default_rng(4)
    sims = [synth.sim_matrix(rng, 300, 300) for _ in range(192)]
    want = tn_fast.tn_batch(sims, tn_max_step=5, tn_top_k=5, max_path=10, min_sim=0.2, min_length=4, max_iou=0.3)
    got, _, status = run_gpu(sims, **VSC_CFG)
    assert got == want
    assert sum(len(b) for b in got) > 100
    assert (status == 0).mean() > 0.9, "aligned 300x300 pairs should stay on the fast pipeline"


def test_long_rows_and_edge_shapes():
    rng = np.random.default_rng(6)
    sims = [synth.sim_matrix(rng, 12, 700), synth.sim_matrix(rng, 700, 12), synth.sim_matrix(rng, 64, 321),
            synth.sim_matrix(rng, 2, 2), np.zeros((9, 9), np.float32), np.full((30, 30), 0.75, np.float32),
            synth.sim_matrix(rng, 1, 1)]
    want = tn_fast.tn_batch(sims, tn_max_step=5, tn_top_k=5, max_path=10, min_sim=0.2, min_length=4, max_iou=0.3)
    got, _, _ = run_gpu(sims, **VSC_CFG)
    assert got == want


def test_empty_batch_and_bad_params():
    from vsc2022_b200 import _lib, vta
    assert vta.build_vta_model("TN").forward_sim([]) == []
    with pytest.raises(NotImplementedError):
        vta.build_vta_model("DTW")
    with pytest.raises(_lib.EngineError):
        vta.build_vta_model("TN", tn_top_k=9).forward_sim([("a", np.zeros((4, 4), np.float32))])
    with pytest.raises(_lib.EngineError):
        vta.build_vta_model("TN", tn_max_step=20).forward_sim([("a", np.zeros((4, 4), np.float32))])


def test_compact_graph_variant_matches():
    """csrc/tn_graph.cu (graph stage on a compact graph, by Kahn generation) against the oracle and the default kernels."""
    from vsc2022_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(8)
    sims = [synth.sim_matrix(rng, 300, 300, dim=512) for _ in range(96)] + _random_cases(seed=12, count=120, max_len=90)
    cfg = dict(tn_max_step=5, tn_top_k=5, max_path=10, min_sim=0.2, min_length=4, max_iou=0.3)
    want = tn_fast.tn_batch(sims, **cfg)
    try:
        _lib.check(lib.vsc_tn_set_graph_variant(1), "vsc_tn_set_graph_variant")
        got, maxsim, status = run_gpu(sims, **VSC_CFG)
    finally:
        lib.vsc_tn_set_graph_variant(0)
    assert got == want
    assert (status[:96] == 0).mean() > 0.9
    for i, s in enumerate(sims[:96]):
        for k, (x1, y1, x2, y2) in enumerate(got[i]):
            assert maxsim[i, k] == s[x1:x2, y1:y2].max()


def test_equal_distance_runs_stay_on_the_fast_pipeline():
    """A chain whose edges were spent leaves a run of nodes with EQUAL distances (dist(v) = dist(u) + 0): ties at the
    maximum between nodes of different Kahn generations, which the smallest generation resolves.  tests/golden/
    tn_equal_distance_run.npz is such a matrix (pair 2618 of the bench workload, as the tensor-core GEMM produced it)."""
    import os
    s = np.load(os.path.join(os.path.dirname(__file__), "golden", "tn_equal_distance_run.npz"))["sims"]
    sims = [s, np.ascontiguousarray(s[:, :296]), np.ascontiguousarray(s[:299])]
    got, _, status = run_gpu(sims, **VSC_CFG)
    want = tn_fast.tn_batch(sims, tn_max_step=5, tn_top_k=5, max_path=10, min_sim=0.2, min_length=4, max_iou=0.3)
    assert got == want
    assert (status == 0).all(), status


@pytest.mark.parametrize("pairs_per_warp", [1, 2, 4])
def test_dp_pairs_per_warp_variants(golden_tn, pairs_per_warp):
    """The longest-path kernel packs 1, 2 or 4 pairs into a warp depending on the batch size (small batches: a warp runs as long
    as its slowest pair).  Every packing against the oracle: goldens, tie-heavy grids (in-kernel exact tie order), ragged
    random shapes and full-size pairs."""
    from vsc2022_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(17)
    sims = [golden_tn[f"sims_{i}"] for i in range(int(golden_tn["n"]))]
    sims += [synth.sim_matrix(rng, 40, 40, quant=4.0) for _ in range(24)] + _random_cases(seed=13, count=90, max_len=90)
    sims += [synth.sim_matrix(rng, 300, 300) for _ in range(40)]
    cfg = dict(tn_max_step=5, tn_top_k=5, max_path=10, min_sim=0.2, min_length=4, max_iou=0.3)
    want = tn_fast.tn_batch(sims, **cfg)
    try:
        _lib.check(lib.vsc_tn_set_dp_pairs_per_warp(pairs_per_warp), "vsc_tn_set_dp_pairs_per_warp")
        got, _, status = run_gpu(sims, **VSC_CFG)
    finally:
        lib.vsc_tn_set_dp_pairs_per_warp(0)
    assert got == want
    assert (status == 0).sum() >= 40      # the 40 full-size pairs at least stay on the fast pipeline


def test_c4_full_size_two_kernels_agree():
    """BASELINE.json configs[3] at FULL size (8000 pairs of 300x300 resident in HBM): the fast pipeline (row top-K, edges,
    longest-path sweeps with the in-kernel tie order) and the one-CTA-per-pair exact-order kernel (literal Kahn positions) are
    independent implementations of the oracle's algorithm -- they must give identical boxes and MaxSim scores on every pair;
    512 of the pairs also go through the oracle itself.  Plus size-independent properties of the result."""
    import torch
    from vsc2022_b200 import vta, workloads
    w = workloads.tn_pairs_device(8000, 300, 300, seed=4, device=torch.device("cuda"))
    model = vta.build_vta_model("TN", **VSC_CFG)
    fast = model.align_device(w.sims, w.off, w.lq, w.lr, 8000, 300, 300, want_maxsim=True).to_host()
    model.force_exact_order = True
    exact = model.align_device(w.sims, w.off, w.lq, w.lr, 8000, 300, 300, want_maxsim=True).to_host()
    boxes, n_boxes, maxsim, status = fast
    assert (status == 0).all() and (exact[3] == 1).all()
    assert np.array_equal(n_boxes, exact[1])
    live = np.arange(boxes.shape[1])[None, :] < n_boxes[:, None]
    assert np.array_equal(boxes[live], exact[0][live]) and np.array_equal(maxsim[live], exact[2][live])
    b = boxes[live]
    assert n_boxes.max() <= 11 and (b[:, 0] <= b[:, 2]).all() and (b[:, 1] <= b[:, 3]).all() and b.min() >= 0 and b.max() < 300
    assert (np.minimum(b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]) > 4).all()            # min_length
    host = w.sims[:512 * 90000].cpu().numpy().reshape(512, 300, 300)
    want = tn_fast.tn_batch(list(host), tn_max_step=5, tn_top_k=5, max_path=10, min_sim=0.2, min_length=4, max_iou=0.3)
    assert all(boxes[i, :n_boxes[i]].tolist() == want[i] for i in range(512))
    assert n_boxes.sum() > 50_000
