"""End to end on the GPU, BASELINE.json configs[4] at test size: synthetic frames -> SSCD descriptors (stage A) ->
score normalisation + global top-K candidates (stage B) -> TN localization (stage C), through the reference-shaped API
(inference_impl.infer_videos, sscd_baseline.search / localize_and_verify).

Stage A has no reference-pinned oracle (random weights), so the checks are structural there (planted copies must be
found) and exact downstream: the oracle restatements of stages B and C run on the SAME descriptors and must give the
same candidate pairs (scores within the split-GEMM tolerance) and the same segment boundaries.
"""
import numpy as np
import pytest

from oracle import localize_numpy, search_numpy

pytestmark = pytest.mark.gpu

FRAMES, HW = 24, 64


@pytest.fixture(scope="module")
def descriptors():
    import torch
    from vsc2022_b200 import inference_impl
    from vsc2022_b200.sscd import SSCDResNet50, TorchReference
    ref = TorchReference(seed=3)
    model = SSCDResNet50(ref.trunk, ref.head)
    rng = np.random.default_rng(12)
    video = lambda: rng.integers(0, 256, size=(FRAMES, HW, HW, 3), dtype=np.uint8)
    ts = np.stack([np.arange(FRAMES) * 1.0, np.arange(FRAMES) * 1.0 + 1.0], axis=1)
    refs = [(f"R{i:06d}", ts, video()) for i in range(8)]
    queries = [(f"Q{i:06d}", ts, video()) for i in range(4)]
    noise = [(f"N{i:06d}", ts, video()) for i in range(4)]
    planted = {0: (5, 4, 2, 14), 2: (1, 8, 6, 10)}   # query -> (ref, query start, ref start, length)
    for q, (r, qs, rs, n) in planted.items():
        queries[q][2][qs:qs + n] = refs[r][2][rs:rs + n]
    run = lambda vids: inference_impl.infer_videos(vids, model, batch_size=64, device=torch.device("cuda"))
    return run(queries), run(refs), run(noise), planted


def test_frames_to_matches(descriptors):
    from vsc2022_b200 import sscd_baseline
    from vsc2022_b200.score_normalization import score_normalize
    queries, refs, noise, planted = descriptors
    assert all(v.feature.shape == (FRAMES, 512) and v.feature.dtype == np.float32 for v in queries + refs)
    # identical frames -> identical descriptors (the forward is batch-invariant)
    for q, (r, qs, rs, n) in planted.items():
        assert np.array_equal(queries[q].feature[qs:qs + n], refs[r].feature[rs:rs + n])

    sn_q, sn_r = score_normalize(queries, refs, noise, beta=1.2)
    cands = sscd_baseline.search(sn_q, sn_r)
    top = {(c.query_id, c.ref_id) for c in cands[:len(planted)]}
    assert top == {(f"Q{q:06d}", f"R{r:06d}") for q, (r, _, _, _) in planted.items()}

    # stage B against the oracle on the same descriptors: same pairs in the same order, scores within 1e-5
    oq, orf = search_numpy.score_normalize([v.feature for v in queries], [v.feature for v in refs],
                                           [v.feature for v in noise], beta=1.2)
    # (the appended -beta * max-similarity column comes from the 3-term bf16 split GEMM; random-weight descriptors of
    # random frames are nearly parallel, the worst case for its error: 4e-6 measured, 1e-5 asserted)
    for mine, theirs in zip(sn_q + sn_r, oq + orf):
        np.testing.assert_allclose(mine.feature, theirs, atol=1e-5, rtol=0)
    want = search_numpy.candidates(oq, orf, int(1200 * len(queries)))[:int(25 * len(queries))]
    want_score = {(f"Q{q:06d}", f"R{r:06d}"): float(s) for q, r, s in want}
    assert {(c.query_id, c.ref_id) for c in cands} == set(want_score)
    for c in cands:
        assert abs(c.score - want_score[(c.query_id, c.ref_id)]) <= 2e-5
    oracle_order = [want_score[(c.query_id, c.ref_id)] for c in cands]   # same order up to near-ties
    assert all(a >= b - 4e-5 for a, b in zip(oracle_order, oracle_order[1:]))

    # stage C: every planted copy yields a match that covers part of the planted range (random-weight descriptors of
    # random frames are nearly parallel, so TN extends the diagonal beyond the copy: boundaries are checked against the
    # oracle below, not against the plant), and the oracle agrees on ALL boxes
    matches = sscd_baseline.localize_and_verify(sn_q, sn_r, cands, score_normalization=True)
    by_pair = {}
    for m in matches:
        by_pair.setdefault((m.query_id, m.ref_id), []).append(m)
    for q, (r, qs, rs, n) in planted.items():
        assert any(m.query_start < qs + n and m.query_end > qs and m.ref_start < rs + n and m.ref_end > rs
                   for m in by_pair[(f"Q{q:06d}", f"R{r:06d}")])
    qd = {v.video_id: v for v in sn_q}
    rd = {v.video_id: v for v in sn_r}
    used = cands[:int(5 * len(queries))]
    pairs = [(qd[c.query_id].feature, qd[c.query_id].timestamps, rd[c.ref_id].feature, rd[c.ref_id].timestamps, c.score)
             for c in used]
    oracle = localize_numpy.localize_all(pairs, 0.5, "max_sim", tn_max_step=5, min_length=4)
    flat = [(c.query_id, c.ref_id) + row[:4] for c, rows in zip(used, oracle) for row in rows]
    assert [(m.query_id, m.ref_id, m.query_start, m.query_end, m.ref_start, m.ref_end) for m in matches] == flat
