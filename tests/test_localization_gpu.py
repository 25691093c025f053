"""GPU parity of the localization mirror (stage C glue) against the reference's test and golden vectors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def unit(x):
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def test_reference_localization_test():
    """tests/test_localization.py:24-66 (float64 features, planted copy a[20:30] = c[30:40], default TN)."""
    from vsc2022_b200.index import VideoFeature
    from vsc2022_b200.localization import VCSLLocalizationMaxSim
    from vsc2022_b200.metrics import CandidatePair
    rng = np.random.default_rng(0)
    a, b, c = unit(rng.normal(size=(45, 64))), unit(rng.normal(size=(30, 64))), unit(rng.normal(size=(60, 64)))
    a[20:30] = c[30:40]
    mk = lambda i, f: VideoFeature(video_id=i, feature=f, timestamps=np.arange(f.shape[0]) * 1.0)
    loc = VCSLLocalizationMaxSim([mk(1, a)], [mk(2, b), mk(3, c)], "TN")
    assert loc.localize(CandidatePair(1, 2, 1.0)) == []
    assert len(loc.localize(CandidatePair(1, 3, 2.0))) >= 1
    matches = loc.localize_all([CandidatePair(1, 2, 1.0), CandidatePair(1, 3, 2.0)])
    assert len(matches) >= 1 and all(m.query_id == 1 and m.ref_id == 3 for m in matches)


def test_golden_localize_all(golden_tn):
    from vsc2022_b200.index import VideoFeature
    from vsc2022_b200.localization import VCSLLocalizationMaxSim
    from vsc2022_b200.metrics import CandidatePair
    g = golden_tn
    loc = VCSLLocalizationMaxSim(
        [VideoFeature(video_id=1, feature=g["loc_a"], timestamps=g["loc_ts_a"])],
        [VideoFeature(video_id=2, feature=g["loc_b"], timestamps=np.arange(30) * 1.0),
         VideoFeature(video_id=3, feature=g["loc_c"], timestamps=g["loc_ts_c"])],
        "TN", similarity_bias=0.5, tn_max_step=5, min_length=4, concurrency=1)
    matches = loc.localize_all([CandidatePair(1, 2, 1.0), CandidatePair(1, 3, 2.0)])
    assert [[m.query_id, m.ref_id] for m in matches] == g["loc_match_ids"].tolist()
    got_ts = np.array([[m.query_start, m.query_end, m.ref_start, m.ref_end] for m in matches])
    assert np.array_equal(got_ts, g["loc_match_ts"])             # segment boundaries: exact
    # identical (copied) frames score 1.0 in the reference; the tensor-core accumulator truncates ~4e-8 per K=16 step
    # of a same-sign sum (tests/test_gemm_gpu.py), BASELINE.json asks for alignment scores within 1e-4
    np.testing.assert_allclose([m.score for m in matches], g["loc_match_score"], atol=3e-6, rtol=0)


@pytest.mark.parametrize("tag,sn", [("raw", False), ("sn", True)])
def test_golden_c1_matching_track(golden_c1, tag, sn):
    """C1 end to end: (score-norm ->) search -> localize_and_verify, against the unmodified reference's output."""
    from vsc2022_b200 import sscd_baseline
    from vsc2022_b200.index import VideoFeature
    from vsc2022_b200.score_normalization import score_normalize
    g = golden_c1
    ts = g["timestamps"]
    vids = lambda x, base: [VideoFeature(video_id=base + i, timestamps=ts, feature=x[i]) for i in range(len(x))]
    queries, refs = vids(g["q"], 0), vids(g["r"], 100)
    if sn:
        queries, refs = score_normalize(queries, refs, vids(g["noise"], 200), beta=1.2)
    cands = sscd_baseline.search(queries, refs)
    assert [[c.query_id, c.ref_id] for c in cands] == g[f"{tag}_cand_ids"].tolist()
    matches = sscd_baseline.localize_and_verify(queries, refs, cands, score_normalization=sn)
    assert [[m.query_id, m.ref_id] for m in matches] == g[f"{tag}_match_ids"].tolist()
    got_ts = np.array([[m.query_start, m.query_end, m.ref_start, m.ref_end] for m in matches]).reshape(-1, 4)
    assert np.array_equal(got_ts, g[f"{tag}_match_ts"])
    np.testing.assert_allclose([m.score for m in matches], g[f"{tag}_match_score"], atol=1e-5, rtol=1e-6)


def test_ragged_batch_and_empty():
    from oracle import localize_numpy
    from vsc2022_b200.index import VideoFeature
    from vsc2022_b200.localization import VCSLLocalizationCandidateScore
    from vsc2022_b200.metrics import CandidatePair
    rng = np.random.default_rng(5)
    lens_q, lens_r = [40, 64, 7, 33], [48, 52, 90, 5, 36]
    grid = lambda n: (rng.integers(-16, 17, size=(n, 32)) / 16.0).astype(np.float32)
    q = [VideoFeature(video_id=i, feature=grid(n), timestamps=np.arange(n) * 0.5) for i, n in enumerate(lens_q)]
    r = [VideoFeature(video_id=100 + i, feature=grid(n), timestamps=np.arange(n) * 2.0) for i, n in enumerate(lens_r)]
    r[1].feature[10:40] = q[1].feature[20:50]          # planted copy
    loc = VCSLLocalizationCandidateScore(q, r, "TN", tn_max_step=5, min_length=4, similarity_bias=0.0)
    assert loc.localize_all([]) == []
    cands = [CandidatePair(a.video_id, b.video_id, float(a.video_id + b.video_id)) for a in q for b in r]
    got = loc.localize_all(cands)
    pairs = [(loc.queries[c.query_id].feature, loc.queries[c.query_id].timestamps, loc.refs[c.ref_id].feature,
              loc.refs[c.ref_id].timestamps, c.score) for c in cands]
    want = localize_numpy.localize_all(pairs, 0.0, "candidate", tn_max_step=5, min_length=4)
    flat = [(c.query_id, c.ref_id) + row[:5] for c, rows in zip(cands, want) for row in rows]
    assert [(m.query_id, m.ref_id, m.query_start, m.query_end, m.ref_start, m.ref_end, m.score) for m in got] == flat
    assert len(flat) >= 1
