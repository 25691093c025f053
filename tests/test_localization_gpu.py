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


@pytest.mark.parametrize("dtype,kind", [(np.float32, "grid"), (np.float16, "grid"), (np.float32, "gauss")])
def test_chunked_localize_all_over_lazily_uploaded_base_arrays(dtype, kind):
    """Large batches are aligned CHUNK pairs at a time while later descriptors are still going up (videos = row views
    of one base array per side, what storage.load_features returns; --store_fp16 arrays included).  Same Match rows as
    the oracle, as the unchunked call, and as a collection of loose arrays."""
    from oracle import localize_numpy
    from vsc2022_b200.index import VideoFeature
    from vsc2022_b200.localization import VCSLLocalizationMaxSim
    from vsc2022_b200.metrics import CandidatePair
    rng = np.random.default_rng(9)
    nq, nr, d = 24, 30, 64
    lens_q, lens_r = rng.integers(20, 70, size=nq), rng.integers(20, 70, size=nr)
    if kind == "grid":
        draw = lambda n: (rng.integers(-16, 17, size=(n, d)) / 16.0).astype(dtype)
    else:
        draw = lambda n: unit(rng.normal(size=(n, d))).astype(dtype)
    Q, R = draw(int(lens_q.sum())), draw(int(lens_r.sum()))
    q_at, r_at = np.concatenate([[0], np.cumsum(lens_q)]), np.concatenate([[0], np.cumsum(lens_r)])
    cand_ids = [(i, int(rng.integers(0, nr))) for i in range(nq) for _ in range(3)]
    for i, j in cand_ids[::2]:          # planted copies
        n = int(min(lens_q[i], lens_r[j], 30)) - 4
        Q[q_at[i] + 2:q_at[i] + 2 + n] = R[r_at[j] + 1:r_at[j] + 1 + n]
    tsq, tsr = np.arange(len(Q)) * 0.5, np.stack([np.arange(len(R)) * 1.0, np.arange(len(R)) * 1.0 + 0.75], axis=1)
    views = lambda: ([VideoFeature(video_id=i, feature=Q[q_at[i]:q_at[i + 1]], timestamps=tsq[q_at[i]:q_at[i + 1]]) for i in range(nq)],
                     [VideoFeature(video_id=100 + j, feature=R[r_at[j]:r_at[j + 1]], timestamps=tsr[r_at[j]:r_at[j + 1]]) for j in range(nr)])
    cands = [CandidatePair(i, 100 + j, 1.0) for i, j in cand_ids]
    cfg = dict(tn_max_step=5, min_length=4, similarity_bias=0.5)
    rows = lambda ms: [(m.query_id, m.ref_id, m.query_start, m.query_end, m.ref_start, m.ref_end, float(m.score)) for m in ms]

    qs, rs = views()
    whole = VCSLLocalizationMaxSim(qs, rs, "TN", **cfg)
    got_whole = whole.localize_all(cands)
    assert whole._dq.lazy is not None and whole._dr.lazy is not None
    assert whole._dr.h2d_bytes <= R.nbytes and whole._dq.h2d_bytes <= Q.nbytes

    qs, rs = views()
    chunked = VCSLLocalizationMaxSim(qs, rs, "TN", **cfg)
    chunked.CHUNK = 8                    # 72 pairs -> 9 chunks; blocks of 16 rows -> many partial uploads
    chunked._stores()[0].BLOCK = chunked._stores()[1].BLOCK = 16
    got_chunked = chunked.localize_all(cands)
    assert rows(got_chunked) == rows(got_whole)
    assert chunked._dq.h2d_bytes <= Q.nbytes and chunked._dr.h2d_bytes <= R.nbytes   # nothing crosses PCIe twice
    assert rows(chunked.localize_all(cands[5:40])) == rows(whole.localize_all(cands[5:40]))   # everything resident now

    loose_q = [VideoFeature(video_id=v.video_id, feature=np.array(v.feature), timestamps=np.array(v.timestamps)) for v in qs]
    loose_r = [VideoFeature(video_id=v.video_id, feature=np.array(v.feature), timestamps=np.array(v.timestamps)) for v in rs]
    assert rows(VCSLLocalizationMaxSim(loose_q, loose_r, "TN", **cfg).localize_all(cands)) == rows(got_whole)

    if kind == "grid":                   # every product exact: the oracle's rows, bit for bit
        pairs = [(Q[q_at[i]:q_at[i + 1]].astype(np.float32), tsq[q_at[i]:q_at[i + 1]], R[r_at[j]:r_at[j + 1]].astype(np.float32),
                  tsr[r_at[j]:r_at[j + 1]], 1.0) for i, j in cand_ids]
        want = localize_numpy.localize_all(pairs, 0.5, "max_sim", tn_max_step=5, min_length=4)
        flat = [(i, 100 + j) + tuple(row[:4]) + (float(row[4]),) for (i, j), rws in zip(cand_ids, want) for row in rws]
        assert rows(got_whole) == flat
    assert len(got_whole) >= 10


def test_growing_operand_scale_overflow_starts_over():
    """A later chunk with values far outside the first chunk's range (fp16 overflow under its scale): detected, the
    collection is prepared again as a whole and the rows equal the unchunked result."""
    from vsc2022_b200.index import VideoFeature
    from vsc2022_b200.localization import VCSLLocalizationMaxSim
    from vsc2022_b200.metrics import CandidatePair
    rng = np.random.default_rng(10)
    n, f, d = 12, 40, 32
    Q = (rng.integers(-16, 17, size=(n * f, d)) / 16.0).astype(np.float32)
    R = (rng.integers(-16, 17, size=(n * f, d)) / 16.0).astype(np.float32)
    Q[6 * f:] *= 64.0                    # the second half of the queries is 64x larger (first-chunk scale has 16x headroom)
    for i in range(n):
        R[i * f + 3:i * f + 33] = Q[i * f + 5:i * f + 35] / (64.0 if i >= 6 else 1.0)
    ts = np.arange(n * f) * 1.0
    mk = lambda X, base: [VideoFeature(video_id=base + i, feature=X[i * f:(i + 1) * f], timestamps=ts[i * f:(i + 1) * f]) for i in range(n)]
    cands = [CandidatePair(i, 100 + i, 1.0) for i in range(n)]
    cfg = dict(tn_max_step=5, min_length=4, similarity_bias=0.5)
    want = VCSLLocalizationMaxSim(mk(Q, 0), mk(R, 100), "TN", **cfg).localize_all(cands)
    loc = VCSLLocalizationMaxSim(mk(Q, 0), mk(R, 100), "TN", **cfg)
    loc.CHUNK = 3
    loc._stores()[0].BLOCK = loc._stores()[1].BLOCK = 8
    got = loc.localize_all(cands)
    assert got == want and len(want) >= n
    assert loc._dq.lazy is None          # settled after the overflow


def test_mixed_collection_settles_the_lazy_mirror():
    """A reference collection whose videos are partly row views of one array and partly arrays of their own: the first batch
    (views only) mirrors the base array lazily, the second batch brings a loose array in -- the mirror goes up whole, becomes
    an ordinary segment, and both batches give the rows of a collection made of loose arrays only."""
    from vsc2022_b200.index import VideoFeature
    from vsc2022_b200.localization import VCSLLocalizationCandidateScore
    from vsc2022_b200.metrics import CandidatePair
    rng = np.random.default_rng(31)
    grid = lambda n: (rng.integers(-16, 17, size=(n, 48)) / 16.0).astype(np.float32)
    R = grid(5 * 50)
    ts = np.arange(250) * 1.0
    extra = grid(64)
    q = grid(60)
    q[5:45] = R[110:150]                 # copy of reference video 2
    q2 = grid(60)
    q2[10:50] = extra[8:48]              # copy of the loose reference video
    refs_views = [VideoFeature(video_id=100 + i, feature=R[i * 50:(i + 1) * 50], timestamps=ts[i * 50:(i + 1) * 50]) for i in range(5)]
    refs_views.append(VideoFeature(video_id=199, feature=extra, timestamps=np.arange(64) * 1.0))
    queries = [VideoFeature(video_id=1, feature=q, timestamps=np.arange(60) * 1.0),
               VideoFeature(video_id=2, feature=q2, timestamps=np.arange(60) * 1.0)]
    cfg = dict(tn_max_step=5, min_length=4, similarity_bias=0.5)
    loc = VCSLLocalizationCandidateScore(queries, refs_views, "TN", **cfg)
    first = loc.localize_all([CandidatePair(1, 100 + i, 0.5) for i in range(5)])
    assert loc._dr.lazy is not None
    second = loc.localize_all([CandidatePair(2, 199, 0.9), CandidatePair(1, 102, 0.7)])
    assert loc._dr.lazy is None and len(loc._dr.segments) == 2
    loose = [VideoFeature(video_id=v.video_id, feature=np.array(v.feature), timestamps=np.array(v.timestamps)) for v in refs_views]
    ref_loc = VCSLLocalizationCandidateScore(queries, loose, "TN", **cfg)
    assert first == ref_loc.localize_all([CandidatePair(1, 100 + i, 0.5) for i in range(5)]) and len(first) >= 1
    assert second == ref_loc.localize_all([CandidatePair(2, 199, 0.9), CandidatePair(1, 102, 0.7)]) and len(second) >= 2


def test_custom_scorer_gets_the_similarity_matrix():
    """A Localization subclass with its own score() (the reference's calling convention, localization.py:66-76: candidate,
    match, box, similarity matrix): the matrices come back from the device, whole and chunked calls agree, and the values are
    the float32 matrices of Q.R^T + bias."""
    from vsc2022_b200.index import VideoFeature
    from vsc2022_b200.localization import VCSLLocalization
    from vsc2022_b200.metrics import CandidatePair

    class MeanOfBox(VCSLLocalization):
        def score(self, candidate, match, box, similarity):
            x1, y1, x2, y2 = box
            assert similarity.dtype == np.float32 and similarity.shape == (len(self.queries[candidate.query_id]),
                                                                            len(self.refs[candidate.ref_id]))
            return float(similarity[x1:x2 + 1, y1:y2 + 1].mean()) + candidate.score

    rng = np.random.default_rng(12)
    grid = lambda n: (rng.integers(-16, 17, size=(n, 32)) / 16.0).astype(np.float32)
    qf = [grid(int(rng.integers(30, 60))) for _ in range(6)]
    rf = [grid(int(rng.integers(30, 60))) for _ in range(7)]
    for i in range(6):
        n = min(len(qf[i]), len(rf[i])) - 6
        qf[i][3:3 + n] = rf[i][2:2 + n]
    qs = [VideoFeature(video_id=i, feature=x, timestamps=np.arange(len(x)) * 1.0) for i, x in enumerate(qf)]
    rs = [VideoFeature(video_id=50 + j, feature=x, timestamps=np.arange(len(x)) * 1.0) for j, x in enumerate(rf)]
    cands = [CandidatePair(i, 50 + j, 0.25 * j) for i in range(6) for j in range(7)]
    cfg = dict(tn_max_step=5, min_length=4, similarity_bias=0.5)
    whole = MeanOfBox(qs, rs, "TN", **cfg).localize_all(cands)
    chunked = MeanOfBox(qs, rs, "TN", **cfg)
    chunked.CHUNK = 5
    assert chunked.localize_all(cands) == whole and len(whole) >= 6
    by_pair = {}
    for m in whole:
        by_pair.setdefault((m.query_id, m.ref_id), []).append(m)
    for (qi, rj), ms in by_pair.items():
        sim = qs[qi].feature @ rs[rj - 50].feature.T + np.float32(0.5)       # grid data: exact in float32
        for m in ms:
            x1, x2, y1, y2 = int(m.query_start), int(m.query_end), int(m.ref_start), int(m.ref_end)
            assert m.score == float(sim[x1:x2 + 1, y1:y2 + 1].mean()) + 0.25 * (rj - 50)
