"""Freeze golden vectors produced by the UNMODIFIED reference running over the shims.

TEST INFRASTRUCTURE.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden

For every fixture it (1) runs the reference's own code (vsc.index / vsc.candidates /
vsc.baseline.score_normalization / vsc.baseline.localization / sscd_baseline glue),
(2) checks the self-contained restatements in oracle/search_numpy.py and
oracle/localize_numpy.py reproduce it exactly, (3) writes inputs + outputs to
tests/golden/*.npz.  The GPU box has no /root/reference; tests there read the npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLDEN = os.path.join(REPO, "tests", "golden")


def _matches_to_array(matches):
    # Match fields: query_id, ref_id, score, query_start, query_end, ref_start, ref_end
    return (np.array([[m.query_id, m.ref_id] for m in matches], dtype=np.int64).reshape(-1, 2),
            np.array([[m.query_start, m.query_end, m.ref_start, m.ref_end] for m in matches],
                     dtype=np.float64).reshape(-1, 4),
            np.array([m.score for m in matches], dtype=np.float32))


def golden_c1():
    from oracle import search_numpy, synth
    from vsc.baseline import sscd_baseline
    from vsc.baseline.score_normalization import score_normalize
    from vsc.candidates import CandidateGeneration, MaxScoreAggregation
    from vsc.index import VideoFeature, VideoIndex

    q, r, noise, ts = synth.c1_videos(seed=1)
    n_q, frames, dim = q.shape
    out = {"q": q, "r": r, "noise": noise, "timestamps": ts}

    def videos(x, base):
        return [VideoFeature(video_id=base + i, timestamps=ts, feature=x[i]) for i in range(len(x))]

    queries, refs, noises = videos(q, 0), videos(r, 100), videos(noise, 200)
    K = 1200 * n_q
    for tag, (qq, rr, sn) in {
        "raw": (queries, refs, False),
        "sn": (*score_normalize(queries, refs, noises, beta=1.2), True),
    }.items():
        if sn:
            out["sn_q"] = np.stack([v.feature for v in qq])
            out["sn_r"] = np.stack([v.feature for v in rr])
            oq, orr = search_numpy.score_normalize(list(q), list(r), list(noise), beta=1.2)
            assert all(np.array_equal(a, b.feature) for a, b in zip(oq, qq))
            assert all(np.array_equal(a, b.feature) for a, b in zip(orr, rr))
        index = VideoIndex(qq[0].feature.shape[1])
        index.add(rr)
        pms = index.search(qq, K)
        flat = [(pm.query_id, pm.ref_id, m.query_timestamps[0], m.query_timestamps[1],
                 m.ref_timestamps[0], m.ref_timestamps[1], m.score) for pm in pms for m in pm.matches]
        out[f"{tag}_pairmatch_ids"] = np.array([[f[0], f[1]] for f in flat], dtype=np.int64)
        out[f"{tag}_pairmatch_ts"] = np.array([f[2:6] for f in flat], dtype=np.float64)
        out[f"{tag}_pairmatch_score"] = np.array([f[6] for f in flat], dtype=np.float32)
        mine = search_numpy.search_pairs([v.feature for v in qq], [v.feature for v in rr], K)
        assert [(a, b + 100) for a, b, _ in mine] == [(pm.query_id, pm.ref_id) for pm in pms]
        assert [s for _, _, ms in mine for _, _, s in ms] == [f[6] for f in flat]

        cands = CandidateGeneration(rr, MaxScoreAggregation()).query(qq, global_k=K)
        out[f"{tag}_cand_ids"] = np.array([[c.query_id, c.ref_id] for c in cands], dtype=np.int64)
        out[f"{tag}_cand_score"] = np.array([c.score for c in cands], dtype=np.float32)
        mine = search_numpy.candidates([v.feature for v in qq], [v.feature for v in rr], K)
        assert [(a, b + 100, s) for a, b, s in mine] == [(c.query_id, c.ref_id, c.score) for c in cands]

        # kNN mode (index.py:108-117,167-177)
        pms = index.search(qq, -3)
        out[f"{tag}_knn_ids"] = np.array([[pm.query_id, pm.ref_id, len(pm.matches)] for pm in pms], dtype=np.int64)
        out[f"{tag}_knn_score"] = np.array([m.score for pm in pms for m in pm.matches], dtype=np.float32)

        matches = sscd_baseline.localize_and_verify(qq, rr, cands, score_normalization=sn)
        ids, tss, sc = _matches_to_array(matches)
        out[f"{tag}_match_ids"], out[f"{tag}_match_ts"], out[f"{tag}_match_score"] = ids, tss, sc
        print(f"c1/{tag}: {len(flat)} frame matches, {len(cands)} candidates, {len(matches)} localized matches")
    np.savez_compressed(os.path.join(GOLDEN, "c1_reference.npz"), **out)


def golden_tn():
    """Seeded similarity matrices -> boxes from the TN authority, through the reference's
    VCSLLocalizationMaxSim.localize_all for the Match rows (localization.py:56-91)."""
    from oracle import synth, tn_fast, tn_networkx
    from vsc.baseline.localization import VCSLLocalizationMaxSim
    from vsc.index import VideoFeature
    from vsc.metrics import CandidatePair

    rng = np.random.default_rng(4)
    shapes = [(32, 32), (45, 60), (60, 45), (7, 90), (90, 7), (3, 3), (1, 40), (40, 1), (5, 4),
              (128, 128), (300, 300), (300, 300), (257, 319)]
    configs = {"vsc": dict(tn_max_step=5, min_length=4), "default": dict()}
    out = {}
    for i, (lq, lr) in enumerate(shapes):
        sims = synth.sim_matrix(rng, lq, lr, dim=64, bias=0.5 if i % 3 else 0.0,
                                quant=(64.0 if i % 4 == 3 else 0.0))
        # the reference's np.argsort is non-stable: the contract needs default == stable here
        top = min(5, lr)
        assert np.array_equal(np.argsort(-sims)[:, :top], np.argsort(-sims, kind="stable")[:, :top]) \
            or i % 4 == 3, f"fixture {i} has top-k ties under the default sort"
        out[f"sims_{i}"] = sims
        for tag, cfg in configs.items():
            boxes = tn_networkx.tn(sims, **cfg)
            assert boxes == tn_fast.tn(sims, **cfg), (i, tag)
            out[f"boxes_{tag}_{i}"] = np.array(boxes, dtype=np.int32).reshape(-1, 4)
        print(f"tn fixture {i}: {lq}x{lr} vsc={len(out[f'boxes_vsc_{i}'])} default={len(out[f'boxes_default_{i}'])}")
    out["n"] = np.array(len(shapes))

    # Match rows through the reference's own localize_all on feature inputs (float32)
    feat_rng = np.random.default_rng(5)
    a = synth.unit_rows(feat_rng.normal(size=(45, 64)).astype(np.float32))
    b = synth.unit_rows(feat_rng.normal(size=(30, 64)).astype(np.float32))
    c = synth.unit_rows(feat_rng.normal(size=(60, 64)).astype(np.float32))
    a[20:30] = c[30:40]
    ts_a = np.arange(45) * 1.0
    ts_c = np.stack([np.arange(60) * 0.5, np.arange(60) * 0.5 + 0.5], axis=1)
    loc = VCSLLocalizationMaxSim(
        [VideoFeature(video_id=1, feature=a, timestamps=ts_a)],
        [VideoFeature(video_id=2, feature=b, timestamps=np.arange(30) * 1.0),
         VideoFeature(video_id=3, feature=c, timestamps=ts_c)], "TN", similarity_bias=0.5,
        tn_max_step=5, min_length=4, concurrency=1)
    matches = loc.localize_all([CandidatePair(1, 2, 1.0), CandidatePair(1, 3, 2.0)])
    ids, tss, sc = _matches_to_array(matches)
    out.update(loc_a=a, loc_b=b, loc_c=c, loc_ts_a=ts_a, loc_ts_c=ts_c, loc_match_ids=ids,
               loc_match_ts=tss, loc_match_score=sc)
    print(f"localize_all fixture: {len(matches)} matches")
    np.savez_compressed(os.path.join(GOLDEN, "tn_reference.npz"), **out)


def main():
    sys.path.insert(0, REPO)
    from oracle.run_reference_tests import reference_on_path
    reference_on_path("/root/reference")
    os.makedirs(GOLDEN, exist_ok=True)
    golden_c1()
    golden_tn()
    for fn in sorted(os.listdir(GOLDEN)):
        print(fn, os.path.getsize(os.path.join(GOLDEN, fn)))


if __name__ == "__main__":
    main()
