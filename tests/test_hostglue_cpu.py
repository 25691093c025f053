"""csrc/hostglue.c (the interpreter-bound loops of localize_all in C) against their Python twins.  CPU only."""
import gc
import sys

import numpy as np
import pytest


@pytest.fixture(scope="module")
def glue():
    from vsc2022_b200 import _hostglue, build_ext
    build_ext.build_hostglue()
    _hostglue._tried = False
    lib = _hostglue.load()
    assert lib is not None
    return lib


def test_match_rows_equal_the_python_construction(glue):
    from vsc2022_b200.metrics import Match
    rng = np.random.default_rng(0)
    n_pairs, n = 50, 4000
    q_ids = [f"Q{i:05d}" for i in range(n_pairs)]
    r_ids = [1000 + i for i in range(n_pairs)]                      # ids pass through untouched, whatever their type
    pair_of = np.sort(rng.integers(0, n_pairs, size=n)).astype(np.int64)
    scores = rng.random(n).astype(np.float32)
    t = [rng.random(n) * 100 for _ in range(4)]
    want = [tuple.__new__(Match, row) for row in zip([q_ids[p] for p in pair_of], [r_ids[p] for p in pair_of], list(scores),
                                                     *(x.tolist() for x in t))]
    for _ in range(3):
        got = glue.vsc_match_rows(Match, q_ids, r_ids, pair_of, scores, *t)
        assert got == want and len(got) == n
        assert all(type(m) is Match for m in got[:50])
        assert type(got[0].score) is np.float32 and type(got[0].query_start) is float and got[7].ref_id == r_ids[pair_of[7]]
        assert got[5]._asdict() == want[5]._asdict()
        del got
        gc.collect()
    before = sys.getrefcount(q_ids[0]), sys.getrefcount(Match)      # no leaked references: 50 more calls change nothing
    for _ in range(50):
        glue.vsc_match_rows(Match, q_ids, r_ids, pair_of, scores, *t)
    gc.collect()
    assert (sys.getrefcount(q_ids[0]), sys.getrefcount(Match)) == before
    assert glue.vsc_match_rows(Match, q_ids, r_ids, pair_of, [1.0] * n, *t)[3].score == 1.0
    assert glue.vsc_match_rows(Match, q_ids, r_ids, pair_of[:0], scores[:0], *(x[:0] for x in t)) == []
    with pytest.raises(ValueError):
        glue.vsc_match_rows(Match, q_ids, r_ids, pair_of, scores[:10], *t)
    with pytest.raises(ValueError):
        glue.vsc_match_rows(Match, q_ids, r_ids, pair_of, scores, t[0][:5], *t[1:])
    bad = pair_of.copy(); bad[3] = n_pairs
    with pytest.raises(IndexError):
        glue.vsc_match_rows(Match, q_ids, r_ids, bad, scores, *t)
    with pytest.raises(TypeError):
        glue.vsc_match_rows(dict, q_ids, r_ids, pair_of, scores, *t)


def test_scan_views_accepts_row_views_and_rejects_everything_else(glue):
    from vsc2022_b200.index import VideoFeature
    rng = np.random.default_rng(1)
    lens = rng.integers(3, 20, size=30)
    at = np.concatenate([[0], np.cumsum(lens)])
    root = rng.random((int(at[-1]), 16)).astype(np.float32)
    troot = np.arange(int(at[-1]) * 2, dtype=np.float64).reshape(-1, 2).copy()     # owns its data: views name it as .base
    mk = lambda i: VideoFeature(video_id=i, feature=root[at[i]:at[i + 1]], timestamps=troot[at[i]:at[i + 1]])
    videos = {i: mk(i) for i in range(30)}
    ids = [7, 3, 29, 0, 11]
    rows, ln = np.empty(5, np.int64), np.empty(5, np.int64)
    assert glue.vsc_scan_views(videos, ids, root, troot, np.ndarray, rows, ln) == 5
    assert rows.tolist() == [int(at[i]) for i in ids] and ln.tolist() == [int(lens[i]) for i in ids]
    t1 = np.arange(int(at[-1]), dtype=np.float64)                    # one-dimensional timestamps
    v1 = {i: VideoFeature(video_id=i, feature=root[at[i]:at[i + 1]], timestamps=t1[at[i]:at[i + 1]]) for i in range(30)}
    assert glue.vsc_scan_views(v1, ids, root, t1, np.ndarray, rows, ln) == 5

    def rejected(change):
        vs = dict(videos)
        vs[3] = change(videos[3])
        return glue.vsc_scan_views(vs, ids, root, troot, np.ndarray, rows, ln) == -1
    import dataclasses
    assert rejected(lambda v: dataclasses.replace(v, feature=np.array(v.feature)))                     # a copy: no base
    assert rejected(lambda v: dataclasses.replace(v, timestamps=troot[at[3] + 1:at[4] + 1]))           # timestamps one row off
    assert rejected(lambda v: dataclasses.replace(v, feature=root[at[3]:at[4], :8], timestamps=v.timestamps))   # column slice
    assert rejected(lambda v: dataclasses.replace(v, feature=root[at[3]:at[4]].view(np.int32)))        # same bytes, other dtype
    assert rejected(lambda v: dataclasses.replace(v, feature=root[at[3]:at[4]][::-1], timestamps=v.timestamps[::-1]))
    assert glue.vsc_scan_views(videos, ids + [999], root, troot, np.ndarray, np.empty(6, np.int64), np.empty(6, np.int64)) == -1
    other = root.copy()
    assert glue.vsc_scan_views(videos, ids, other, troot, np.ndarray, rows, ln) == -1                  # views of another array
