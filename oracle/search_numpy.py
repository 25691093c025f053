"""numpy restatement of the reference's stage-B search path (TEST INFRASTRUCTURE).

Self-contained (no /root/reference at run time) so that it travels to the GPU
box.  ``oracle/make_golden.py`` proves it equal to the UNMODIFIED reference
(run over the shims) on seeded inputs and freezes those outputs in
``tests/golden``.

Restated reference code
    vsc/index.py:96-140    VideoIndex.search            -> search_pairs()
    vsc/index.py:142-165   _global_threshold_knn_search -> global_topk()
    vsc/index.py:167-177   _knn_search                  -> knn()
    vsc/candidates.py:24-40 MaxScoreAggregation / CandidateGeneration.query -> candidates()
    vsc/baseline/score_normalization.py:31-105 score_normalize -> score_normalize()
The FAISS arithmetic behind them is the shim in oracle/shims/faiss (see its
header for the third-party pinning statement).
"""
from typing import List, Sequence, Tuple

import numpy as np

from oracle.shims import faiss as _faiss
from oracle.shims.faiss.contrib import exhaustive_search as _es

METRIC_INNER_PRODUCT = _faiss.METRIC_INNER_PRODUCT
METRIC_L2 = _faiss.METRIC_L2


def _rows(feats: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Concatenate per-video features; return (matrix, video-of-row, frame-of-row)."""
    mat = np.concatenate([np.asarray(f, dtype=np.float32) for f in feats], axis=0)
    vid = np.concatenate([np.full(len(f), i, dtype=np.int64) for i, f in enumerate(feats)])
    frm = np.concatenate([np.arange(len(f), dtype=np.int64) for f in feats])
    return mat, vid, frm


def global_topk(xq: np.ndarray, xb: np.ndarray, k: int, metric: int = METRIC_INNER_PRODUCT):
    """index.py:142-165 -- (query row, ref row, score) of the global top-k frame pairs.

    Ordered best first; equal scores keep (query row asc, ref row asc) because the
    reference builds the list in that order and Python's sort is stable.
    """
    index = _faiss.IndexFlat(xb.shape[1], metric)
    index.add(xb)
    keep_max = metric == METRIC_INNER_PRODUCT
    radius = -1e10 if keep_max else 1e10
    _, lims, dis, ids = _es.range_search_max_results(
        index, _es.exponential_query_iterator(xq), radius, max_results=2 * k, min_results=k)
    rows = np.repeat(np.arange(len(xq), dtype=np.int64), np.diff(lims.astype(np.int64)))
    order = np.argsort(-dis if keep_max else dis, kind="stable")[:k]
    return rows[order], ids[order], dis[order]


def knn(xq: np.ndarray, xb: np.ndarray, k: int, metric: int = METRIC_INNER_PRODUCT):
    """index.py:167-177 -- per query row its k best refs, emitted row by row."""
    index = _faiss.IndexFlat(xb.shape[1], metric)
    index.add(xb)
    dis, ids = index.search(xq, k)
    rows = np.repeat(np.arange(len(xq), dtype=np.int64), k)
    return rows, ids.ravel(), dis.ravel()


def search_pairs(query_feats, ref_feats, global_k: int, metric: int = METRIC_INNER_PRODUCT):
    """index.py:96-140 -- frame matches grouped per (query video, ref video).

    Returns a list of (query_video_index, ref_video_index, [(q_frame, r_frame, score), ...])
    in the reference's dict-insertion order (first appearance in the hit list).
    """
    xq, q_vid, q_frm = _rows(query_feats)
    xb, r_vid, r_frm = _rows(ref_feats)
    if global_k < 0:
        i, j, s = knn(xq, xb, -global_k, metric)
    else:
        i, j, s = global_topk(xq, xb, global_k, metric)
    groups = {}
    for ii, jj, ss in zip(i, j, s):
        groups.setdefault((int(q_vid[ii]), int(r_vid[jj])), []).append(
            (int(q_frm[ii]), int(r_frm[jj]), ss))
    return [(q, r, m) for (q, r), m in groups.items()]


def candidates(query_feats, ref_feats, global_k: int, metric: int = METRIC_INNER_PRODUCT):
    """candidates.py:24-40 -- (query video, ref video, max frame score), best first (stable)."""
    pairs = search_pairs(query_feats, ref_feats, global_k, metric)
    out = [(q, r, np.max([m[2] for m in ms])) for q, r, ms in pairs]
    return sorted(out, key=lambda c: c[2], reverse=True)


def l2_normalize_rows(x: np.ndarray) -> np.ndarray:
    """sklearn.preprocessing.normalize(x) (norm='l2', axis=1): rows / sqrt(sum x^2); zero rows stay."""
    x = np.asarray(x)
    norms = np.sqrt(np.einsum("ij,ij->i", x, x))
    norms[norms == 0.0] = 1.0
    return x / norms[:, np.newaxis]


def score_normalize(query_feats: List[np.ndarray], ref_feats: List[np.ndarray],
                    noise_feats: List[np.ndarray], l2_normalize: bool = True,
                    replace_dim: bool = True, beta: float = 1.0):
    """score_normalization.py:31-105 (the id-overlap guard lives with the caller)."""
    if replace_dim:
        stacked = np.concatenate(noise_feats, axis=0)
        drop = stacked.var(axis=0).argmin()
        query_feats, ref_feats, noise_feats = [
            [np.delete(f, drop, axis=1) for f in group]
            for group in (query_feats, ref_feats, noise_feats)]
    if l2_normalize:
        query_feats, ref_feats, noise_feats = [
            [l2_normalize_rows(f) for f in group]
            for group in (query_feats, ref_feats, noise_feats)]
    index = _faiss.IndexFlat(noise_feats[0].shape[1], METRIC_INNER_PRODUCT)
    for f in noise_feats:
        index.add(f)
    new_q = []
    for f in query_feats:
        sim, _ = index.search(f, 1)
        new_q.append(np.concatenate([f, -beta * sim[:, :1]], axis=1))
    new_r = [np.concatenate([f, np.ones_like(f[:, :1])], axis=1) for f in ref_feats]
    return new_q, new_r
