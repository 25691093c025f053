"""Restatement of faiss.contrib.exhaustive_search (the two functions vsc uses).

TEST INFRASTRUCTURE (oracle).  Published algorithm (faiss >= 1.7.3,
contrib/exhaustive_search.py), restated from its documented behaviour; anchored
on the reference call site vsc/index.py:147-154 and pinned by the reference's
own known-answer test tests/test_candidates.py:72-83.

exponential_query_iterator(xq, start_bs=32, max_bs=20000)
    yields consecutive query slices of size 32, 64, 128, ...; the size doubles
    after a slice whenever it is still < max_bs (so the largest slice is 32768).

range_search_max_results(index, query_iterator, radius, max_results,
                         min_results, ngpu)
    for each slice: range_search with the CURRENT radius, append; when the
    running total exceeds max_results, tighten: radius := the
    (min_results+1)-th best stored score (numpy partition), then re-filter
    every stored slice with the STRICT test (> for IP, < for L2).  The radius
    only ever tightens.  Returns (radius, lims, D, I) over all queries.
"""
import numpy as np

from .. import METRIC_INNER_PRODUCT


def exponential_query_iterator(xq, start_bs=32, max_bs=20000):
    n = len(xq)
    size = start_bs
    at = 0
    while at < n:
        piece = xq[at:at + size]
        yield piece
        if size < max_bs:
            size *= 2
        at += len(piece)


def _tighten(batches, target, keep_max):
    every = np.hstack([d for _, d, _ in batches])
    assert len(every) > target
    if keep_max:
        every.partition(len(every) - target - 1)
        radius = float(every[-1 - target])
    else:
        every.partition(target)
        radius = float(every[target])
    total = 0
    for b, (counts, dis, ids) in enumerate(batches):
        keep = dis > radius if keep_max else dis < radius
        owner = np.repeat(np.arange(len(counts)), counts)
        counts = np.bincount(owner[keep], minlength=len(counts)).astype(counts.dtype)
        batches[b] = (counts, dis[keep], ids[keep])
        total += int(keep.sum())
    return radius, total


def range_search_max_results(index, query_iterator, radius, max_results=None,
                             min_results=None, shard=False, ngpu=0, clip_to_min=False):
    if min_results is None:
        min_results = int(0.8 * max_results)
    if max_results is None:
        max_results = int(min_results * 1.5)
    keep_max = index.metric_type == METRIC_INNER_PRODUCT
    batches = []
    total = 0
    for piece in query_iterator:
        lims, dis, ids = index.range_search(piece, radius)
        counts = (lims[1:] - lims[:-1]).astype(np.int64)
        batches.append((counts, dis, ids))
        total += len(dis)
        if total > max_results:
            radius, total = _tighten(batches, min_results, keep_max)
    if clip_to_min and total > min_results:
        radius, total = _tighten(batches, min_results, keep_max)
    counts = np.hstack([c for c, _, _ in batches]) if batches else np.zeros(0, np.int64)
    lims = np.zeros(len(counts) + 1, dtype=np.uint64)
    lims[1:] = np.cumsum(counts)
    D = np.hstack([d for _, d, _ in batches]) if batches else np.zeros(0, np.float32)
    I = np.hstack([i for _, _, i in batches]) if batches else np.zeros(0, np.int64)
    return radius, lims, D, I
