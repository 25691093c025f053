"""Make the engine importable under the reference's module names (drop-in use).

    import vsc2022_b200.compat; vsc2022_b200.compat.install()
    from vsc.index import VideoIndex            # -> vsc2022_b200.index
    from vcsl.vta import build_vta_model        # -> vsc2022_b200.vta
    import faiss; faiss.METRIC_INNER_PRODUCT    # constants + index_factory only

Nothing is installed unless `install()` is called, and existing real packages are left alone unless force=True.
"""
import importlib
import sys
import types

_ALIASES = {
    "vsc.index": "vsc2022_b200.index",
    "vsc.candidates": "vsc2022_b200.candidates",
    "vsc.metrics": "vsc2022_b200.metrics",
    "vsc.storage": "vsc2022_b200.storage",
    "vsc.baseline.score_normalization": "vsc2022_b200.score_normalization",
    "vsc.baseline.localization": "vsc2022_b200.localization",
    "vsc.descriptor_eval_lib": "vsc2022_b200.descriptor_eval_lib",
    "vsc.baseline.sscd_baseline": "vsc2022_b200.sscd_baseline",
    "vcsl.vta": "vsc2022_b200.vta",
}


def install(force: bool = False):
    def package(name):
        if name not in sys.modules or force:
            mod = types.ModuleType(name)
            mod.__path__ = []
            sys.modules[name] = mod
        return sys.modules[name]

    for alias, target in _ALIASES.items():
        parts = alias.split(".")
        for i in range(1, len(parts)):
            package(".".join(parts[:i]))
        try:
            mod = importlib.import_module(target)
        except ModuleNotFoundError:
            continue
        sys.modules[alias] = mod
        setattr(sys.modules[".".join(parts[:-1])], parts[-1], mod)
    if "faiss" not in sys.modules or force:
        faiss = faiss_module()
        sys.modules["faiss"] = faiss
        sys.modules["faiss.contrib"] = faiss.contrib
        sys.modules["faiss.contrib.exhaustive_search"] = faiss.contrib.exhaustive_search


class _ExponentialQueries:
    """What `exponential_query_iterator(xq)` returns: iterable over the slices FAISS would yield (32, 64, ... rows), and
    it remembers the whole query matrix so that range_search_max_results can hand the schedule to the engine in one call."""

    def __init__(self, xq, start_bs=32, max_bs=20000):
        self.xq, self.start_bs, self.max_bs = xq, start_bs, max_bs

    def __iter__(self):
        from .index import exponential_batches
        for lo, hi in exponential_batches(len(self.xq), self.start_bs, self.max_bs):
            yield self.xq[lo:hi]


def _range_search_max_results(index, query_iterator, radius, max_results=None, min_results=None, shard=False, ngpu=0,
                              clip_to_min=False):
    """faiss.contrib.exhaustive_search.range_search_max_results over the GPU engine (the call of vsc/index.py:147-154).
    Returns (radius, lims, D, I) like FAISS: results grouped per query row, in database order inside a row."""
    import numpy as np
    from . import _lib
    from .index import METRIC_INNER_PRODUCT, exponential_batches
    torch = _lib.require_cuda()
    if max_results is None and min_results is None:
        raise ValueError("range_search_max_results: max_results and/or min_results required")
    if min_results is None:
        min_results = int(0.8 * max_results)
    if max_results is None:
        max_results = int(min_results * 1.5)
    if clip_to_min or shard:
        raise NotImplementedError("clip_to_min / shard are not on the vsc2022 path")
    if isinstance(query_iterator, _ExponentialQueries):
        xq, start_bs, max_bs = query_iterator.xq, query_iterator.start_bs, query_iterator.max_bs
    else:   # any iterable of query blocks: the radius schedule depends on the block sizes, only FAISS's own is built in
        blocks = [np.asarray(b) for b in query_iterator]
        xq, start_bs, max_bs = (np.concatenate(blocks) if blocks else np.zeros((0, index.d), np.float32)), 32, 20000
        if [len(b) for b in blocks] != [hi - lo for lo, hi in exponential_batches(len(xq))]:
            raise NotImplementedError("only the batch sizes of exponential_query_iterator (32, 64, ...) are supported")
    if (start_bs, max_bs) != (32, 20000):
        raise NotImplementedError("only exponential_query_iterator's default batch sizes are supported")
    keep_max = index.metric_type == METRIC_INNER_PRODUCT
    if radius != (-1e10 if keep_max else 1e10):
        raise NotImplementedError("the engine starts from the unbounded radius vsc uses (-1e10 / 1e10)")
    score, row, col, final = index.range_search_max_results(xq, int(max_results), int(min_results))
    nq, nb = len(xq), max(index.ntotal, 1)
    order = torch.argsort(row * nb + col)
    row, col, score = row[order].cpu().numpy(), col[order].cpu().numpy(), score[order].cpu().numpy()
    lims = np.zeros(nq + 1, dtype=np.int64)
    lims[1:] = np.cumsum(np.bincount(row, minlength=nq))
    return final, lims, score, col


def faiss_module():
    """A module object with the part of the faiss API the reference touches (vsc/index.py:11-13,82,94,145-154,169-174;
    score_normalization.py:87-96; tests/test_index.py:8,43), backed by the GPU engine."""
    from . import index
    faiss = types.ModuleType("faiss")
    faiss.METRIC_INNER_PRODUCT, faiss.METRIC_L2 = index.METRIC_INNER_PRODUCT, index.METRIC_L2
    faiss.index_factory = index.index_factory
    faiss.IndexFlat = lambda d, metric=index.METRIC_L2: index.FlatIndex(d, metric)
    faiss.IndexFlatIP = lambda d: index.FlatIndex(d, index.METRIC_INNER_PRODUCT)
    faiss.IndexFlatL2 = lambda d: index.FlatIndex(d, index.METRIC_L2)
    faiss.get_num_gpus = lambda: 0   # callers then use the index they were given, which already runs on the GPU
    faiss.index_cpu_to_all_gpus = lambda idx, *a, **k: idx
    contrib = types.ModuleType("faiss.contrib")
    contrib.__path__ = []
    es = types.ModuleType("faiss.contrib.exhaustive_search")
    es.exponential_query_iterator = lambda xq, start_bs=32, max_bs=20000: _ExponentialQueries(xq, start_bs, max_bs)
    es.range_search_max_results = _range_search_max_results
    contrib.exhaustive_search = es
    faiss.contrib = contrib
    faiss.__path__ = []
    return faiss
