"""numpy restatement of the slice of FAISS the vsc2022 reference touches.

TEST INFRASTRUCTURE (oracle).  FAISS is a third-party dependency of the
reference that is NOT vendored under /root/reference and cannot be installed in
this image (no wheel, no network); the reference does not pin a version
(docs/installation.md:9-11 installs conda ``faiss-gpu``); the contrib function
used below exists with inner-product support since faiss 1.7.3.

Call sites this shim serves
    vsc/index.py:82          faiss.index_factory(dim, "Flat", metric)
    vsc/index.py:94          index.add(x)
    vsc/index.py:145         index.metric_type == faiss.METRIC_INNER_PRODUCT
    vsc/index.py:147-154     faiss.contrib.exhaustive_search.range_search_max_results
    vsc/index.py:169-174     faiss.get_num_gpus(), index.search(x, k)
    vsc/baseline/score_normalization.py:87-96   index.search(x, 1)
    tests/test_index.py:43   faiss.METRIC_L2

Semantics restated (published FAISS behaviour for IndexFlat):
    * METRIC_INNER_PRODUCT scores are x @ y.T, larger is better;
      METRIC_L2 scores are SQUARED euclidean distances, smaller is better.
    * search(x, k) returns the k best per row, best first; missing slots
      (k > ntotal) are filled with id -1 and -inf / +inf.
    * range_search(x, radius) returns, per query and in database order, every
      entry with score > radius (IP) or distance < radius (L2) -- STRICT.
All arithmetic is float32 (numpy sgemm), like FAISS's BLAS path.
"""
import numpy as np

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


def get_num_gpus():
    return 0


class IndexFlat:
    def __init__(self, d, metric=METRIC_L2):
        self.d = int(d)
        self.metric_type = int(metric)
        self._chunks = []
        self._xb = np.zeros((0, self.d), dtype=np.float32)

    @property
    def ntotal(self):
        return self._xb.shape[0]

    def add(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        self._xb = np.concatenate([self._xb, x], axis=0)

    def _scores(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        if self.metric_type != METRIC_INNER_PRODUCT and x.shape[0] < 20:
            # FAISS computes small L2 batches (nq < distance_compute_blas_threshold
            # = 20) directly as sum((x - y)^2) rather than through the norm expansion.
            diff = x[:, None, :] - self._xb[None, :, :]
            return (diff * diff).sum(axis=2, dtype=np.float32)
        ip = x @ self._xb.T
        if self.metric_type == METRIC_INNER_PRODUCT:
            return ip
        qn = (x * x).sum(axis=1, dtype=np.float32)[:, None]
        bn = (self._xb * self._xb).sum(axis=1, dtype=np.float32)[None, :]
        return (qn + bn - np.float32(2.0) * ip).astype(np.float32)

    def search(self, x, k):
        s = self._scores(x)
        nq, nb = s.shape
        keep_max = self.metric_type == METRIC_INNER_PRODUCT
        kk = min(k, nb)
        order = np.argsort(-s if keep_max else s, axis=1, kind="stable")[:, :kk]
        D = np.full((nq, k), -np.inf if keep_max else np.inf, dtype=np.float32)
        I = np.full((nq, k), -1, dtype=np.int64)
        D[:, :kk] = np.take_along_axis(s, order, axis=1)
        I[:, :kk] = order
        return D, I

    def range_search(self, x, radius):
        s = self._scores(x)
        hit = s > radius if self.metric_type == METRIC_INNER_PRODUCT else s < radius
        lims = np.zeros(s.shape[0] + 1, dtype=np.uint64)
        lims[1:] = np.cumsum(hit.sum(axis=1))
        rows, cols = np.nonzero(hit)  # row-major: per query, database order
        return lims, s[rows, cols].astype(np.float32), cols.astype(np.int64)


IndexFlatIP = lambda d: IndexFlat(d, METRIC_INNER_PRODUCT)  # noqa: E731
IndexFlatL2 = lambda d: IndexFlat(d, METRIC_L2)  # noqa: E731


def index_factory(d, description, metric=METRIC_L2):
    if description != "Flat":
        raise NotImplementedError(f"oracle shim only restates 'Flat', got {description!r}")
    return IndexFlat(d, metric)


def index_cpu_to_all_gpus(index, co=None, ngpu=-1):
    return index
