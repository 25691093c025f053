"""Literal CPU restatement of VCSL's temporal-network (TN) alignment on networkx.

TEST INFRASTRUCTURE (oracle) -- the authority for stage C.

Third-party algorithm: alipay/VCSL, file ``vcsl/vta.py`` (functions ``tn`` and
``build_vta_model``; helper ``iou`` in ``vcsl/utils.py``), pinned by the
reference at commit c39269d5c3a252a6fba63ccc924d2f791a5bfece
(/root/reference/.SUBMODULES.json, .gitmodules).  The submodule is MISSING from
/root/reference and there is no network, so this file restates the published
algorithm (Tan et al., "Scalable detection of partial near-duplicate videos by
visual-temporal consistency", ACM MM 2009, as re-implemented in VCSL) and is
anchored on the reference's call sites:
    vsc/baseline/localization.py:44-46   build_vta_model(model_type, **kwargs)
    vsc/baseline/localization.py:58      model.forward_sim([(key, sim), ...])
    vsc/baseline/localization.py:59-64   same length/order as input, key echoed
    vsc/baseline/sscd_baseline.py:118-135  tn_max_step=5, min_length=4, concurrency=16
and pinned by tests/test_localization.py:45-66 (existence / absence of matches).
Box coordinates are NOT pinned by any reference test: parity for them is
"pinned by this restatement" ("Oracle pinning").

Numerical contract (what runs in THIS image: numpy 2.x, networkx 3.6.1):
    * similarities stay float32 end to end; numpy-2 weak-scalar promotion keeps
      ``0 + np.float32`` and ``np.float32 + 0.0`` in float32, so longest-path
      distances and the path score are float32 sums taken in path order;
      ``score / ave_length > min_sim`` and ``sim >= min_sim`` compare in float32.
    * row top-k order: ``np.argsort(-sims)`` in VCSL uses numpy's default
      (non-stable, SIMD/CPU-dependent) sort, so the order of EXACTLY equal
      values is not defined by the reference itself.  The contract here is the
      stable order (equal values: lower reference index first); the golden
      generator checks that the default and stable orders agree on every
      fixture.
    * networkx semantics relied on (networkx/algorithms/dag.py, 3.6.1):
      topological order = Kahn generations (zero-in-degree nodes in insertion
      order, children in adjacency order); per node the FIRST maximal
      predecessor wins; the end node is the FIRST maximum in topological order.
    * "sink" node (``tn(..., sink=...)``): upstream VCSL has a "link sink node"
      step that could not be read here (source missing).  Contract = "none".
      The two possible readings are implemented as a switch so that their
      effect is measured, not argued (tests/test_oracle_cpu.py::
      test_vcsl_sink_node_variants):
        "dedicated"  an extra node (Lq, Lr) after all real nodes, zero-weight
                     edges from every node within tn_max_step of it, stripped
                     from the path.  With first-max tie-breaking it is never the
                     end node and never lies inside a path: boxes are identical
                     to "none" on every golden and random fixture.
        "last_node"  the LAST REAL node (last query row, lowest top-k rank)
                     plays the sink: its incoming edges weigh 0 and it is
                     stripped from paths.  This changes the boxes of 13 of the
                     186 test matrices (paths that end in the last row).
      Which reading the pinned VCSL commit implements stays UNVERIFIED.
"""
from typing import List, Sequence, Tuple

import networkx as nx
import numpy as np
from networkx.algorithms.dag import dag_longest_path


def inclusive_iou(box: np.ndarray, kept: np.ndarray) -> np.ndarray:
    """IoU of one box against kept boxes with +1 (inclusive index) extents."""
    if len(box) == 0 or len(kept) == 0:
        return np.array(0)
    top_left = np.maximum(box[:, None, :2], kept[:, :2])
    bottom_right = np.minimum(box[:, None, 2:], kept[:, 2:])
    extent = np.maximum(bottom_right - top_left + 1, 0)
    inter = extent[:, :, 0] * extent[:, :, 1]
    area_box = (box[:, 2] - box[:, 0] + 1) * (box[:, 3] - box[:, 1] + 1)
    area_kept = (kept[:, 2] - kept[:, 0] + 1) * (kept[:, 3] - kept[:, 1] + 1)
    return inter / (area_box[:, None] + area_kept - inter)


def row_topk(sims: np.ndarray, top: int) -> Tuple[np.ndarray, np.ndarray]:
    """Per-row indices/values of the ``top`` largest entries, best first (stable)."""
    order = np.argsort(-sims, axis=1, kind="stable")[:, :top]
    return order, np.take_along_axis(sims, order, axis=-1)


def build_graph(sims: np.ndarray, tn_max_step: int, tn_top_k: int, min_sim: float):
    """Nodes: 0 = source (-1,-1), then 1 + q*top + rank.  Edges per C1-C4."""
    n_q = sims.shape[0]
    top = min(tn_top_k, sims.shape[1])
    ref_of, sim_of = row_topk(sims, top)
    graph = nx.DiGraph()
    graph.add_node(0)
    for node in range(1, 1 + n_q * top):
        graph.add_node(node)
    for q_src in range(n_q):
        r_src = ref_of[q_src]
        linked = np.empty((0,), dtype=np.int64)  # refs already linked from q_src
        # C1: destination row within tn_max_step of the source row
        for q_dst in range(q_src + 1, min(n_q, q_src + tn_max_step)):
            r_dst = ref_of[q_dst]
            delta = r_dst[:, None] - r_src[None, :]  # [dst rank, src rank]
            c2 = (delta > 0) & (delta < tn_max_step)
            if len(linked):
                before = linked[None, None, :] < r_src[None, :, None]
                after = linked[None, None, :] > r_dst[:, None, None]
                c3 = np.all(before | after, axis=-1)
            else:
                c3 = np.ones(c2.shape, dtype=bool)
            s_dst = np.repeat(sim_of[q_dst].reshape(-1, 1), top, axis=1)
            c4 = s_dst >= min_sim
            dst_rank, src_rank = np.where(c2 & c3 & c4)  # row-major: dst rank, then src rank
            weights = s_dst[dst_rank, src_rank]
            linked = np.unique(np.concatenate([linked, r_dst[dst_rank]]))
            graph.add_edges_from(
                (q_src * top + a + 1, q_dst * top + b + 1, {"weight": w})
                for a, b, w in zip(src_rank, dst_rank, weights)
            )
    return graph, ref_of, top


def tn(sims: np.ndarray, tn_max_step: int = 10, tn_top_k: int = 5, max_path: int = 10,
       min_sim: float = 0.2, min_length: int = 5, max_iou: float = 0.3, sink: str = "none",
       **_ignored) -> List[List[int]]:
    """Temporal-network alignment of one (Lq x Lr) similarity matrix.

    Returns up to ``max_path + 1`` boxes ``[q_min, r_min, q_max, r_max]``
    (frame INDICES, inclusive).  ``sink``: see the module docstring.
    """
    graph, ref_of, top = build_graph(sims, tn_max_step, tn_top_k, min_sim)
    n_real = sims.shape[0] * top
    sink_node = None
    if sink != "none" and n_real > 0:
        def coords(node):
            if node == 0:
                return (-1, -1)
            if node == n_real + 1:
                return (sims.shape[0], sims.shape[1])
            return ((node - 1) // top, int(ref_of[(node - 1) // top][(node - 1) % top]))
        if sink == "dedicated":
            sink_node = n_real + 1
            graph.add_node(sink_node)
        elif sink == "last_node":
            sink_node = n_real
        else:
            raise ValueError(f"unknown sink variant {sink!r}")
        jq, jr = coords(sink_node)
        for i in range(0, sink_node):
            iq, ir = coords(i)
            if jq > iq and jr > ir and jq - iq <= tn_max_step and jr - ir <= tn_max_step:
                graph.add_edge(i, sink_node, weight=0)
    boxes: List[List[int]] = []
    sweep = 0
    while sweep <= max_path:
        chain = dag_longest_path(graph)
        for u, v in zip(chain[:-1], chain[1:]):
            graph.add_edge(u, v, weight=0.0)  # spent: later sweeps gain nothing here
        chain = [n for n in chain if n != 0 and n != sink_node]
        if not chain:
            break
        qs = [(n - 1) // top for n in chain]
        rs = [int(ref_of[(n - 1) // top][(n - 1) % top]) for n in chain]
        score = 0.0
        for q, r in zip(qs, rs):
            score += sims[q][r]
        if score > 0:
            q_lo, q_hi, r_lo, r_hi = min(qs), max(qs), min(rs), max(rs)
        else:
            q_lo = q_hi = r_lo = r_hi = 0
        mean_extent = (r_hi - r_lo + q_hi - q_lo) / 2
        overlap = inclusive_iou(np.array([[q_lo, r_lo, q_hi, r_hi]]), np.array(boxes))
        if (mean_extent != 0 and score / mean_extent > min_sim
                and min(r_hi - r_lo, q_hi - q_lo) > min_length
                and overlap.max() < max_iou):
            boxes.append([int(q_lo), int(r_lo), int(q_hi), int(r_hi)])
        sweep += 1
    return boxes


def _tn_job(args):
    key, sims, cfg = args
    return key, tn(sims, **cfg)


class TemporalNetwork:
    """Object returned by ``build_vta_model("TN", ...)``."""

    def __init__(self, concurrency: int = 4, **cfg):
        self.concurrency = int(concurrency)
        self.cfg = cfg

    def forward_sim(self, data: Sequence[Tuple[str, np.ndarray]]):
        jobs = [(key, sims, self.cfg) for key, sims in data]
        if self.concurrency <= 1 or len(jobs) <= 1:
            return [_tn_job(j) for j in jobs]
        import multiprocessing as mp
        import os
        with mp.get_context("fork").Pool(min(self.concurrency, os.cpu_count() or 1)) as pool:
            return pool.map(_tn_job, jobs)


def build_vta_model(method: str = "DTW", concurrency: int = 4, **config):
    if method != "TN":
        raise NotImplementedError(
            f"oracle restates only the 'TN' aligner used by vsc2022 (got {method!r})")
    return TemporalNetwork(concurrency, **config)
