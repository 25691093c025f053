"""The rule tn_dp_kernel uses to order two nodes of one Kahn generation (csrc/tn_pipeline.cu: topo_compare_exact and the
resolve pass) against networkx's own topological order, on the oracle's graphs (CPU only, no engine involved).

Rule: generation 0 is in node-id order; a node of generation g >= 1 is placed by (position of its LAST parent -- the
parent of generation g-1 that comes last in that generation -- , node id).  Ancestors of a generation-g node in row q
lie in rows [q - (step-1)*g, q): the kernel only fills last parents for that row range.
"""
import networkx as nx
import numpy as np

from oracle import synth, tn_networkx


def _last_parent_order(graph, top, step):
    gens = list(nx.topological_generations(graph))
    gen_of = {v: g for g, nodes in enumerate(gens) for v in nodes}
    last = {}

    def compare(a, b, g):
        while True:
            if a == b:
                return 0
            if g <= 0:
                return -1 if a < b else 1
            pa, pb = last[a], last[b]
            if pa == pb:
                return -1 if a < b else 1
            a, b, g = pa, pb, g - 1

    for v in sorted(graph.nodes):           # ascending id: every ancestor is done first
        g = gen_of[v]
        best = None
        for u in graph.pred[v]:              # insertion order = ascending predecessor id (slot order in the kernel)
            if gen_of[u] != g - 1:
                continue
            if best is None or compare(u, best, g - 1) > 0:
                best = u
        if best is not None:
            last[v] = best
            row = lambda n: (n - 1) // top
            assert row(v) - (step - 1) * g <= row(best) < row(v)
    return gens, compare


def test_last_parent_rule_reproduces_networkx_generation_order():
    import functools
    rng = np.random.default_rng(5)
    checked = 0
    for it in range(60):
        lq, lr = int(rng.integers(8, 70)), int(rng.integers(8, 70))
        s = synth.sim_matrix(rng, lq, lr, bias=0.5, quant=[4.0, 8.0, 0.0][it % 3])
        step, top = (5, 5) if it % 2 else (4, 3)
        graph, _, top = tn_networkx.build_graph(s, step, top, 0.2)
        gens, compare = _last_parent_order(graph, top, step)
        for g, nodes in enumerate(gens):
            mine = sorted(nodes, key=functools.cmp_to_key(lambda a, b: compare(a, b, g)))
            assert mine == list(nodes), (it, g)
            checked += len(nodes) > 1
    assert checked > 100
