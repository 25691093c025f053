"""The compact-graph formulation of the temporal network (csrc/tn_graph.cu), pinned on the CPU.

oracle/tn_graph_model.c executes the kernels' algorithm -- bitmap edge screening, compaction to active nodes + CSR edges,
longest paths by Kahn generation, generation-ordered re-relaxation after every extracted chain -- sequentially; here it
must reproduce oracle/tn_fast.c (itself pinned against the networkx restatement) wherever it does not hand the pair
back (table overflow, ties that generations cannot break).
"""
import numpy as np
import pytest

from oracle import synth, tn_fast, tn_graph_model, tn_networkx

CFGS = [
    dict(tn_max_step=5, tn_top_k=5, min_length=4),                                       # vsc2022 parameters
    dict(tn_max_step=3, tn_top_k=3, min_length=2),
    dict(tn_max_step=7, tn_top_k=5, min_length=3, max_path=4, min_sim=0.45, max_iou=0.1),
    dict(tn_max_step=5, tn_top_k=8, min_length=4),
    dict(tn_max_step=9, tn_top_k=4, min_length=5, max_path=2),
]


def test_random_matrices_match_tn_fast():
    rng = np.random.default_rng(21)
    finished = handed_back = 0
    for it in range(1200):
        lq, lr = int(rng.integers(1, 110)), int(rng.integers(1, 110))
        quant = [0.0, 0.0, 8.0, 64.0][it % 4]          # coarse grids force exact ties
        s = synth.sim_matrix(rng, lq, lr, bias=[0.5, 0.0][it % 2], quant=quant)
        cfg = CFGS[it % len(CFGS)]
        boxes, status = tn_graph_model.tn(s, **cfg)
        if status != tn_graph_model.STATUS_OK:
            handed_back += 1
            continue
        finished += 1
        assert boxes == tn_fast.tn(s, **cfg), (it, lq, lr, quant, cfg)
    assert finished > 500, (finished, handed_back)


def test_full_size_pairs_and_networkx():
    rng = np.random.default_rng(22)
    cfg = dict(tn_max_step=5, tn_top_k=5, min_length=4)
    done = 0
    for it in range(24):
        s = synth.sim_matrix(rng, 300, 300, dim=512 if it % 2 else 64)
        boxes, status = tn_graph_model.tn(s, **cfg)
        if status == tn_graph_model.STATUS_OK:
            done += 1
            assert boxes == tn_fast.tn(s, **cfg)
            if it < 4:
                assert boxes == tn_networkx.tn(s, **cfg)
    assert done >= 20


def test_edge_shapes():
    rng = np.random.default_rng(23)
    cfg = dict(tn_max_step=5, tn_top_k=5, min_length=4)
    for s in (synth.sim_matrix(rng, 1, 1), synth.sim_matrix(rng, 2, 2), np.zeros((9, 9), np.float32),
              synth.sim_matrix(rng, 12, 400), synth.sim_matrix(rng, 200, 3), np.full((30, 30), 0.75, np.float32)):
        boxes, status = tn_graph_model.tn(s, **cfg)
        if status == tn_graph_model.STATUS_OK:
            assert boxes == tn_fast.tn(s, **cfg), s.shape


def test_wide_parameter_sets_are_rejected():
    with pytest.raises(ValueError):
        tn_graph_model.tn(np.zeros((8, 8), np.float32))     # VCSL defaults: 9 * 5 = 45 predecessor slots
