"""numpy restatement of the reference's stage-C glue (TEST INFRASTRUCTURE).

Restated reference code
    vsc/baseline/localization.py:33-36,49-54  similarity = Q @ R.T (+ bias)
    vsc/baseline/localization.py:56-79        localize_all: boxes -> timestamps
    vsc/baseline/localization.py:88-96        MaxSim / CandidateScore scoring
    vsc/index.py:26-30                        get_timestamps (N or Nx2)
The aligner itself is oracle/tn_networkx.py (or tn_fast for big workloads).
"""
from typing import Callable, List, Sequence, Tuple

import numpy as np

from oracle import tn_networkx


def stamp(timestamps: np.ndarray, idx: int) -> Tuple[float, float]:
    t = timestamps[idx]
    if timestamps.ndim == 1:
        return (t, t)
    return (t[0], t[1])


def similarity(q_feat: np.ndarray, r_feat: np.ndarray, bias: float = 0.0) -> np.ndarray:
    return np.matmul(q_feat, r_feat.T) + bias


def localize_all(pairs: Sequence[Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, float]],
                 similarity_bias: float, scoring: str, tn_fn: Callable = None, **tn_cfg):
    """pairs: (q_feat, q_timestamps, r_feat, r_timestamps, candidate_score) per candidate.

    Returns per candidate a list of (query_start, query_end, ref_start, ref_end, score, box).
    scoring: "max_sim" (localization.py:88-91) or "candidate" (:94-96).
    """
    tn_fn = tn_fn or tn_networkx.tn
    out: List[List[tuple]] = []
    for q_feat, q_ts, r_feat, r_ts, cand_score in pairs:
        sim = similarity(q_feat, r_feat, similarity_bias)
        rows = []
        for box in tn_fn(sim, **tn_cfg):
            x1, y1, x2, y2 = box
            if scoring == "max_sim":
                # NOTE the slice is EXCLUSIVE of row x2 / column y2 (reference quirk)
                score = sim[x1:x2, y1:y2].max() - similarity_bias
            else:
                score = cand_score
            rows.append((stamp(q_ts, x1)[0], stamp(q_ts, x2)[1],
                         stamp(r_ts, y1)[0], stamp(r_ts, y2)[1], score, tuple(box)))
        out.append(rows)
    return out
