"""Timing of the matching-track metric on large prediction sets: this package's numpy interval
sweeps against the UNMODIFIED reference `vsc.metrics.match_metric` (imported from /root/reference over the oracle shims),
same inputs, same result.  Runs where the reference tree is mounted (the build container; CPU only).

    python tools/probe_metrics.py [n_pairs ...]
"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle.run_reference_tests import reference_on_path  # noqa: E402

reference_on_path()
import vsc.metrics as ref_metrics  # noqa: E402  (the reference's own file)
from vsc2022_b200 import metrics as our_metrics  # noqa: E402


def workload(n_pairs, rng):
    """n_pairs video pairs; ~1.5 ground-truth segments and ~9.5 predictions per pair (what the TN stage emits at
    configs[3]: 76 k predictions for 8000 pairs), predictions scattered around the truth with random scores."""
    gts, preds = [], []
    for p in range(n_pairs):
        q, r = f"Q{p // 5:06d}", f"R{(p * 7919) % (n_pairs // 5 + 1):06d}"
        for _ in range(int(rng.integers(1, 3))):
            a, b, n = rng.uniform(0, 240), rng.uniform(0, 240), rng.uniform(10, 60)
            gts.append(dict(query_id=q, ref_id=r, query_start=a, query_end=a + n, ref_start=b, ref_end=b + n, score=1.0))
            for _ in range(int(rng.integers(3, 8))):
                da, db, m = rng.normal(0, 8), rng.normal(0, 8), rng.uniform(5, 50)
                preds.append(dict(query_id=q, ref_id=r, query_start=max(0, a + da), query_end=max(0, a + da) + m,
                                  ref_start=max(0, b + db), ref_end=max(0, b + db) + m, score=float(rng.uniform(0, 1))))
    return gts, preds


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [200, 1000, 8000]
    rng = np.random.default_rng(0)
    print(f"{'pairs':>6s} {'gt':>7s} {'preds':>7s} {'reference s':>12s} {'this repo s':>12s} {'speed-up':>9s}  segment AP (ref / ours)")
    for n in sizes:
        gts, preds = workload(n, rng)
        t0 = time.perf_counter()
        want = ref_metrics.match_metric([ref_metrics.Match(**g) for g in gts], [ref_metrics.Match(**p) for p in preds])
        t1 = time.perf_counter()
        got = our_metrics.match_metric([our_metrics.Match(**g) for g in gts], [our_metrics.Match(**p) for p in preds])
        t2 = time.perf_counter()
        assert abs(want.ap - got.ap) < 1e-12, (want.ap, got.ap)
        print(f"{n:6d} {len(gts):7d} {len(preds):7d} {t1 - t0:12.2f} {t2 - t1:12.2f} {(t1 - t0) / (t2 - t1):8.1f}x  {want.ap:.6f} / {got.ap:.6f}")


if __name__ == "__main__":
    main()
