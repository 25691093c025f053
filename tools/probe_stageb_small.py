"""Probe: where the host time of score_normalize + search goes at the small pipeline size (dev tool)."""
import cProfile
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from vsc2022_b200 import sscd_baseline  # noqa: E402
from vsc2022_b200.index import VideoFeature  # noqa: E402
from vsc2022_b200.score_normalization import score_normalize  # noqa: E402

nq, nr, nn, fr, d = 50, 400, 50, 40, 512
rng = np.random.default_rng(0)
ts = np.stack([np.arange(fr) * 1.0, np.arange(fr) * 1.0 + 1.0], axis=1)
mk = lambda p, n: [VideoFeature(video_id=f"{p}{i:06d}", timestamps=ts, feature=rng.normal(size=(fr, d)).astype(np.float32)) for i in range(n)]
q, r, noise = mk("Q", nq), mk("R", nr), mk("N", nn)
for i in range(0, nq, 2):
    q[i].feature[8:28] = r[(i * 37) % nr].feature[4:24]


def run():
    sn_q, sn_r = score_normalize(q, r, noise, beta=1.2)
    cands = sscd_baseline.search(sn_q, sn_r)
    matches = sscd_baseline.localize_and_verify(sn_q, sn_r, cands, score_normalization=True)
    torch.cuda.synchronize()
    return len(cands), len(matches)


for _ in range(2):
    print(run())
t0 = time.perf_counter(); run(); print("wall", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
