"""Probe: where the time of the reference-facing localize_all goes at configs[3] size (host descriptors in, Match rows out).

    python tools/probe_e2e.py [n_pairs]
Wall-clock phases with a device synchronisation after each (so the sum exceeds the un-instrumented call, which overlaps
the upload with host work), then the un-instrumented call.
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from vsc2022_b200 import gemm  # noqa: E402
from vsc2022_b200.index import VideoFeature  # noqa: E402
from vsc2022_b200.localization import VCSLLocalizationMaxSim  # noqa: E402
from vsc2022_b200.metrics import CandidatePair  # noqa: E402
from vsc2022_b200.workloads import C4Workload  # noqa: E402

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
import os  # noqa: E402
from vsc2022_b200 import localization as _loc  # noqa: E402
if os.environ.get("VSC_E2E_BLOCK"):     # sweep of the upload granularity
    _loc._DeviceVideos.BLOCK = int(os.environ["VSC_E2E_BLOCK"])
if os.environ.get("VSC_E2E_BRIDGE"):
    _loc._DeviceVideos.BRIDGE = int(os.environ["VSC_E2E_BRIDGE"])
if os.environ.get("VSC_E2E_CHUNK"):
    _loc.VCSLLocalization.CHUNK = int(os.environ["VSC_E2E_CHUNK"])
print("BLOCK", _loc._DeviceVideos.BLOCK, "BRIDGE", _loc._DeviceVideos.BRIDGE, "CHUNK", _loc.VCSLLocalization.CHUNK)
f, dim = 300, 512
wl = C4Workload(n_pairs, frames=f, dim=dim)
q_ids = sorted(set(wl.pair_query.tolist()))
r_ids = sorted(set(wl.pair_ref.tolist()))
q_host = torch.empty((len(q_ids) * f, dim), dtype=torch.float32, pin_memory=True)
r_host = torch.empty((len(r_ids) * f, dim), dtype=torch.float32, pin_memory=True)
wl.videos_for(0, n_pairs, q_base=q_host.numpy(), r_base=r_host.numpy())
ts_q = np.tile(np.arange(f, dtype=np.float64), len(q_ids)).copy()
ts_r = np.tile(np.arange(f, dtype=np.float64), len(r_ids)).copy()
qh, rh = q_host.numpy(), r_host.numpy()
queries = [VideoFeature(video_id=f"Q{q:06d}", timestamps=ts_q[i * f:(i + 1) * f], feature=qh[i * f:(i + 1) * f]) for i, q in enumerate(q_ids)]
refs = [VideoFeature(video_id=f"R{r:06d}", timestamps=ts_r[i * f:(i + 1) * f], feature=rh[i * f:(i + 1) * f]) for i, r in enumerate(r_ids)]
cands = [CandidatePair(f"Q{int(wl.pair_query[p]):06d}", f"R{int(wl.pair_ref[p]):06d}", 1.0) for p in range(n_pairs)]
sync = torch.cuda.synchronize


def whole():
    loc = VCSLLocalizationMaxSim(queries, refs, "TN", tn_max_step=5, min_length=4, concurrency=16, similarity_bias=0.5)
    return loc.localize_all(cands)


for _ in range(2):
    whole()
sync()
for rep in range(0 if os.environ.get('VSC_E2E_QUICK') else 2):
    t = [time.perf_counter()]
    loc = VCSLLocalizationMaxSim(queries, refs, "TN", tn_max_step=5, min_length=4, concurrency=16, similarity_bias=0.5)
    t.append(time.perf_counter())
    dq, dr = loc._stores()
    qi = [c.query_id for c in cands]
    ri = [c.ref_id for c in cands]
    dq.ensure(qi); dr.ensure(ri)
    t.append(time.perf_counter())
    sync()
    t.append(time.perf_counter())
    loc._operands()
    sync()
    t.append(time.perf_counter())
    m = loc.localize_all(cands)
    sync()
    t.append(time.perf_counter())
    names = ["construct (dicts)", "ensure (host walk, async upload issued)", "upload completes", "prepare panels", "rest of localize_all (meta, GEMM+TN, D2H, Match rows)"]
    print(f"--- instrumented pass {rep}: {len(m)} matches")
    for nme, a, b in zip(names, t[:-1], t[1:]):
        print(f"  {nme:60s} {1e3 * (b - a):8.2f} ms")
    print(f"  {'sum':60s} {1e3 * (t[-1] - t[0]):8.2f} ms   h2d {(dq.h2d_bytes + dr.h2d_bytes) / 1e9:.2f} GB")
for _ in range(2):
    loc = VCSLLocalizationMaxSim(queries, refs, "TN", tn_max_step=5, min_length=4, concurrency=16, similarity_bias=0.5)
    loc.profile = {}
    sync(); t0 = time.perf_counter(); loc.localize_all(cands); sync(); dt = time.perf_counter() - t0
    print(f"host phases of one chunked call ({1e3 * dt:.1f} ms):", {k: (round(1e3 * v, 2) if isinstance(v, float) and k != "t_start" else v) for k, v in loc.profile.items() if k != "t_start"})
ts = []
for _ in range(5):
    sync(); t0 = time.perf_counter(); whole(); sync(); ts.append(time.perf_counter() - t0)
print("un-instrumented localize_all, fresh object: ms", [round(1e3 * x, 2) for x in ts])
# raw copy rate of the same bytes
sync(); t0 = time.perf_counter(); a = q_host.to("cuda", non_blocking=True); b = r_host.to("cuda", non_blocking=True); sync()
dt = time.perf_counter() - t0
print(f"raw pinned H2D of the same arrays: {1e3 * dt:.2f} ms = {(q_host.numel() + r_host.numel()) * 4 / dt / 1e9:.1f} GB/s")
if os.environ.get('VSC_E2E_QUICK'):
    sys.exit(0)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); whole(); sync(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
