"""Probe: stage C from descriptors at configs[3] size (8000 pairs of 300 x 300, 512-d), device resident.

    python tools/probe_pairs.py [n_pairs] [frames] [dim] [grid|gauss]
Prints CUDA-event times of vcsl_tn_batch_from_features (with / without MaxSim), of vsc_pair_similarity + vcsl_tn_batch,
and the per-stage times of the fast pipeline.
"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from vsc2022_b200 import _lib, gemm, vta  # noqa: E402

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 512
kind = sys.argv[4] if len(sys.argv) > 4 else "gauss"
dev = torch.device("cuda")
n_q = max(1, n_pairs // 5)
n_r = n_q
gen = torch.Generator(device=dev)
gen.manual_seed(3)
Q = torch.nn.functional.normalize(torch.randn((n_q * L, dim), generator=gen, device=dev), dim=1)
R = torch.nn.functional.normalize(torch.randn((n_r * L, dim), generator=gen, device=dev), dim=1)
rng = np.random.default_rng(0)
qi = np.repeat(np.arange(n_q), 5)[:n_pairs]
ri = rng.integers(0, n_r, size=n_pairs)
for p in range(0, n_pairs, 2):   # planted copies: 20-80 frames of the reference inside the query
    n = int(rng.integers(20, 81)); a = int(rng.integers(0, L - n + 1)); b = int(rng.integers(0, L - n + 1))
    Q[qi[p] * L + a: qi[p] * L + a + n] = torch.nn.functional.normalize(
        R[ri[p] * L + b: ri[p] * L + b + n] + 0.1 / dim ** 0.5 * torch.randn((n, dim), generator=gen, device=dev), dim=1)
if kind == "grid":
    Q, R = Q.bfloat16().float(), R.bfloat16().float()
oq, orr = gemm.prepare_pair(Q, R)
pairing = gemm.Pairing(oq, orr)
print(f"pairs {n_pairs}  frames {L}  dim {dim}  split {pairing.split}  K' {pairing.k}")
meta = np.stack([qi * L, np.full(n_pairs, L), ri * L, np.full(n_pairs, L)]).astype(np.int32)
d_meta = torch.from_numpy(meta).to(dev)
params = vta.tn_params(tn_max_step=5, min_length=4)
lib = _lib.load()
lib.vsc_tn_set_profiling(1)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), r


def stages():
    out = (ctypes.c_float * 4)()
    lib.vsc_tn_last_stage_ms(out)
    return [round(x, 3) for x in out]


direct = lambda ms: vta.tn_batch_from_features(oq.panel, orr.panel, pairing.k, d_meta[0], d_meta[1], d_meta[2], d_meta[3], n_pairs,
                                               L, L, L, 0.5, params, want_maxsim=ms, fmt=pairing)
t, res = timed(lambda: direct(False))
print(f"from_features (boxes only)      {t:.3f} ms  {n_pairs / t / 1e3:.2f} M pairs/s  stages[topk, edges, dp, maxsim] {stages()}")
boxes, nb, _, st = res.to_host()
print("  boxes", int(nb.sum()), "status", np.bincount(st, minlength=3).tolist())
t, res2 = timed(lambda: direct(True))
print(f"from_features (+MaxSim scores)  {t:.3f} ms  {n_pairs / t / 1e3:.2f} M pairs/s  stages {stages()}")
sims = torch.empty((n_pairs * L * L + 4,), dtype=torch.float32, device=dev)
off = torch.arange(n_pairs, device=dev, dtype=torch.int64) * (L * L)
t, _ = timed(lambda: vta.pair_similarity(oq.panel, orr.panel, pairing.k, d_meta[0], d_meta[1], d_meta[2], d_meta[3], n_pairs,
                                         L, L, 0.5, sims, off, fmt=pairing))
flops = 2.0 * n_pairs * L * L * pairing.k
print(f"pair_similarity (matrices out)  {t:.3f} ms  {flops / t / 1e9:.0f} TFLOP/s  {n_pairs * L * L * 4 / t / 1e6:.0f} GB/s written")
model = vta.TN(tn_max_step=5, min_length=4)
t, res3 = timed(lambda: model.align_device(sims, off, d_meta[1], d_meta[3], n_pairs, L, L, want_maxsim=False))
print(f"vcsl_tn_batch on those matrices {t:.3f} ms  stages {stages()}")
b3, nb3, _, _ = res3.to_host()
same = bool((nb3 == nb).all()) and all((b3[i, :nb[i]] == boxes[i, :nb[i]]).all() for i in range(n_pairs))
print("  boxes identical to the from-features path:", same)
