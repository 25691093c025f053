"""Quick device-side timing of the TN kernel on the C4 workload (dev tool, not the bench)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from vsc2022_b200 import _lib, vta, workloads  # noqa: E402

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = torch.device("cuda", 0)
w = workloads.tn_pairs_device(n_pairs, L, L, seed=4, device=dev)
model = vta.build_vta_model("TN", tn_max_step=5, min_length=4)
for want_maxsim in (False, True):
    for _ in range(3):
        res = model.align_device(w.sims, w.off, w.lq, w.lr, n_pairs, L, L, want_maxsim=want_maxsim)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(5):
        res = model.align_device(w.sims, w.off, w.lq, w.lr, n_pairs, L, L, want_maxsim=want_maxsim)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 5
    gbs = n_pairs * L * L * 4 / ms / 1e6
    print(f"maxsim={want_maxsim} pairs={n_pairs} {L}x{L}: {ms:.3f} ms/step, {n_pairs / ms * 1e3:.0f} pairs/s, {gbs:.0f} GB/s algorithmic")
bx, nb, ms_, st = res.to_host()
print("boxes/pair mean", nb.mean(), "exact-kernel pairs", int(st.sum()))
