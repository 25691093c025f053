"""Per-stage device time of the TN pipeline at several batch sizes (dev tool): shows which stages are
latency-bound (time flat in the pair count) and which are throughput-bound."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from vsc2022_b200 import _lib, vta, workloads  # noqa: E402

L = 300
dev = torch.device("cuda", 0)
lib = _lib.load()
model = vta.build_vta_model("TN", tn_max_step=5, min_length=4)
sizes = [int(x) for x in sys.argv[1:]] or [148, 592, 1184, 2368, 4736, 8000, 16000]
for n_pairs in sizes:
    w = workloads.tn_pairs_device(n_pairs, L, L, seed=4, device=dev)
    for _ in range(3):
        model.align_device(w.sims, w.off, w.lq, w.lr, n_pairs, L, L, want_maxsim=False)
    torch.cuda.synchronize()
    lib.vsc_tn_set_profiling(1)
    acc = [0.0] * 4
    reps = 5
    for _ in range(reps):
        model.align_device(w.sims, w.off, w.lq, w.lr, n_pairs, L, L, want_maxsim=False)
        torch.cuda.synchronize()
        st = (ctypes.c_float * 4)()
        lib.vsc_tn_last_stage_ms(st)
        acc = [a + s for a, s in zip(acc, st)]
    lib.vsc_tn_set_profiling(0)
    t1, te, t2, _ = [a / reps for a in acc]
    print(f"pairs={n_pairs:6d}  topk {t1:.3f} ms  edges {te:.3f} ms  dp {t2:.3f} ms  sum {t1 + te + t2:.3f} ms  "
          f"-> {n_pairs * L * L * 4 / (t1 + te + t2) / 1e6:.0f} GB/s", flush=True)
    c = (ctypes.c_ulonglong * 8)()
    lib.vsc_tn_debug_counters(c)
    if c[7]:
        n = c[7]
        names = ["first sweep", "end search", "walk", "zero+score", "box filter", "incremental"]
        print("    dp phase kcycles per warp: " + ", ".join(f"{nm} {c[i] / n / 1e3:.1f}" for i, nm in enumerate(names)) +
              f"; incremental layer steps per warp-call {c[6] / n:.0f}", flush=True)
    del w
