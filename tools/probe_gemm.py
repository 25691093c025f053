"""Device-side timing of the descriptor GEMM epilogues at the C3 shape (dev tool)."""
import sys

import torch

sys.path.insert(0, ".")
from vsc2022_b200 import gemm  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
d = 512
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(3)
q = torch.randn((nq, d), generator=g, device=dev).bfloat16().float()
r = torch.randn((nr, d), generator=g, device=dev).bfloat16().float()
oa, ob = gemm.prepare_pair(q, r)
print("split:", gemm.Pairing(oa, ob).split, "k:", gemm.Pairing(oa, ob).k)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


flops = 2.0 * nq * nr * oa.k
ms = timeit(lambda: gemm.gemm_rowmax(oa, ob))
print(f"rowmax {nq}x{nr}x{oa.k}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s")
hits = gemm.HitBuffer(8_000_000, dev)
thr = float(torch.quantile((q[:256] @ r[:20000].T).flatten(), 0.9999))
def emit():
    hits.counters.zero_()
    gemm.gemm_emit(oa, ob, hits, thr, thr)
ms = timeit(emit)
print(f"emit (thr={thr:.1f}) : {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s  counters {hits.read_counters()}")
ref = torch.matmul(q[:4096].bfloat16(), r[:8192].bfloat16().T)
ms = timeit(lambda: torch.matmul(q.bfloat16()[:32768], r.bfloat16()[:65536].T), 3)
print(f"cuBLAS bf16 32768x65536x512 (writes C): {ms:.3f} ms {2.0*32768*65536*512/ms/1e9:.1f} TFLOP/s")
