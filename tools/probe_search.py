"""Stage B at the C3 scale: 40k query x 200k ref x 512-d descriptors, score-norm + global top-K (dev tool)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from vsc2022_b200 import _lib, gemm  # noqa: E402
from vsc2022_b200.index import FlatIndex, METRIC_INNER_PRODUCT, VideoIndex  # noqa: E402

nqv, nrv, frames, d = 1250, 6250, 32, 512
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(3)
q = torch.randn((nqv * frames, d), generator=g, device=dev).bfloat16().float()
r = torch.randn((nrv * frames, d), generator=g, device=dev).bfloat16().float()
noise = torch.randn((200000, d), generator=g, device=dev).bfloat16().float()
# 5% of query videos carry a 16-frame copy of a random ref
for v in range(0, nqv, 20):
    rv = (v * 7919) % nrv
    q[v * frames + 8:v * frames + 24] = r[rv * frames + 4:rv * frames + 20]
K = 1200 * nqv


def sync_time(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n, out


index = VideoIndex(d)
index.index.add_device(r)
l0 = _lib.launch_count()
t, res = sync_time(lambda: index.global_topk_device(q, K))
print(f"global top-K search (K={K}): {t*1e3:.1f} ms  -> {(q.shape[0] + r.shape[0]) / t / 1e6:.2f} M descriptors/s; "
      f"{res[2].numel()} hits, launches/call {(_lib.launch_count() - l0) / 4:.0f}")
flops = 2.0 * q.shape[0] * r.shape[0] * d
print(f"   GEMM-equivalent {flops / t / 1e12:.0f} TFLOP/s over the whole search (single pass = {flops/1e12:.2f} TFLOP)")

nidx = FlatIndex(d, METRIC_INNER_PRODUCT); nidx.add_device(noise)
t, best = sync_time(lambda: nidx.max_similarity(q))
print(f"score-norm 1-NN vs 200k noise: {t*1e3:.1f} ms ({2.0*q.shape[0]*noise.shape[0]*d/t/1e12:.0f} TFLOP/s incl. operand prep)")

# split (fp32-class) path on L2-normalised features
qn = torch.nn.functional.normalize(q[:, :511]); rn = torch.nn.functional.normalize(r[:, :511])
idx2 = VideoIndex(511); idx2.index.add_device(rn)
t, res = sync_time(lambda: idx2.global_topk_device(qn, K), n=2)
print(f"global top-K on normalised fp32 descriptors (3-term split, K'={3*512}): {t*1e3:.1f} ms")
