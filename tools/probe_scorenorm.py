"""Probe: phases of score_normalize_device at configs[2] size (40k query, 200k ref, 200k noise, 512-d Gaussian), CUDA events."""
import sys

import torch

sys.path.insert(0, ".")
from vsc2022_b200 import gemm  # noqa: E402
from vsc2022_b200.index import METRIC_INNER_PRODUCT, FlatIndex  # noqa: E402
from vsc2022_b200.score_normalization import l2norm_dropdim, lowvar_dim, score_normalize_device  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(3)
q = torch.randn((40_000, 512), generator=g, device=dev)
r = torch.randn((200_000, 512), generator=g, device=dev)
noise = torch.randn((200_000, 512), generator=g, device=dev)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


t, _ = timed(lambda: score_normalize_device(q, r, noise, True, True, 1.2)); print(f"score_normalize_device total {t:.2f} ms")
t, drop = timed(lambda: lowvar_dim(noise)); print(f"  lowvar_dim(noise)            {t:.2f} ms")
t, (qn, kept) = timed(lambda: l2norm_dropdim(q, drop, True, extra_column=True)); print(f"  l2norm_dropdim(q)            {t:.2f} ms")
t, _ = timed(lambda: l2norm_dropdim(r, drop, True, extra_column=True, tail=1.0)); print(f"  l2norm_dropdim(r)            {t:.2f} ms")
t, (nn, _) = timed(lambda: l2norm_dropdim(noise, drop, True, extra_column=False)); print(f"  l2norm_dropdim(noise)        {t:.2f} ms")
t, (oa, ob) = timed(lambda: gemm.prepare_pair(qn[:, :kept], nn)); print(f"  prepare_pair(q, noise)       {t:.2f} ms")
t, _ = timed(lambda: gemm.gemm_rowmax(oa, ob)); print(f"  gemm_rowmax split            {t:.2f} ms   (k = {gemm.Pairing(oa, ob).k})")
t, _ = timed(lambda: gemm.gemm_rowmax(oa, ob, precise=False)); print(f"  gemm_rowmax single pass      {t:.2f} ms")
idx = FlatIndex(kept, METRIC_INNER_PRODUCT); idx.add_device(nn, copy=False)
t, _ = timed(lambda: idx.max_similarity(qn[:, :kept])); print(f"  index.max_similarity         {t:.2f} ms")
