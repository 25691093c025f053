"""One device-scheduled global top-K search at configs[2] size (for an ncu launch list): Gaussian float32 by default,
`grid` = descriptors rounded to 11 bits (single product)."""
import sys

import torch

sys.path.insert(0, ".")
from vsc2022_b200.index import VideoIndex  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "gauss"
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(3)
nqv, nrv, frames, d = 1250, 6250, 32, 512
q = torch.nn.functional.normalize(torch.randn((nqv * frames, d), generator=g, device=dev), dim=1)
r = torch.nn.functional.normalize(torch.randn((nrv * frames, d), generator=g, device=dev), dim=1)
if kind == "grid":
    q, r = q.half().float(), r.half().float()
index = VideoIndex(d)
index.index.add_device(r, copy=False)
K = 1200 * nqv
for _ in range(2):
    out = index.global_topk_device(q, K)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = index.global_topk_device(q, K); e1.record(); torch.cuda.synchronize()
print(kind, "search ms", e0.elapsed_time(e1), "pairs", out[0].numel())
