"""Probe: error of the fp16-split tensor-core GEMM against float64, where it comes from (dev tool)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from vsc2022_b200 import gemm  # noqa: E402

rng = np.random.default_rng(3)
d = int(sys.argv[1]) if len(sys.argv) > 1 else 511
a = rng.normal(size=(257, d)).astype(np.float32); b = rng.normal(size=(513, d)).astype(np.float32)
a /= np.linalg.norm(a, axis=1, keepdims=True); b /= np.linalg.norm(b, axis=1, keepdims=True)
b[:40] = a[100:140]
ref = a.astype(np.float64) @ b.astype(np.float64).T
da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
oa, ob = gemm.prepare_pair(da, db)
def report(name, c):
    err = np.abs(c.astype(np.float64) - ref)
    ident = err[100:140, :40].diagonal()
    print(f"{name:28s} max {err.max():.3e}  median {np.median(err):.3e}  identical-rows max {ident.max():.3e} mean signed {np.mean((c.astype(np.float64) - ref)[100:140, :40].diagonal()):+.3e}")
report("fp16 split (3 products)", gemm.gemm_store(oa, ob).cpu().numpy())
report("fp16 hi only (1 product)", gemm.gemm_store(oa, ob, precise=False).cpu().numpy())
report("numpy float32 matmul", a @ b.T)
torch.backends.cuda.matmul.allow_tf32 = False
report("torch fp32 matmul (cuBLAS)", (da @ db.T).cpu().numpy())
# emulate: hi/lo exact split in float64 to isolate the accumulation error
inv = float(oa.inv_scale.item()) * float(ob.inv_scale.item())
pa, pb = oa.panel[:257].float().cpu().numpy().astype(np.float64), ob.panel[:513].float().cpu().numpy().astype(np.float64)
exact = (pa @ pb.T) * inv
report("panels multiplied in float64", exact.astype(np.float64))
