"""Probe (dev tool): (1) latency of the exact-order kernel on a few pairs, (2) single-pass GEMM: strided vs contiguous hi panel."""
import ctypes, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from vsc2022_b200 import _lib, gemm, vta, workloads

dev = torch.device("cuda")
def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for n in (1, 8, 148):
    w = workloads.tn_pairs_device(n, 300, 300, seed=4, device=dev, dim=512)
    m = vta.build_vta_model("TN", tn_max_step=5, min_length=4)
    t0 = timed(lambda: m.align_device(w.sims, w.off, w.lq, w.lr, n, 300, 300, want_maxsim=False))
    m.force_exact_order = True
    t1 = timed(lambda: m.align_device(w.sims, w.off, w.lq, w.lr, n, 300, 300, want_maxsim=False))
    print(f"pairs {n}: fast pipeline {t0:.3f} ms, exact-order kernel {t1:.3f} ms")

g = torch.Generator(device=dev); g.manual_seed(1)
q = torch.randn((40000, 512), generator=g, device=dev).half().float()
r = torch.randn((200000, 512), generator=g, device=dev).half().float()
oa, ob = gemm.prepare_pair(q, r)
p = gemm.Pairing(oa, ob)
print("split", p.split)
flops = 2.0 * 40000 * 200000 * 512
ms = timed(lambda: gemm.gemm_rowmax(oa, ob), 5)
print(f"rowmax, hi part inside the 3*kpad panel (row stride 3072 B): {ms:.3f} ms  {flops / ms / 1e9:.0f} TFLOP/s")
# contiguous copy of the hi parts
lib = _lib.load()
ha, hb = oa.panel[:, 1024:].contiguous(), ob.panel[:, 1024:].contiguous()
out = torch.empty((40000,), dtype=torch.float32, device=dev)
fmt = _lib.GemmFormat(1, 512, 512, p.scale.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ms = timed(lambda: lib.vsc_gemm_rowmax(ha.data_ptr(), 40000, hb.data_ptr(), 200000, 512, out.data_ptr(), ctypes.byref(fmt), st), 5)
print(f"rowmax, contiguous hi panels (row stride 1024 B): {ms:.3f} ms  {flops / ms / 1e9:.0f} TFLOP/s")
ms = timed(lambda: gemm.gemm_rowmax(oa, ob), 5)
print(f"rowmax strided again: {ms:.3f} ms")
