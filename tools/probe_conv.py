"""Time / profile single convolution-GEMM shapes of the SSCD trunk at batch 128 x 288^2 (dev tool)."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from vsc2022_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", 0)
sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
shapes = [  # (name, M, K, N, residual)
    ("l1.c3  72x72 64->256 +res", 128 * 72 * 72, 64, 256, True),
    ("l1.down 72x72 64->256", 128 * 72 * 72, 64, 256, False),
    ("l1.c1  72x72 256->64", 128 * 72 * 72, 256, 64, False),
    ("l1.c2  72x72 576->64", 128 * 72 * 72, 576, 64, False),
    ("l2.c3  36x36 128->512 +res", 128 * 36 * 36, 128, 512, True),
    ("l3.c3  18x18 256->1024 +res", 128 * 18 * 18, 256, 1024, True),
    ("l3.c2  18x18 2304->256", 128 * 18 * 18, 2304, 256, False),
    ("l4.c2  9x9 4608->512", 128 * 9 * 9, 4608, 512, False),
]
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, m, k, n, has_res in shapes:
    if only and only not in name:
        continue
    a = torch.randn((m, k), device=dev).bfloat16()
    w = torch.randn((n, k), device=dev).bfloat16()
    bias = torch.randn((n,), device=dev)
    res = torch.randn((m, n), device=dev).bfloat16() if has_res else None
    out = torch.empty((m, n), device=dev, dtype=torch.bfloat16)
    def run():
        _lib.check(lib.vsc_gemm_conv(a.data_ptr(), m, w.data_ptr(), n, k, bias.data_ptr(),
                                     res.data_ptr() if has_res else None, 1, out.data_ptr(), n, sp), "conv")
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    bytes_ = 2.0 * (m * k + m * n * (2 if has_res else 1))
    print(f"{name:30s} {ms * 1e3:7.1f} us  {bytes_ / ms / 1e6:7.0f} GB/s  {2.0 * m * k * n / ms / 1e9:7.0f} TFLOP/s")
