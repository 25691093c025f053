"""Stage A throughput: SSCD ResNet-50 on synthetic 288x288 frames (dev tool)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from vsc2022_b200 import _lib  # noqa: E402
from vsc2022_b200.sscd import SSCDResNet50, TorchReference, normalize_pixels  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
ref = TorchReference(seed=0)
ours = SSCDResNet50(ref.trunk, ref.head)
g = torch.Generator(device="cuda"); g.manual_seed(0)
frames = torch.randint(0, 256, (n, 288, 288, 3), generator=g, device="cuda", dtype=torch.uint8)
for _ in range(2):
    ours.forward(frames[:batch], batch=batch)
torch.cuda.synchronize()
l0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = ours.forward(frames, batch=batch)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"ours: {n} frames in {ms:.1f} ms -> {n / ms * 1e3:.0f} frames/s, {n * 13.513e9 / ms / 1e9:.0f} TFLOP/s, "
      f"{(_lib.launch_count() - l0)} launches")
x = normalize_pixels(frames[:256])
with torch.autocast("cuda", dtype=torch.bfloat16):
    ref(x[:64])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ref(x.contiguous(memory_format=torch.channels_last))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"torch bf16 autocast (cuDNN, channels_last): 256 frames in {dt*1e3:.1f} ms -> {256/dt:.0f} frames/s")
