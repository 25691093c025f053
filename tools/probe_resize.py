"""Probe: GPU frame resize (vsc_resize_u8) on full-HD frames, CUDA events (dev tool)."""
import sys

import torch

sys.path.insert(0, ".")
from vsc2022_b200.preprocess import InferenceTransforms as T, build_transforms  # noqa: E402

g = torch.Generator(device="cuda"); g.manual_seed(0)
frames = torch.randint(0, 256, (64, 1080, 1920, 3), generator=g, device="cuda", dtype=torch.uint8)
for t in T:
    fn = build_transforms(t)
    for _ in range(2):
        out = fn(frames)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = fn(frames)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{t.name:20s} 64 x 1080x1920 -> {tuple(out.shape[1:3])}: {ms:.3f} ms = {64 / ms * 1e3:.0f} frames/s, "
          f"{frames.numel() / ms / 1e6:.0f} GB/s of input")
