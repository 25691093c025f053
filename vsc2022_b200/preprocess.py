"""Frame preprocessing ahead of the SSCD model, on the GPU.

Mirror of `build_transforms` (vsc/baseline/inference_impl.py:39-69) and `InferenceTransforms`
(vsc/baseline/inference.py:28-34).  The reference composes torchvision transforms on PIL images:

    RESIZE_288          Resize(288)                     short edge -> 288, aspect ratio kept
    RESIZE_320_CENTER   Resize(320) + CenterCrop(320)
    RESIZE_224_SQUARE   Resize((224, 224))
    then ToTensor() (/255, HWC -> CHW) and Normalize(mean, std).

Here decoded uint8 RGB frames [n, H, W, 3] (numpy or torch, host or device) are resized by `vsc_resize_u8`
(csrc/resize.cu): Pillow's two-pass fixed-point bilinear resample, bit for bit, crop folded in.  The output stays
uint8 NHWC on the device; ToTensor + Normalize happen inside the stem kernel of the model (csrc/sscd_ops.cu), so a
`GpuTransform` followed by `SSCDResNet50` computes what the reference's transform followed by its model computes.

Geometry rules restated from torchvision (third party, unpinned by the reference; 0.26 in this image):
`Resize(int)`: short edge = size, long edge = int(size * long / short) (functional._compute_resized_output_size);
`CenterCrop`: top = int(round((h - ch) / 2.0)), left likewise (functional.center_crop), Python's round.
Coefficients restated from Pillow `src/libImaging/Resample.c` (precompute_coeffs, bilinear_filter,
normalize_coeffs_8bpc).
"""
import ctypes
import enum
import functools
import math
from typing import Tuple

import numpy as np

from . import _lib

PRECISION_BITS = 32 - 8 - 2


class InferenceTransforms(enum.Enum):
    # Aspect-ratio preserving resize to 288
    RESIZE_288 = enum.auto()
    # Resize the short edge to 320, then take the center crop
    RESIZE_320_CENTER = enum.auto()
    # Resize to 224x224
    RESIZE_224_SQUARE = enum.auto()


@functools.lru_cache(maxsize=64)
def pil_coefficients(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Pillow's bilinear resampling windows and fixed-point weights for one axis.

    Returns (bounds int32 [out_size, 2] = (first input pixel, count), weights int32 [out_size, ksize], ksize).
    Every floating-point operation is the double-precision operation of Resample.c, in its order."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale                        # bilinear: support 1
    ksize = int(math.ceil(support)) * 2 + 1
    xx = np.arange(out_size, dtype=np.float64)
    center = 0.0 + (xx + 0.5) * scale
    ss = 1.0 / filterscale
    xmin = np.trunc(center - support + 0.5).astype(np.int64)      # (int) truncates towards zero
    xmin = np.maximum(xmin, 0)
    xmax = np.trunc(center + support + 0.5).astype(np.int64)
    xmax = np.minimum(xmax, in_size) - xmin
    k = np.zeros((out_size, ksize), dtype=np.float64)
    ww = np.zeros(out_size, dtype=np.float64)
    for x in range(ksize):                                         # sequential accumulation, like the C loop
        live = x < xmax
        arg = np.abs(((x + xmin).astype(np.float64) - center + 0.5) * ss)
        w = np.where(live & (arg < 1.0), 1.0 - arg, 0.0)
        k[:, x] = w
        ww = ww + w
    nz = ww != 0.0
    k[nz] = k[nz] / ww[nz, None]
    fixed = np.where(k < 0, np.trunc(-0.5 + k * (1 << PRECISION_BITS)), np.trunc(0.5 + k * (1 << PRECISION_BITS)))
    bounds = np.stack([xmin, xmax], axis=1).astype(np.int32)
    return bounds, np.ascontiguousarray(fixed.astype(np.int32)), ksize


def resized_geometry(transform: InferenceTransforms, h: int, w: int) -> Tuple[int, int, int, int, int, int]:
    """(rh, rw, top, left, oh, ow): the frame is resized to rh x rw, the window [top, top+oh) x [left, left+ow) kept."""
    if transform == InferenceTransforms.RESIZE_224_SQUARE:
        return 224, 224, 0, 0, 224, 224
    size = 288 if transform == InferenceTransforms.RESIZE_288 else 320
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    rw, rh = (new_short, new_long) if w <= h else (new_long, new_short)
    if transform == InferenceTransforms.RESIZE_288:
        return rh, rw, 0, 0, rh, rw
    top, left = int(round((rh - size) / 2.0)), int(round((rw - size) / 2.0))
    return rh, rw, top, left, size, size


class GpuTransform:
    """Callable standing in for the Compose the reference builds: uint8 frames [n, H, W, 3] -> uint8 CUDA tensor
    [n, oh, ow, 3] resized (and cropped) like PIL; normalisation is left to the model's stem kernel."""

    def __init__(self, transform: InferenceTransforms, device=None):
        self.transform, self.device = transform, device
        self._tables = {}

    def _device_tables(self, torch, dev, in_size, out_size):
        key = (in_size, out_size, str(dev))
        if key not in self._tables:
            bounds, kk, ksize = pil_coefficients(in_size, out_size)
            self._tables[key] = (torch.from_numpy(bounds).to(dev), torch.from_numpy(kk).to(dev), ksize)
        return self._tables[key]

    def __call__(self, frames):
        torch = _lib.require_cuda()
        lib = _lib.load()
        if not isinstance(frames, torch.Tensor):
            frames = torch.from_numpy(np.ascontiguousarray(frames))
        if frames.dim() == 3:
            frames = frames[None]
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3:
            raise ValueError("frames must be uint8 [n, H, W, 3]")
        dev = torch.device(self.device) if self.device is not None else (
            frames.device if frames.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        frames = frames.to(dev, non_blocking=True).contiguous()
        n, h, w, _ = frames.shape
        rh, rw, top, left, oh, ow = resized_geometry(self.transform, h, w)
        out = torch.empty((n, oh, ow, 3), dtype=torch.uint8, device=dev)
        if n == 0:
            return out
        xb, xk, xks = self._device_tables(torch, dev, w, rw)
        yb, yk, yks = self._device_tables(torch, dev, h, rh)
        tmp = torch.empty((n, h, ow, 3), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.vsc_resize_u8(frames.data_ptr(), n, h, w, rh, rw, top, left, oh, ow, xb.data_ptr(), xk.data_ptr(), xks,
                                   yb.data_ptr(), yk.data_ptr(), yks, tmp.data_ptr(), out.data_ptr(),
                                   ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(rc, "vsc_resize_u8")
        return out


def build_transforms(transform: InferenceTransforms, device=None) -> GpuTransform:
    """inference_impl.py:39-69 for decoded uint8 frames; accepts the enum member or its name (the CLI's --transforms)."""
    if isinstance(transform, str):
        transform = InferenceTransforms[transform]
    elif not isinstance(transform, InferenceTransforms):
        transform = InferenceTransforms[transform.name]      # the reference's own enum class
    return GpuTransform(transform, device)
