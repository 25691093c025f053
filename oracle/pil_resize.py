"""CPU restatement of Pillow's bilinear Image.resize for 8-bit RGB (TEST INFRASTRUCTURE -- oracle).

Third party: Pillow `src/libImaging/Resample.c` (ImagingResampleHorizontal_8bpc / ImagingResampleVertical_8bpc over the
coefficient tables of precompute_coeffs + normalize_coeffs_8bpc).  Pillow is not part of /root/reference (the reference
reaches it through torchvision.transforms.Resize, vsc/baseline/inference_impl.py:39-69) and its version is unpinned
upstream; this image has Pillow 12.2, and tests/test_preprocess_cpu.py pins this restatement against it on every
geometry the transforms produce.  The GPU kernel (csrc/resize.cu) is then compared with PIL itself.
"""
import numpy as np

from vsc2022_b200.preprocess import PRECISION_BITS, pil_coefficients


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One separable pass over `axis` (0 = rows / vertical, 1 = columns / horizontal) of uint8 [h, w, 3]."""
    bounds, kk, ksize = pil_coefficients(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)                       # [in, other, 3]
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for o in range(out_size):
        lo, cnt = int(bounds[o, 0]), int(bounds[o, 1])
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[o, :cnt].astype(np.int64), src[lo:lo + cnt], axes=(0, 0))
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear(img: np.ndarray, rh: int, rw: int) -> np.ndarray:
    """uint8 [h, w, 3] -> uint8 [rh, rw, 3]: horizontal pass into a uint8 temporary, then vertical (Resample.c order)."""
    return _pass(_pass(img, rw, 1), rh, 0)
