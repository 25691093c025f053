"""Frame descriptors between the stages without host round trips.

The reference hands descriptors from stage to stage as numpy arrays inside `VideoFeature`s (and as .npz files between
processes: vsc/storage.py:13-68).  Here a `VideoFeature.feature` may also be a CUDA tensor: `score_normalize(...,
on_device=True)` returns row views of one device matrix, and `CandidateGeneration`, `VideoIndex` and the localization
classes take them as they are.  Host arrays that are row views of one big array -- what `storage.load_features` returns --
are uploaded with a single copy.
"""
import dataclasses
from typing import List, Sequence

import numpy as np

from . import _lib


def is_device_tensor(x) -> bool:
    return hasattr(x, "is_cuda") and bool(x.is_cuda)


def root_of(arr: np.ndarray, cache: dict = None):
    """(root array, first row of `arr` inside it) if `arr` is a block of whole rows of a base array (1-D or 2-D,
    C-contiguous), else None.  `cache` memoises the per-root facts (this runs once per video of a collection)."""
    root = arr.base
    if root.__class__ is not np.ndarray and not isinstance(root, np.ndarray):
        return None
    base = root.base
    while base is not None and isinstance(base, np.ndarray):
        root, base = base, base.base
    info = cache.get(id(root)) if cache is not None else None
    if info is None:
        ok = root.ndim in (1, 2) and root.flags.c_contiguous and root.shape[0] > 0
        info = (root, ok, root.__array_interface__["data"][0], root.strides, root.shape[1:], root.dtype, root.shape[0])
        if cache is not None:
            cache[id(root)] = info
    _, ok, ptr, strides, tail, dtype, rows = info
    if not ok or arr.dtype != dtype or arr.shape[1:] != tail or (arr.shape[0] and arr.strides != strides):
        return None
    row, rem = divmod(arr.__array_interface__["data"][0] - ptr, strides[0])
    if rem or row < 0 or row + arr.shape[0] > rows:
        return None
    return root, row


def features_matrix(features: Sequence, device):
    """All frames of `features` (VideoFeatures) as one float32 CUDA matrix [sum(len), d], rows in list order.
    Device tensors are used in place when they are consecutive row views of one tensor; host arrays that are consecutive
    row views of one array go up in a single copy; anything else is concatenated first."""
    torch = _lib.require_cuda()
    if not features:
        return torch.zeros((0, 0), dtype=torch.float32, device=device)
    feats = [f.feature for f in features]
    if all(is_device_tensor(x) for x in feats):
        first = feats[0]
        row_elems = first.shape[1]
        consecutive = first.dim() == 2 and first.is_contiguous()
        at = first.storage_offset()
        for x in feats:
            if not (consecutive and x.dim() == 2 and x.is_contiguous() and x.shape[1] == row_elems and x.dtype == first.dtype
                    and x.untyped_storage().data_ptr() == first.untyped_storage().data_ptr() and x.storage_offset() == at):
                consecutive = False
                break
            at += x.shape[0] * row_elems
        if consecutive:
            rows = (at - first.storage_offset()) // max(row_elems, 1)
            mat = torch.as_strided(first, (rows, row_elems), (row_elems, 1), first.storage_offset())
        else:
            mat = torch.cat(feats)
        return mat.to(device=device, dtype=torch.float32)
    feats = [np.asarray(x) for x in feats]
    cache = {}
    where = [root_of(x, cache) for x in feats]
    if all(w is not None for w in where) and all(w[0] is where[0][0] for w in where):
        at = where[0][1]
        ok = True
        for w, x in zip(where, feats):
            if w[1] != at:
                ok = False
                break
            at += x.shape[0]
        if ok and where[0][0].dtype in (np.float32, np.float16):
            host = where[0][0][where[0][1]:at]
            return torch.from_numpy(host).to(device, non_blocking=True).float()
    host = np.concatenate([np.asarray(x, dtype=np.float32) for x in feats], axis=0)
    return torch.from_numpy(host).to(device)


def split_rows(features: Sequence, mat, on_device: bool) -> List:
    """The inverse: VideoFeatures like `features` whose descriptors are the rows of `mat` (device views, or host arrays)."""
    src = mat if on_device else mat.cpu().numpy()
    out, at = [], 0
    for f in features:
        n = len(f)
        out.append(dataclasses.replace(f, feature=src[at:at + n]))
        at += n
    return out


def to_host(features: Sequence) -> List:
    """VideoFeatures with numpy descriptors (for storage.store_features and other host consumers)."""
    if not any(is_device_tensor(f.feature) for f in features):
        return list(features)
    torch = _lib.require_cuda()
    mat = features_matrix(features, features[0].feature.device)
    return split_rows(features, mat, on_device=False)
