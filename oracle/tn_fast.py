"""ctypes front end of oracle/tn_fast.c (TEST INFRASTRUCTURE; see that file's header)."""
import ctypes
import os
import subprocess
from typing import List, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_tn.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tn_fast.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_build/liboracle_tn.so"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.tn_fast.restype = ctypes.c_int
        _lib.tn_fast.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_double,
                                 ctypes.c_double, ctypes.c_void_p]
        _lib.tn_fast_batch.restype = ctypes.c_int
        _lib.tn_fast_batch.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 4 + [
            ctypes.c_float, ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def tn(sims: np.ndarray, tn_max_step: int = 10, tn_top_k: int = 5, max_path: int = 10,
       min_sim: float = 0.2, min_length: int = 5, max_iou: float = 0.3) -> List[List[int]]:
    lib = _load()
    sims = np.ascontiguousarray(sims, dtype=np.float32)
    boxes = np.zeros((max_path + 1, 4), dtype=np.int32)
    n = lib.tn_fast(sims.ctypes.data, sims.shape[0], sims.shape[1], tn_max_step, tn_top_k,
                    max_path, min_sim, float(min_length), float(max_iou), boxes.ctypes.data)
    if n < 0:
        raise ValueError("tn_fast: unsupported parameters")
    return boxes[:n].tolist()


def tn_batch(sims: Sequence[np.ndarray], tn_max_step: int = 10, tn_top_k: int = 5,
             max_path: int = 10, min_sim: float = 0.2, min_length: int = 5,
             max_iou: float = 0.3) -> List[List[List[int]]]:
    lib = _load()
    n = len(sims)
    lq = np.array([s.shape[0] for s in sims], dtype=np.int32)
    lr = np.array([s.shape[1] for s in sims], dtype=np.int32)
    off = np.zeros(n, dtype=np.int64)
    if n:
        off[1:] = np.cumsum(lq[:-1].astype(np.int64) * lr[:-1])
    flat = np.concatenate([np.asarray(s, dtype=np.float32).ravel() for s in sims]) if n else np.zeros(0, np.float32)
    boxes = np.zeros((n, max_path + 1, 4), dtype=np.int32)
    cnt = np.zeros(n, dtype=np.int32)
    rc = lib.tn_fast_batch(flat.ctypes.data, off.ctypes.data, lq.ctypes.data, lr.ctypes.data, n,
                           tn_max_step, tn_top_k, max_path, min_sim, float(min_length),
                           float(max_iou), boxes.ctypes.data, cnt.ctypes.data)
    if rc != 0:
        raise ValueError("tn_fast_batch: unsupported parameters")
    return [boxes[i, :cnt[i]].tolist() for i in range(n)]
