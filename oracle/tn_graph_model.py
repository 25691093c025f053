"""ctypes front end of oracle/tn_graph_model.c: the CPU model of the compact-graph TN formulation that
vsc2022_b200/csrc/tn_graph.cu runs (TEST INFRASTRUCTURE; see that file's header)."""
import ctypes
import os
import subprocess
from typing import List, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_tn_graph.so")
_lib = None

STATUS_OK, STATUS_HAND_BACK = 0, 3   # 3: table overflow or a tie generations cannot break -> exact-order kernel


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("tn_graph_model.c", "tn_fast.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_build/liboracle_tn_graph.so"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.tn_graph_model.restype = ctypes.c_int
        _lib.tn_graph_model.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_double,
                                        ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def tn(sims: np.ndarray, tn_max_step: int = 10, tn_top_k: int = 5, max_path: int = 10, min_sim: float = 0.2,
       min_length: int = 5, max_iou: float = 0.3) -> Tuple[List[List[int]], int]:
    """(boxes, status); boxes are only meaningful when status == STATUS_OK.  Raises for parameter sets outside the
    compact formulation ((tn_max_step - 1) * tn_top_k > 32)."""
    lib = _load()
    sims = np.ascontiguousarray(sims, dtype=np.float32)
    boxes = np.zeros((max_path + 1, 4), dtype=np.int32)
    status = ctypes.c_int32(0)
    n = lib.tn_graph_model(sims.ctypes.data, sims.shape[0], sims.shape[1], tn_max_step, tn_top_k, max_path, min_sim,
                           float(min_length), float(max_iou), boxes.ctypes.data, ctypes.byref(status))
    if n < 0:
        raise ValueError("tn_graph_model: parameters outside the compact formulation")
    return boxes[:n].tolist(), int(status.value)
