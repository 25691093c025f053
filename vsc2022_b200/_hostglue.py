"""Loader of csrc/_hostglue.so (csrc/hostglue.c): the two interpreter-bound loops of localize_all in C.

Optional by design: it is host-side glue with a Python twin next to every call site (same results, tests/test_hostglue_cpu.py);
`load()` returns None when the library has not been built.  This is NOT a fallback for device work -- that lives in
libvsc_b200.so and fails loudly when missing (_lib.py)."""
import ctypes
import os

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "_hostglue.so")
_lib = None
_tried = False


def load():
    global _lib, _tried
    if not _tried:
        _tried = True
        if os.path.exists(_PATH) and not os.environ.get("VSC_NO_HOSTGLUE"):
            lib = ctypes.PyDLL(_PATH)
            lib.vsc_match_rows.restype = ctypes.py_object
            lib.vsc_match_rows.argtypes = [ctypes.py_object] * 9
            lib.vsc_scan_views.restype = ctypes.py_object
            lib.vsc_scan_views.argtypes = [ctypes.py_object] * 7
            _lib = lib
    return _lib
