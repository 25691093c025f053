"""Make the engine importable under the reference's module names (drop-in use).

    import vsc2022_b200.compat; vsc2022_b200.compat.install()
    from vsc.index import VideoIndex            # -> vsc2022_b200.index
    from vcsl.vta import build_vta_model        # -> vsc2022_b200.vta
    import faiss; faiss.METRIC_INNER_PRODUCT    # constants + index_factory only

Nothing is installed unless `install()` is called, and existing real packages are left alone unless force=True.
"""
import importlib
import sys
import types

_ALIASES = {
    "vsc.index": "vsc2022_b200.index",
    "vsc.candidates": "vsc2022_b200.candidates",
    "vsc.metrics": "vsc2022_b200.metrics",
    "vsc.storage": "vsc2022_b200.storage",
    "vsc.baseline.score_normalization": "vsc2022_b200.score_normalization",
    "vsc.baseline.localization": "vsc2022_b200.localization",
    "vsc.descriptor_eval_lib": "vsc2022_b200.descriptor_eval_lib",
    "vsc.baseline.sscd_baseline": "vsc2022_b200.sscd_baseline",
    "vcsl.vta": "vsc2022_b200.vta",
}


def install(force: bool = False):
    def package(name):
        if name not in sys.modules or force:
            mod = types.ModuleType(name)
            mod.__path__ = []
            sys.modules[name] = mod
        return sys.modules[name]

    for alias, target in _ALIASES.items():
        parts = alias.split(".")
        for i in range(1, len(parts)):
            package(".".join(parts[:i]))
        try:
            mod = importlib.import_module(target)
        except ModuleNotFoundError:
            continue
        sys.modules[alias] = mod
        setattr(sys.modules[".".join(parts[:-1])], parts[-1], mod)
    if "faiss" not in sys.modules or force:
        from . import index
        faiss = types.ModuleType("faiss")
        faiss.METRIC_INNER_PRODUCT, faiss.METRIC_L2 = index.METRIC_INNER_PRODUCT, index.METRIC_L2
        faiss.index_factory = index.index_factory
        faiss.get_num_gpus = lambda: 0   # callers then use the index they were given, which already runs on the GPU
        sys.modules["faiss"] = faiss
