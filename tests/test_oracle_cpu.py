"""CPU suite: pins the oracle (reference tests over the shims, golden vectors,
C formulation vs networkx) and checks the C-ABI library loads with every symbol."""
import os
import re

import numpy as np
import pytest

from oracle import localize_numpy, search_numpy, synth, tn_fast, tn_networkx

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/vsc"), reason="reference tree not mounted")
def test_unmodified_reference_tests_pass_over_shims():
    import subprocess
    import sys
    out = subprocess.run([sys.executable, "-m", "oracle.run_reference_tests"], cwd=REPO,
                         capture_output=True, text=True)
    assert "reference tests run=20 failures=0 errors=0 skipped=0" in out.stdout, out.stdout + out.stderr


def test_restatement_matches_golden_c1(golden_c1):
    g = golden_c1
    q, r, noise = list(g["q"]), list(g["r"]), list(g["noise"])
    for tag, (qq, rr) in {"raw": (q, r), "sn": search_numpy.score_normalize(q, r, noise, beta=1.2)}.items():
        if tag == "sn":
            assert np.array_equal(np.stack(qq), g["sn_q"]) and np.array_equal(np.stack(rr), g["sn_r"])
        cands = search_numpy.candidates(qq, rr, 2400)
        assert [[a, b + 100] for a, b, _ in cands] == g[f"{tag}_cand_ids"].tolist()
        assert np.array_equal(np.array([s for _, _, s in cands], np.float32), g[f"{tag}_cand_score"])
        pairs = search_numpy.search_pairs(qq, rr, 2400)
        flat = [[a, b + 100] for a, b, ms in pairs for _ in ms]
        assert flat == g[f"{tag}_pairmatch_ids"].tolist()


def test_known_answer_candidates():
    """tests/test_candidates.py:17-83 of the reference, restated on arrays."""
    q = [np.eye(3, dtype=np.float32)]
    refs = [np.array([[0, 0, 0], [0, 0, 0], [0, 1, 0], [0, 2, 0], [0, 0, 0]], np.float32),
            np.array([[0, 0, 0], [1, 0, 0], [1, 0, 0]], np.float32),
            np.array([[0, 0, 0], [0, 0, 0.25], [0, 0, 0]], np.float32)]
    assert search_numpy.candidates(q, refs, 6) == [(0, 0, 2.0), (0, 1, 1.0), (0, 2, 0.25)]


def test_tn_golden_and_c_formulation(golden_tn):
    g = golden_tn
    for i in range(int(g["n"])):
        sims = g[f"sims_{i}"]
        for tag, cfg in (("vsc", dict(tn_max_step=5, min_length=4)), ("default", {})):
            want = g[f"boxes_{tag}_{i}"].tolist()
            assert tn_fast.tn(sims, **cfg) == want
            if sims.size <= 128 * 128:
                assert tn_networkx.tn(sims, **cfg) == want


def test_c_formulation_matches_networkx_random():
    rng = np.random.default_rng(0)
    for it in range(250):
        lq, lr = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        sims = synth.sim_matrix(rng, lq, lr, bias=[0.5, 0.0][it % 2], quant=[0.0, 8.0, 4.0, 16.0][it % 4])
        cfg = dict(tn_max_step=[5, 10, 3, 7][it % 4], tn_top_k=[5, 5, 3, 6][(it // 4) % 4],
                   min_length=[4, 5, 2][it % 3])
        assert tn_networkx.tn(sims, **cfg) == tn_fast.tn(sims, **cfg), (it, lq, lr, cfg)


def test_localize_restatement_matches_golden(golden_tn):
    g = golden_tn
    a, b, c = g["loc_a"], g["loc_b"], g["loc_c"]
    pairs = [(a, g["loc_ts_a"], b, np.arange(30) * 1.0, 1.0), (a, g["loc_ts_a"], c, g["loc_ts_c"], 2.0)]
    rows = localize_numpy.localize_all(pairs, 0.5, "max_sim", tn_max_step=5, min_length=4)
    flat = [r for per in rows for r in per]
    assert np.array_equal(np.array([r[:4] for r in flat]), g["loc_match_ts"])
    assert np.array_equal(np.array([r[4] for r in flat], np.float32), g["loc_match_score"])


def test_library_exports_every_declared_symbol():
    from vsc2022_b200 import _lib, build_ext
    build_ext.build_library()
    lib = _lib.load()
    header = open(os.path.join(REPO, "include", "vsc_b200.h")).read()
    declared = set(re.findall(r"\b(vsc_[a-z0-9_]+|vcsl_[a-z0-9_]+|sscd_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes found in include/vsc_b200.h"
    for name in declared:
        assert hasattr(lib, name), f"libvsc_b200.so does not export {name}"
    assert set(_lib.EXPORTS) == declared
    assert lib.vsc_abi_version() >= 1


def test_ctypes_prototypes_agree_with_the_header():
    """Every prototype of include/vsc_b200.h against the ctypes declaration in _lib.py: same number of parameters,
    pointers bound as pointers, 64-bit integers as 64-bit, floats as floats (an ABI drift crashes instead of failing)."""
    import ctypes
    from vsc2022_b200 import _lib
    header = open(os.path.join(REPO, "include", "vsc_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    protos = re.findall(r"\b(?:int|int64_t|const char \*)\s*\*?\s*((?:vsc|vcsl)_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", header, flags=re.S)
    assert len(protos) == len(_lib.EXPORTS), (len(protos), len(_lib.EXPORTS))
    for name, params in protos:
        params = " ".join(params.split())
        args = [] if params in ("", "void") else [a.strip() for a in params.split(",")]
        _, argtypes = _lib.EXPORTS[name]
        assert len(args) == len(argtypes), f"{name}: header has {len(args)} parameters, ctypes {len(argtypes)}"
        for a, t in zip(args, argtypes):
            if "*" in a or a.startswith("vsc_stream_t"):
                assert t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or issubclass(t, ctypes._Pointer), (name, a, t)
            elif a.startswith(("int64_t", "uint64_t")):
                assert t in (ctypes.c_int64, ctypes.c_uint64), (name, a, t)
            elif a.startswith(("int32_t", "int ", "uint32_t")):
                assert t in (ctypes.c_int32, ctypes.c_int, ctypes.c_uint32), (name, a, t)
            elif a.startswith("float"):
                assert t is ctypes.c_float, (name, a, t)
            elif a.startswith("double"):
                assert t is ctypes.c_double, (name, a, t)
            else:
                raise AssertionError(f"{name}: unrecognised parameter type {a!r}")


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from vsc2022_b200 import _lib, vta
    with pytest.raises(_lib.EngineError):
        vta.build_vta_model("TN").forward_sim([("a", np.zeros((4, 4), np.float32))])


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/vsc_b200.h must compile as C99 (no C++ or torch types)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "use_header.c"
    src.write_text('#include "vsc_b200.h"\nint main(void) { vsc_tn_params p; (void)p; return 0; }\n')
    out = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(REPO, "include"),
                          str(src)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_vcsl_sink_node_variants():
    """upstream VCSL links a zero-weight "sink"; its source is missing, so every reading of it is run.
    A dedicated sink node must never change a box (goldens + random matrices, ties included); the reading in which the
    last real node plays the sink is counted."""
    import numpy as np
    from oracle import synth, tn_fast, tn_networkx
    golden = np.load(os.path.join(REPO, "tests", "golden", "tn_reference.npz"))
    cases = [(golden[f"sims_{i}"], dict(tn_max_step=5, min_length=4)) for i in range(int(golden["n"]))]
    cases += [(golden[f"sims_{i}"], {}) for i in range(int(golden["n"]))]
    rng = np.random.default_rng(77)
    for it in range(160):
        lq, lr = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        s = synth.sim_matrix(rng, lq, lr, bias=[0.5, 0.0][it % 2], quant=[0.0, 8.0, 0.0, 64.0][it % 4])
        cases.append((s, dict(tn_max_step=5, min_length=4) if it % 2 else dict(tn_max_step=4, tn_top_k=3, min_length=2)))
    changed_by_last_node = 0
    for s, cfg in cases:
        base = tn_networkx.tn(s, **cfg)
        assert tn_networkx.tn(s, sink="dedicated", **cfg) == base, (s.shape, cfg)
        if (cfg.get("tn_max_step", 10) - 1) * min(cfg.get("tn_top_k", 5), s.shape[1]) <= 64:
            assert tn_fast.tn(s, **cfg) == base
        changed_by_last_node += tn_networkx.tn(s, sink="last_node", **cfg) != base
    print(f"sink = last real node changes the boxes of {changed_by_last_node} of {len(cases)} matrices")
    assert changed_by_last_node <= len(cases) // 4
