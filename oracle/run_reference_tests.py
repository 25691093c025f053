"""Run the reference's OWN unit tests, unmodified, on top of the oracle shims.

TEST INFRASTRUCTURE.  Only works where /root/reference is mounted (the build
container); the GPU box never runs this.  Pins the shims: all 20 reference test
methods must pass (13 metrics, 2 storage, 1 candidates, 2 index, 2 localization).

    python -m oracle.run_reference_tests [/root/reference]
"""
import importlib.util
import os
import sys
import unittest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def reference_on_path(ref_root="/root/reference"):
    """Make ``import vsc`` resolve to the unmodified reference, engines to the shims."""
    if not os.path.isdir(os.path.join(ref_root, "vsc")):
        raise FileNotFoundError(f"reference tree not found at {ref_root}")
    for p in (ref_root, REPO, os.path.join(HERE, "shims")):
        if p in sys.path:
            sys.path.remove(p)
    # shims first: /root/reference/vcsl is an empty package with a dangling
    # symlink and would otherwise shadow the TN stand-in.
    sys.path[:0] = [os.path.join(HERE, "shims"), REPO, ref_root]
    for name in [m for m in sys.modules if m.split(".")[0] in ("vsc", "vcsl", "faiss", "matplotlib")]:
        del sys.modules[name]


def load_suite(ref_root="/root/reference"):
    reference_on_path(ref_root)
    suite = unittest.TestSuite()
    loader = unittest.TestLoader()
    tests_dir = os.path.join(ref_root, "tests")
    for fn in sorted(os.listdir(tests_dir)):
        if not (fn.startswith("test_") and fn.endswith(".py")):
            continue
        spec = importlib.util.spec_from_file_location(f"_vsc_ref_{fn[:-3]}", os.path.join(tests_dir, fn))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        suite.addTests(loader.loadTestsFromModule(mod))
    return suite


def main(argv):
    ref_root = argv[1] if len(argv) > 1 else "/root/reference"
    suite = load_suite(ref_root)
    result = unittest.TextTestRunner(verbosity=1).run(suite)
    print(f"reference tests run={result.testsRun} failures={len(result.failures)} "
          f"errors={len(result.errors)} skipped={len(result.skipped)}")
    return 0 if result.wasSuccessful() and not result.skipped else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv))
