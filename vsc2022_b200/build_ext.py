"""Builds vsc2022_b200/csrc/libvsc_b200.so with nvcc for sm_100a (in-tree, no JIT cache)."""
import glob
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
REPO = os.path.dirname(PKG)
LIB = os.path.join(CSRC, "libvsc_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
    "-I", os.path.join(REPO, "include"), "-I", CSRC,
] + (["-DVSC_TN_COUNTERS"] if os.environ.get("VSC_TN_COUNTERS") else [])


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libvsc_b200.so")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(REPO, "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu in csrc/ to objects (parallel) and link the shared library."""
    if not force and not is_stale():
        return LIB
    nvcc = find_nvcc()
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            print(out)
        if p.returncode:
            failed.append(src)
    if failed:
        raise RuntimeError(f"nvcc failed for {failed}")
    subprocess.check_call([nvcc, "-shared", "-o", LIB, *objs, "-lcudart"])
    return LIB


HOSTGLUE = os.path.join(CSRC, "_hostglue.so")


def build_hostglue(force: bool = False) -> str:
    """csrc/hostglue.c (CPython API, buffer protocol) -> csrc/_hostglue.so with gcc."""
    import sysconfig
    src = os.path.join(CSRC, "hostglue.c")
    if not force and os.path.exists(HOSTGLUE) and os.path.getmtime(HOSTGLUE) >= os.path.getmtime(src):
        return HOSTGLUE
    gcc = shutil.which("gcc") or shutil.which("cc")
    if not gcc:
        raise RuntimeError("gcc not found: cannot build _hostglue.so")
    subprocess.check_call([gcc, "-O2", "-Wall", "-shared", "-fPIC", "-I", sysconfig.get_paths()["include"], src, "-o", HOSTGLUE])
    return HOSTGLUE


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_hostglue(force="--force" in sys.argv))
