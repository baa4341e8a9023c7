"""Build libpmce_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The shared object lands next to the sources' package (`pmce_b200/libpmce_b200.so`); it is
git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# PMCE_B200_PROFILING=1: a separate library compiled with -DPMCE_PROFILING, the only build in which the result-corrupting
# timing knobs (PMCE_TC_NULL / PMCE_TC_DBG / PMCE_ATTN_DEBUG, tools/gemm_sweep.py) are honoured
PROFILING = os.environ.get("PMCE_B200_PROFILING", "0") == "1"
LIB = os.path.join(HERE, "libpmce_b200_prof.so" if PROFILING else "libpmce_b200.so")
SOURCES = ["api.cu", "layout.cpp"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "pmce_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into libpmce_b200.so. Returns the library path."""
    if not force and not _stale():
        return LIB
    cmd = [_nvcc(), "-O3", "-std=c++17", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC", "-shared", *(["-DPMCE_PROFILING"] if PROFILING else []),
           "-Xptxas", "-v" if verbose else "-warn-spills", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libpmce_b200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
