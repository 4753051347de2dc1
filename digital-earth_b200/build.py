"""Build libde.so (the C-ABI + CUDA kernels) for sm_100a with plain nvcc, in-tree.

    python digital-earth_b200/build.py [--force]

Three translation units:
  de_kernels.cu  -DDE_EXACT=1 -fmad=false            parity flavour (IEEE source-order arithmetic)
  de_kernels.cu  -DDE_EXACT=0 -use_fast_math         product flavour
  de_wavefront.cu            -use_fast_math          persistent-thread wavefront integrator
  de_api.cu                                          host side of the C-ABI
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libde%s.so" % os.environ.get("DE_LIB_SUFFIX", ""))
WF_DEFS = os.environ.get("DE_WF_DEFS", "").split()  # tuning sweeps: -DWF_WARPS=.. -DWF_SLOTS=..
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"] + ARCH
UNITS = [
    ("de_kernels.cu", "de_kernels_exact.o", ["-DDE_EXACT=1", "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false"]),
    ("de_kernels.cu", "de_kernels_fast.o", ["-DDE_EXACT=0", "-use_fast_math"]),
    ("de_wavefront.cu", "de_wavefront%s.o" % os.environ.get("DE_LIB_SUFFIX", ""), ["-DDE_EXACT=0", "-use_fast_math"] + WF_DEFS),
    ("de_api.cu", "de_api.o", []),
]


def _sources():
    return [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".cu", ".cuh", ".h"))] + [
        os.path.join(os.path.dirname(HERE), "include", "de_api.h"), os.path.abspath(__file__)]


def up_to_date():
    return os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in _sources())


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    bdir = os.path.join(SRC, "build")
    os.makedirs(bdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")

    def compile_one(u):
        src, obj, flags = u
        cmd = [nvcc] + COMMON + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(SRC, src), "-o", os.path.join(bdir, obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(len(UNITS)) as ex:
        logs = list(ex.map(compile_one, UNITS))
    if verbose:
        print("\n".join(logs))
    cmd = [nvcc, "-shared", "-o", OUT] + ARCH + [os.path.join(bdir, u[1]) for u in UNITS] + ["-Xcompiler", "-fvisibility=hidden"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed: %s\n%s" % (r.stdout, r.stderr))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
