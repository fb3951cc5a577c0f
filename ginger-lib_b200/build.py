"""Build recipe for libg753.so (hand-written CUDA for sm_100a, in-tree so it travels with gpurun).

nvcc cross-compiles without a GPU.  Each translation unit (the C ABI + NTT, and one per curve
group for the MSM pipeline) is compiled in parallel and linked into one shared library.
Entry: build_library(force=False).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libg753.so")
SOURCES = ["capi.cu", "msm_g0.cu", "msm_g1.cu", "msm_g2.cu", "msm_g3.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".inc"))]
    out.append(os.path.join(HERE, "..", "include", "g753.h"))
    return out


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _nvcc():
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return nvcc if os.path.exists(nvcc) else "nvcc"


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu into ginger-lib_b200/libg753.so for sm_100a. Returns the path."""
    deps = _deps()
    if not force and not _stale(LIB, deps):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if verbose:
            with open(os.path.join(OBJ, src + ".ptxas.log"), "w") as fh:
                fh.write(proc.stderr)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, proc.stdout, proc.stderr))
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + proc.stdout + proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
