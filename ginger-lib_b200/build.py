"""Build recipe for libg753.so (hand-written CUDA for sm_100a, in-tree so it travels with gpurun).

nvcc cross-compiles without a GPU.  Each translation unit (the C ABI + NTT, and one per curve
group for the MSM pipeline) is compiled in parallel and linked into one shared library.
Entry: build_library(force=False).
"""
import hashlib
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


def source_hash():
    """sha256 over every source the library is built from (csrc/*, include/g753.h) and the compiler
    flags.  It is compiled into the library (g753_source_hash()), so a loaded libg753.so can always be
    matched against the tree it claims to come from."""
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for path in sorted(_deps()):
        h.update(os.path.basename(path).encode() + b"\0")
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def built_hash():
    """the source hash recorded by the build that produced the in-tree library (None: never built)"""
    try:
        return open(os.path.join(OBJ, "source_hash.txt")).read().strip()
    except OSError:
        return None


def _stale(target, want_hash):
    return not os.path.exists(target) or built_hash() != want_hash


def _nvcc():
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return nvcc if os.path.exists(nvcc) else "nvcc"


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu into ginger-lib_b200/libg753.so for sm_100a. Returns the path."""
    want = source_hash()
    if not force and not _stale(LIB, want):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    define = ["-DG753_SOURCE_HASH=\"%s\"" % want]

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + define + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
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
    with open(os.path.join(OBJ, "source_hash.txt"), "w") as fh:
        fh.write(want + "\n")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
