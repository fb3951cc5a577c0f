"""The C ABI from a COMPILED caller (integration/c_caller/abi_caller.c, plain C99 against include/g753.h):
compiled and linked here without a GPU (the program must then report the library's refusal to run - there
is no CPU fallback), run for real on a B200.  Also: the Rust -sys crate declares every function of the
header (the crate cannot be compiled in this image)."""
import importlib
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = importlib.import_module("ginger-lib_b200")
SRC = os.path.join(ROOT, "integration", "c_caller", "abi_caller.c")
LIBDIR = os.path.join(ROOT, "ginger-lib_b200")


def build_caller(tmp_path):
    importlib.import_module("__graft_entry__").build()
    exe = str(tmp_path / "abi_caller")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC,
                           "-L", LIBDIR, "-lg753", "-Wl,-rpath," + LIBDIR, "-o", exe])
    return exe


def generator_file(tmp_path):
    params = importlib.import_module("ginger-lib_b200.params")
    gen = np.array([[(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(12)] for v in params.GENERATOR_MONT[0]],
                   dtype=np.uint64)
    path = str(tmp_path / "gen_g1.bin")
    gen.tofile(path)
    return path


def test_c_caller_links_and_refuses_without_device(tmp_path):
    exe = build_caller(tmp_path)
    if G.ffi.Library().device_count_safe() > 0:
        pytest.skip("a GPU is present: see test_c_caller_on_gpu")
    proc = subprocess.run([exe, generator_file(tmp_path)], capture_output=True, text=True, timeout=120)
    assert proc.returncode == 3, proc.stdout + proc.stderr
    assert "no CUDA device" in proc.stdout and "g753" in proc.stdout


@pytest.mark.gpu
def test_c_caller_on_gpu(tmp_path):
    exe = build_caller(tmp_path)
    proc = subprocess.run([exe, generator_file(tmp_path)], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert "all checks passed" in proc.stdout


def test_rust_sys_crate_declares_every_function():
    header = open(os.path.join(ROOT, "include", "g753.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(g753_[a-z0-9_]+)\s*\(", header)))
    rust = open(os.path.join(ROOT, "integration", "algebra-cuda-sys", "src", "lib.rs")).read()
    bound = sorted(set(re.findall(r"pub fn (g753_[a-z0-9_]+)\s*\(", rust)))
    assert declared == bound
