"""ctypes loader for oracle/libref753.so (the C++ restatement of the reference algorithms).

TEST INFRASTRUCTURE ONLY - see the header of oracle/ref753.cpp.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libref753.so")

_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "ref753.cpp")
        if not os.path.exists(LIB) or os.path.getmtime(src) > os.path.getmtime(LIB):
            subprocess.check_call(["make", "-s", "-C", HERE])
        L = ctypes.CDLL(LIB)
        vp, sz, i, u = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint
        L.ref753_hardware_threads.restype = u
        L.ref753_constant.argtypes = [i, i, vp]
        L.ref753_field_op.argtypes = [i, i, vp, vp, vp, sz]
        L.ref753_ext_op.argtypes = [i, i, vp, vp, vp]
        L.ref753_point_op.argtypes = [i, i, vp, vp, vp]
        L.ref753_msm.argtypes = [i, vp, vp, sz, vp, sz, vp, u]
        L.ref753_msm_windows.argtypes = [i, vp, vp, sz, vp, sz, u, u, vp, u]
        L.ref753_msm_window_bits.argtypes = [sz]
        L.ref753_msm_window_bits.restype = u
        L.ref753_walk.argtypes = [i, vp, vp, sz, vp, u]
        L.ref753_fft.argtypes = [i, vp, u, i, u]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


GROUP_K = {0: 1, 1: 2, 2: 1, 3: 3}


def hardware_threads():
    return int(lib().ref753_hardware_threads())


def field_op(field, op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 12)
    bb = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 12) if b is not None else None
    out = np.zeros_like(a)
    assert lib().ref753_field_op(field, op, _p(a), _p(bb), _p(out), a.shape[0]) == 0
    return out


def ext_op(ext, op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1)
    bb = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1) if b is not None else None
    out = np.zeros_like(a)
    assert lib().ref753_ext_op(ext, op, _p(a), _p(bb), _p(out)) == 0
    return out


def point_op(group, op, a, b=None):
    k = GROUP_K[group]
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1)
    bb = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1) if b is not None else None
    out = np.zeros(3 * k * 12 + 1, dtype=np.uint64)
    assert lib().ref753_point_op(group, op, _p(a), _p(bb), _p(out)) == 0
    return out


def msm(group, coords, infinity, scalars, nthreads=None):
    """VariableBaseMSM::multi_scalar_mul restated: returns the (3, k*12) GroupProjective limbs."""
    k = GROUP_K[group]
    coords = np.ascontiguousarray(coords, dtype=np.uint64).reshape(-1, 2 * k * 12)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 12)
    inf = np.ascontiguousarray(infinity, dtype=np.uint8) if infinity is not None else None
    out = np.zeros((3, k * 12), dtype=np.uint64)
    nt = nthreads or hardware_threads()
    assert lib().ref753_msm(group, _p(coords), _p(inf), coords.shape[0], _p(scalars), scalars.shape[0], _p(out), nt) == 0
    return out


def msm_window_bits(n_scalars):
    """c of variable_base.rs:14-18 for this many scalars; the MSM has ceil(753 / c) windows"""
    return int(lib().ref753_msm_window_bits(n_scalars))


def msm_windows(group, coords, infinity, scalars, first, count, nthreads=None):
    """`count` of the per-window tasks of the same MSM (variable_base.rs:30-70), windows [first, first+count):
    returns the window sums, (count, 3, k*12) GroupProjective limbs.  The bounded sample of bench.py's
    CPU baseline; folded over all windows it IS multi_scalar_mul (tests/test_oracle_cpp.py)."""
    k = GROUP_K[group]
    coords = np.ascontiguousarray(coords, dtype=np.uint64).reshape(-1, 2 * k * 12)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 12)
    inf = np.ascontiguousarray(infinity, dtype=np.uint8) if infinity is not None else None
    out = np.zeros((count, 3, k * 12), dtype=np.uint64)
    nt = nthreads or hardware_threads()
    assert lib().ref753_msm_windows(group, _p(coords), _p(inf), coords.shape[0], _p(scalars), scalars.shape[0],
                                    first, count, _p(out), nt) == 0
    return out


def normalize(group, xyz):
    """into_affine: returns (xy limbs (2, k*12), infinity flag)"""
    k = GROUP_K[group]
    out = point_op(group, 3, xyz)
    return out[:2 * k * 12].reshape(2, k * 12), bool(out[2 * k * 12])


def walk(group, p0_xy, d_xy, n, nthreads=None):
    """bases P_i = P_0 + i*D (affine, Montgomery limbs) - the large-input generator"""
    k = GROUP_K[group]
    p0 = np.ascontiguousarray(p0_xy, dtype=np.uint64).reshape(-1)
    d = np.ascontiguousarray(d_xy, dtype=np.uint64).reshape(-1)
    out = np.zeros((n, 2 * k * 12), dtype=np.uint64)
    assert lib().ref753_walk(group, _p(p0), _p(d), n, _p(out), nthreads or hardware_threads()) == 0
    return out


def fft(field, data, mode, nthreads=None):
    """EvaluationDomain::{fft, ifft, coset_fft, coset_ifft} restated (best_fft split included)"""
    a = np.array(data, dtype=np.uint64, copy=True).reshape(-1, 12)
    n = a.shape[0]
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    rc = lib().ref753_fft(field, _p(a), log_n, mode, nthreads or hardware_threads())
    if rc == 4:
        return None
    assert rc == 0
    return a
