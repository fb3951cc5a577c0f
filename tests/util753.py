"""Shared helpers for the tests: oracle objects <-> the C ABI's numpy limb layout."""
import importlib
import json
import os

import numpy as np

from oracle import g753 as O

HERE = os.path.dirname(os.path.abspath(__file__))
G = importlib.import_module("ginger-lib_b200")
ffi = G.ffi

GROUPS = {
    ffi.MNT4_G1: O.MNT4_G1,
    ffi.MNT4_G2: O.MNT4_G2,
    ffi.MNT6_G1: O.MNT6_G1,
    ffi.MNT6_G2: O.MNT6_G2,
}
FIELDS = {ffi.FIELD_MNT6_FR: O.MNT6_FR, ffi.FIELD_MNT4_FR: O.MNT4_FR}


def int_to_limbs(x):
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(12)]


def ints_to_array(vals):
    """list of ints -> (n, 12) uint64"""
    a = np.zeros((len(vals), 12), dtype=np.uint64)
    for r, v in enumerate(vals):
        a[r] = int_to_limbs(v)
    return a


def array_to_ints(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 12)
    return [sum(int(a[r, i]) << (64 * i) for i in range(12)) for r in range(a.shape[0])]


def points_to_arrays(curve, pts):
    """affine oracle points (None = infinity) -> (coords (n, 2k*12) Montgomery, infinity (n,))"""
    F = curve.F.base
    k = curve.F.k
    coords = np.zeros((len(pts), 2 * k * 12), dtype=np.uint64)
    inf = np.zeros(len(pts), dtype=np.uint8)
    for r, P in enumerate(pts):
        if P is None:
            inf[r] = 1
            # the reference's GroupAffine::zero() is (0, 1, infinity=true); any payload is legal
            coords[r, k * 12:(k + 1) * 12] = int_to_limbs(F.to_mont(1))
            continue
        flat = [F.to_mont(c) for c in P[0]] + [F.to_mont(c) for c in P[1]]
        coords[r] = np.concatenate([np.array(int_to_limbs(v), dtype=np.uint64) for v in flat])
    return coords, inf


def projective_to_point(curve, xyz):
    """(3, k*12) Montgomery limbs -> normalised affine oracle point; checks canonical limbs"""
    F = curve.F.base
    k = curve.F.k
    vals = array_to_ints(np.asarray(xyz).reshape(-1, 12))
    assert len(vals) == 3 * k
    for v in vals:
        assert v < F.p, "non-canonical coordinate returned"
    c = [F.from_mont(v) for v in vals]
    X, Y, Z = tuple(c[:k]), tuple(c[k:2 * k]), tuple(c[2 * k:])
    return curve.from_projective(X, Y, Z)


_PARAMS = None


def g2_generator(curve):
    global _PARAMS
    if _PARAMS is None:
        _PARAMS = json.load(open(os.path.join(HERE, "golden", "reference_params.json")))
    key = "curves_mnt4753_g2" if curve is O.MNT4_G2 else "curves_mnt6753_g2"
    F = curve.F.base
    k = curve.F.k
    gx = tuple(F.from_mont(int(_PARAMS[key]["consts"]["G2_GENERATOR_X_C%d" % i][0]["value"], 16)) for i in range(k))
    gy = tuple(F.from_mont(int(_PARAMS[key]["consts"]["G2_GENERATOR_Y_C%d" % i][0]["value"], 16)) for i in range(k))
    return (gx, gy)


def sample_points(curve, count, seed):
    """distinct random-looking points: x-sampling on G1 (cofactor 1), an additive walk from
    the reference generator on G2 (cofactor != 1)"""
    rng = O.SplitMix64(seed)
    if curve.F.k == 1:
        return [O.random_g1_point(rng, curve) for _ in range(count)]
    g = g2_generator(curve)
    steps = [curve.mul(g, (rng.next() << 64) | rng.next() | 1) for _ in range(4)]
    pts, cur = [], curve.mul(g, rng.next() | 1)
    for i in range(count):
        pts.append(cur)
        cur = curve.add(cur, steps[i % 4])
    return pts


def sample_scalars(curve, count, seed):
    rng = O.SplitMix64(seed)
    fr = O.MNT4_FR if curve.r == O.MNT4_FR.p else O.MNT6_FR
    return [O.random_field_element(rng, fr) for _ in range(count)]


def field_array(field, vals):
    """canonical ints -> Montgomery limb array"""
    return ints_to_array([field.to_mont(v) for v in vals])


def array_field(field, a):
    out = []
    for v in array_to_ints(a):
        assert v < field.p, "non-canonical element returned"
        out.append(field.from_mont(v))
    return out


def build_emul(flags=(), tag=""):
    """(re)build the TEST-ONLY host-emulation library of the C ABI when it is missing or stale;
    `flags` / `tag`: a differently configured build (e.g. the opt-in kernel forms) under its own name"""
    import subprocess
    src = os.path.join(HERE, "host_emul", "capi_emul.cpp")
    lib = os.path.join(HERE, "host_emul", "libg753_emul%s.so" % tag)
    csrc = os.path.join(HERE, "..", "ginger-lib_b200", "csrc")
    deps = [src, os.path.join(HERE, "..", "include", "g753.h")] + [
        os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".cu", ".inc"))]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++"] + list(flags) +
                              ["-o", lib, src])
    return lib
