"""Pin the C++ restatement (oracle/ref753.cpp - the checker at large sizes and the timed CPU
"port") against the Python oracle and the reference's raw-limb KATs."""
import json
import os

import numpy as np
import pytest

from oracle import g753 as O
from oracle import ref753 as R
from util753 import (FIELDS, GROUPS, array_field, array_to_ints, field_array, ints_to_array, points_to_arrays,
                     projective_to_point, sample_points, sample_scalars)

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "reference_kat.json")))
PAR = json.load(open(os.path.join(HERE, "golden", "reference_params.json")))


@pytest.mark.parametrize("field,key,pkey,F", [(0, "fields_mnt4753_tests", "fields_mnt4753_fq", O.MNT4_FQ),
                                              (1, "fields_mnt6753_tests", "fields_mnt6753_fq", O.MNT6_FQ)])
def test_constants_and_field_kats(field, key, pkey, F):
    import ctypes
    out = np.zeros(12, dtype=np.uint64)
    c = PAR[pkey]["consts"]
    for which, name in ((0, "R"), (1, "R2"), (2, "ROOT_OF_UNITY")):
        assert R.lib().ref753_constant(field, which, out.ctypes.data_as(ctypes.c_void_p)) == 0
        assert array_to_ints(out)[0] == int(c[name][0]["value"], 16)
    assert R.lib().ref753_constant(field, 3, out.ctypes.data_as(ctypes.c_void_p)) == 0
    assert int(out[0]) == PAR[pkey]["ints"]["INV"]
    t = KAT[key]["tests"]
    a, b, m = [int(x["value"], 16) for x in t["test_fq_mul_assign"]]
    assert array_to_ints(R.field_op(field, 0, ints_to_array([a]), ints_to_array([b]))) == [m]
    a, s = [int(x["value"], 16) for x in t["test_fq_squaring"]]
    assert array_to_ints(R.field_op(field, 3, ints_to_array([a]))) == [F.to_mont(s)]
    rng = O.SplitMix64(41 + field)
    xs = [0, 1, F.p - 1] + [O.random_field_element(rng, F) for _ in range(200)]
    ys = [F.p - 1, 0, 1] + [O.random_field_element(rng, F) for _ in range(200)]
    A, B = ints_to_array(xs), ints_to_array(ys)
    assert array_to_ints(R.field_op(field, 0, A, B)) == [F.mont_mul(x, y) for x, y in zip(xs, ys)]
    assert array_to_ints(R.field_op(field, 1, A, B)) == [(x + y) % F.p for x, y in zip(xs, ys)]
    assert array_to_ints(R.field_op(field, 2, A, B)) == [(x - y) % F.p for x, y in zip(xs, ys)]
    assert array_to_ints(R.field_op(field, 4, A)) == [(-x) % F.p for x in xs]
    assert array_to_ints(R.field_op(field, 5, A[1:20])) == [F.to_mont(F.inv(F.from_mont(x))) for x in xs[1:20]]
    assert array_to_ints(R.field_op(field, 6, A)) == [F.to_mont(x) for x in xs]
    assert array_to_ints(R.field_op(field, 7, A)) == [F.from_mont(x) for x in xs]


@pytest.mark.parametrize("ext,E", [(2, O.FQ2_MNT4), (3, O.FQ3_MNT6)])
def test_ext_ops(ext, E):
    F = E.base
    rng = O.SplitMix64(0xE0 + ext)
    for _ in range(6):
        a = tuple(O.random_field_element(rng, F) for _ in range(E.k))
        b = tuple(O.random_field_element(rng, F) for _ in range(E.k))
        A, B = field_array(F, a), field_array(F, b)
        assert tuple(array_field(F, R.ext_op(ext, 0, A, B))) == E.mul(a, b)
        assert tuple(array_field(F, R.ext_op(ext, 3, A))) == E.sqr(a)
        assert tuple(array_field(F, R.ext_op(ext, 5, A))) == E.inv(a)
        assert tuple(array_field(F, R.ext_op(ext, 1, A, B))) == E.add(a, b)
        assert tuple(array_field(F, R.ext_op(ext, 2, A, B))) == E.sub(a, b)


def proj_of(C, P, z=None):
    """homogeneous projective limbs (X, Y, Z) of an affine oracle point, optional scaling z"""
    E, F = C.F, C.F.base
    if P is None:
        t = [E.zero(), E.one(), E.zero()]
    else:
        z = z or E.one()
        t = [E.mul(P[0], z), E.mul(P[1], z), z]
    return field_array(F, [c for el in t for c in el])


@pytest.mark.parametrize("group", sorted(GROUPS))
def test_group_law(group):
    """the reference's projective formulas (dbl-2007-bl, madd-1998-cmo, add-1998-cmo-2) incl.
    their special cases, against the affine oracle; replays the shape of the curve KATs"""
    C = GROUPS[group]
    E, F = C.F, C.F.base
    P, Q = sample_points(C, 2, 0xC0 + group)
    rng = O.SplitMix64(3)
    z = tuple(O.random_field_element(rng, F) for _ in range(E.k))
    for A, B in [(P, Q), (P, P), (P, C.neg(P)), (None, Q), (P, None), (None, None)]:
        got = R.point_op(group, 0, proj_of(C, A, z), proj_of(C, B))[:3 * E.k * 12]
        assert projective_to_point(C, got) == C.add(A, B)
        if B is not None:
            aff, _ = points_to_arrays(C, [B])
            got = R.point_op(group, 4, proj_of(C, A, z), aff)[:3 * E.k * 12]
            assert projective_to_point(C, got) == C.add(A, B)
        got = R.point_op(group, 1, proj_of(C, A, z))[:3 * E.k * 12]
        assert projective_to_point(C, got) == C.double(A)
    s = O.random_field_element(rng, F) % C.r
    aff, _ = points_to_arrays(C, [P])
    got = R.point_op(group, 2, aff, ints_to_array([s]))[:3 * E.k * 12]
    assert projective_to_point(C, got) == C.mul(P, s)
    xy, inf = R.normalize(group, proj_of(C, P, z))
    assert not inf and tuple(array_field(F, xy.reshape(-1, 12))) == tuple(P[0]) + tuple(P[1])


@pytest.mark.parametrize("group", sorted(GROUPS))
def test_msm_vs_python(group):
    C = GROUPS[group]
    n = 40 if C.F.k == 1 else 12
    pts = sample_points(C, n, 0xA0 + group)
    sc = sample_scalars(C, n, 0xB0 + group)
    pts[1] = None
    sc[2], sc[3], sc[4] = 0, 1, C.r - 1
    pts[6], sc[6] = pts[5], sc[5]
    pts[7], sc[7] = C.neg(pts[5]), sc[5]
    coords, inf = points_to_arrays(C, pts)
    got = projective_to_point(C, R.msm(group, coords, inf, ints_to_array(sc)))
    assert got == O.msm_naive(C, pts, sc)
    got = projective_to_point(C, R.msm(group, coords, inf, ints_to_array(sc[:9]), nthreads=1))   # c = 3 branch
    assert got == O.msm_naive(C, pts, sc[:9])
    got = projective_to_point(C, R.msm(group, coords[:5], inf[:5], ints_to_array(sc)))           # zip-truncation
    assert got == O.msm_naive(C, pts[:5], sc)


@pytest.mark.parametrize("group", [0, 3])
def test_msm_window_sample(group):
    """msm_windows (the bounded sample the CPU baseline times): the per-window sums, folded high to low with
    c doublings in between (variable_base.rs:72-82), are multi_scalar_mul"""
    C = GROUPS[group]
    n = 40 if C.F.k == 1 else 33
    pts = sample_points(C, n, 0xA7 + group)
    sc = sample_scalars(C, n, 0xB7 + group)
    sc[3] = 1
    coords, inf = points_to_arrays(C, pts)
    arr = ints_to_array(sc)
    c = R.msm_window_bits(n)
    windows = (753 + c - 1) // c
    lo = [projective_to_point(C, w) for w in R.msm_windows(group, coords, inf, arr, 0, 5)]
    hi = [projective_to_point(C, w) for w in R.msm_windows(group, coords, inf, arr, 5, windows - 5, nthreads=2)]
    total = None
    for w in reversed(lo + hi):
        for _ in range(c):
            total = C.double(total) if total is not None else None
        total = C.add(total, w)
    assert total == O.msm_naive(C, pts, sc)


def test_walk_generator():
    for group in (0, 1):
        C = GROUPS[group]
        P0, D = sample_points(C, 2, 0x99 + group)
        c0, _ = points_to_arrays(C, [P0])
        cd, _ = points_to_arrays(C, [D])
        got = R.walk(group, c0, cd, 5000, nthreads=3)
        F = C.F.base
        k = C.F.k
        for i in (0, 1, 2, 4095, 4096, 4999):
            want = C.add(P0, C.mul(D, i))
            v = array_field(F, got[i].reshape(-1, 12))
            assert (tuple(v[:k]), tuple(v[k:])) == want


@pytest.mark.parametrize("field", sorted(FIELDS))
def test_fft_vs_python(field):
    F = FIELDS[field]
    rng = O.SplitMix64(0xF0 + field)
    for log_n, threads in ((0, 1), (3, 1), (6, 4), (8, 16)):
        n = 1 << log_n
        a = [O.random_field_element(rng, F) for _ in range(n)]
        ref = O.EvaluationDomain(F, n)
        arr = field_array(F, a)
        assert array_field(F, R.fft(field, arr, 0, threads)) == ref.fft(a)
        assert array_field(F, R.fft(field, arr, 1, threads)) == ref.ifft(a)
        assert array_field(F, R.fft(field, arr, 2, threads)) == ref.coset_fft(a)
        assert array_field(F, R.fft(field, arr, 3, threads)) == ref.coset_ifft(a)
    if field == 0:
        assert R.fft(field, np.zeros((1 << 15, 12), dtype=np.uint64), 0, 1) is None   # EvaluationDomain::new -> None
