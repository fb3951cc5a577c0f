"""Check the device limb arithmetic (csrc/fq.cuh, slots.cuh, ec_slots.cuh) on the CPU.

The PTX carry-chain primitives are emulated on the host (tests/host_emul/emul.cpp), so this
exercises exactly the limb logic the GPU runs - minus ptxas.  It is a test of kernel source,
not a product path: the product library has no host implementation.
"""
import ctypes
import json
import os
import subprocess

import pytest

from oracle import g753 as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "emul.cpp")
LIB = os.path.join(HERE, "host_emul", "libemul.so")
CSRC = os.path.join(HERE, "..", "ginger-lib_b200", "csrc")


# the default build and the opt-in rolled multiplier (slots.cuh, G753_ROLLED): identical results
@pytest.fixture(scope="module", params=["", "_rolled"])
def emul(request):
    lib = LIB.replace(".so", request.param + ".so")
    flags = ["-DG753_ROLLED=1"] if request.param else []
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("fq.cuh", "slots.cuh", "ec_slots.cuh", "device.cuh", "constants.inc")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC"] + flags + ["-o", lib, SRC])
    return ctypes.CDLL(lib)


U32x24 = ctypes.c_uint32 * 24


def to_buf(vals):
    """list of ints (768-bit each) -> contiguous u32 buffer"""
    buf = (ctypes.c_uint32 * (24 * len(vals)))()
    for k, v in enumerate(vals):
        for i in range(24):
            buf[24 * k + i] = (v >> (32 * i)) & 0xFFFFFFFF
    return buf


def from_buf(buf, count):
    return [sum(buf[24 * k + i] << (32 * i) for i in range(24)) for k in range(count)]


FIELDS = [(0, O.MNT4_FQ), (1, O.MNT6_FQ)]


@pytest.mark.parametrize("fid,F", FIELDS)
def test_field_ops(emul, fid, F):
    rng = O.SplitMix64(0xF1E1D + fid)
    p = F.p
    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, (1 << 752), (1 << 752) - 1, F.R, F.R2, 0xFFFFFFFF,
            (1 << 64) - 1, p >> 1, p - (1 << 32), p - (1 << 64) + 1]
    vals = edge + [O.random_field_element(rng, F) for _ in range(60)]
    pairs = [(a, b) for a in edge for b in edge] + list(zip(vals, reversed(vals)))
    out = U32x24()

    def run(op, a, b=0):
        emul.emul_field_op(fid, op, to_buf([a]), to_buf([b]), out)
        return from_buf(out, 1)[0]

    for a, b in pairs:
        assert run(0, a, b) == F.mont_mul(a, b), (hex(a), hex(b))
        assert run(13, a, b) == F.mont_mul(a, b), (hex(a), hex(b))
        assert run(1, a, b) == (a + b) % p
        assert run(2, a, b) == (a - b) % p
    for a in vals:
        assert run(3, a) == F.mont_mul(a, a)
        assert run(14, a) == F.mont_mul(a, a)
        assert run(4, a) == (-a) % p
        assert run(6, a) == F.to_mont(a)
        assert run(7, a) == F.from_mont(a)
        assert run(8, a) == 11 * a % p
        assert run(9, a) == 13 * a % p
        assert run(10, a) == 26 * a % p
        assert run(11, a) == 121 * a % p
        assert run(12, a) == 2 * a % p
    for a in vals[1:8] + vals[-4:]:
        if a % p:
            x = F.from_mont(a)  # fq_inv maps xR -> x^-1 R
            assert run(5, a) == F.to_mont(F.inv(x))


@pytest.mark.parametrize("fid,F", FIELDS)
def test_two_products_one_reduction(emul, fid, F):
    """fq_mul2 / s_mul2: (a b + c d) / R mod p with a single Montgomery reduction, canonical output"""
    rng = O.SplitMix64(0x2B0D + fid)
    p = F.p
    edge = [0, 1, p - 1, p - 2, (1 << 752), F.R, (1 << 752) - 1, p - (1 << 32)]
    quads = [(a, b, c, d) for a in edge[:4] for b in edge[2:6] for c in edge[1:5] for d in edge[3:]]
    quads += [(p - 1, p - 1, p - 1, p - 1), (0, 0, 0, 0), (p - 1, p - 1, 0, 5), (0, 7, p - 1, p - 1)]
    quads += [tuple(O.random_field_element(rng, F) for _ in range(4)) for _ in range(100)]
    out = U32x24()
    for a, b, c, d in quads:
        want = (F.mont_mul(a, b) + F.mont_mul(c, d)) % p
        emul.emul_field_mul2(fid, -1, to_buf([a]), to_buf([b]), to_buf([c]), to_buf([d]), out)
        assert from_buf(out, 1)[0] == want
        emul.emul_field_mul2(fid, 0, to_buf([a]), to_buf([b]), to_buf([c]), to_buf([d]), out)
        assert from_buf(out, 1)[0] == want
        emul.emul_field_mul2(fid, 1, to_buf([a]), to_buf([b]), to_buf([c]), to_buf([d]), out)
        assert from_buf(out, 1)[0] == (F.mont_mul(a, b) - F.mont_mul(c, d)) % p
        emul.emul_field_mul2(fid, 2, to_buf([a]), to_buf([b]), to_buf([c]), to_buf([d]), out)
        assert from_buf(out, 1)[0] == (F.mont_mul(a, b) + 13 * F.mont_mul(c, d)) % p


def test_reference_mul_kat_through_device_code(emul):
    """The reference's raw-limb multiplication KAT (fields/mnt4753/tests.rs:605-650,
    fields/mnt6753/tests.rs:811) through the device multiplier source."""
    kat = json.load(open(os.path.join(HERE, "golden", "reference_kat.json")))
    out = U32x24()
    for fid, key in ((0, "fields_mnt4753_tests"), (1, "fields_mnt6753_tests")):
        a, b, c = [int(it["value"], 16) for it in kat[key]["tests"]["test_fq_mul_assign"]]
        emul.emul_field_op(fid, 0, to_buf([a]), to_buf([b]), out)
        assert from_buf(out, 1)[0] == c


EXTS = [(2, O.FQ2_MNT4), (3, O.FQ3_MNT6)]


@pytest.mark.parametrize("ext,E", EXTS)
def test_ext_ops(emul, ext, E):
    F = E.base
    k = E.k
    rng = O.SplitMix64(0xE87 + ext)
    out = (ctypes.c_uint32 * (24 * k))()

    def mont(t):
        return [F.to_mont(c) for c in t]

    def run(op, a, b=None):
        b = b if b is not None else E.zero()
        emul.emul_ext_op(ext, op, to_buf(mont(a)), to_buf(mont(b)), out)
        return tuple(F.from_mont(v) for v in from_buf(out, k))

    elems = [E.zero(), E.one(), tuple([F.p - 1] * k), (0,) * (k - 1) + (1,)]
    elems += [tuple(O.random_field_element(rng, F) for _ in range(k)) for _ in range(12)]
    for a in elems:
        for b in elems[:6] + elems[-3:]:
            assert run(0, a, b) == E.mul(a, b)
            assert run(1, a, b) == E.add(a, b)
            assert run(2, a, b) == E.sub(a, b)
        assert run(3, a) == E.sqr(a)
        assert run(4, a) == E.neg(a)
        assert run(12, a) == E.add(a, a)
        assert run(23, a) == E.sqr(a)                 # destination aliases the operand
        for b in elems[1:3] + elems[-2:]:
            assert run(20, a, b) == E.mul(a, b)       # d aliases a
            assert run(21, a, b) == E.mul(a, b)       # d aliases b


CURVES = [(0, O.MNT4_G1), (1, O.MNT4_G2), (2, O.MNT6_G1), (3, O.MNT6_G2)]


def g2_generator(curve):
    par = json.load(open(os.path.join(HERE, "golden", "reference_params.json")))
    key = "curves_mnt4753_g2" if curve is O.MNT4_G2 else "curves_mnt6753_g2"
    F = curve.F.base
    k = curve.F.k
    gx = tuple(F.from_mont(int(par[key]["consts"]["G2_GENERATOR_X_C%d" % i][0]["value"], 16)) for i in range(k))
    gy = tuple(F.from_mont(int(par[key]["consts"]["G2_GENERATOR_Y_C%d" % i][0]["value"], 16)) for i in range(k))
    return (gx, gy)


def sample_points(curve, count, seed):
    rng = O.SplitMix64(seed)
    if curve.F.k == 1:
        return [O.random_g1_point(rng, curve) for _ in range(count)]
    g = g2_generator(curve)
    return [curve.mul(g, rng.next() | 1) for _ in range(count)]


@pytest.mark.parametrize("cid,C", CURVES)
def test_curve_ops(emul, cid, C):
    E = C.F
    F = E.base
    k = E.k
    pts = sample_points(C, 2, 0xC0 + cid)

    def enc_aff(P):
        if P is None:
            return [0] * (2 * k)
        return [F.to_mont(c) for c in P[0]] + [F.to_mont(c) for c in P[1]]

    def enc_xyzz(P, scale=None):
        """an XYZZ representative of P with a non-trivial Z (zz = s^2, zzz = s^3)"""
        if P is None:
            return [0] * k + [F.to_mont(1)] + [0] * (k - 1) + [0] * (2 * k)
        s = scale if scale is not None else E.one()
        zz = E.sqr(s)
        zzz = E.mul(zz, s)
        x = E.mul(P[0], zz)
        y = E.mul(P[1], zzz)
        return [F.to_mont(c) for t in (x, y, zz, zzz) for c in t]

    out = (ctypes.c_uint32 * (24 * 4 * k))()

    def dec_xyzz():
        v = [F.from_mont(x) for x in from_buf(out, 4 * k)]
        x, y, zz, zzz = (tuple(v[i * k:(i + 1) * k]) for i in range(4))
        if E.is_zero(zz):
            return None
        assert E.mul(E.sqr(zz), zz) == E.sqr(zzz)
        return (E.mul(x, E.inv(zz)), E.mul(y, E.inv(zzz)))

    rng = O.SplitMix64(77)
    scale = tuple(O.random_field_element(rng, F) for _ in range(k))
    P, Q = pts[0], pts[1]
    cases = [(P, Q), (P, P), (P, C.neg(P)), (None, Q), (P, None), (None, None), (C.double(P), P)]
    for A, B in cases:
        for sc in (None, scale):
            if B is not None:   # the mixed addition is only ever fed finite bases (k_msm_digits drops the rest)
                emul.emul_curve_op(cid, 0, to_buf(enc_xyzz(A, sc)), to_buf(enc_aff(B)), out)
                assert dec_xyzz() == C.add(A, B)
                emul.emul_curve_op(cid, 6, to_buf(enc_xyzz(A, sc)), to_buf(enc_aff(B)), out)
                assert dec_xyzz() == C.add(A, C.neg(B))
                # the accumulation kernel's form (six slots, fused products on the prime-field curves)
                emul.emul_curve_op(cid, 7, to_buf(enc_xyzz(A, sc)), to_buf(enc_aff(B)), out)
                assert dec_xyzz() == C.add(A, B)
                emul.emul_curve_op(cid, 8, to_buf(enc_xyzz(A, sc)), to_buf(enc_aff(B)), out)
                assert dec_xyzz() == C.add(A, C.neg(B))
            emul.emul_curve_op(cid, 1, to_buf(enc_xyzz(A, sc)), to_buf(enc_xyzz(B, scale)), out)
            assert dec_xyzz() == C.add(A, B)
            emul.emul_curve_op(cid, 5, to_buf(enc_xyzz(A, sc)), to_buf(enc_xyzz(B, scale)), out)
            assert dec_xyzz() == C.add(A, B)
        emul.emul_curve_op(cid, 2, to_buf(enc_xyzz(A, scale)), to_buf([0]), out)
        assert dec_xyzz() == C.double(A)
        # homogeneous projective output (GroupProjective layout)
        emul.emul_curve_op(cid, 3, to_buf(enc_xyzz(A, scale)), to_buf([0]), out)
        v = [F.from_mont(x) for x in from_buf(out, 3 * k)]
        X, Y, Z = (tuple(v[i * k:(i + 1) * k]) for i in range(3))
        assert C.from_projective(X, Y, Z) == A
        if A is None:
            assert (X, Y, Z) == (E.zero(), E.one(), E.zero())
