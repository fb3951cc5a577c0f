"""Pin the Python oracle against the reference's own known-answer tests and constants.

The vectors in tests/golden/reference_kat.json / reference_params.json were extracted from
the reference's test sources by tests/golden/make_golden.py (SURVEY.md 8c lists them).
Each test below replays one reference #[test] with the oracle in place of ginger-lib.
"""
import json
import os

import pytest

from oracle import g753 as O

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "reference_kat.json")))
PAR = json.load(open(os.path.join(HERE, "golden", "reference_params.json")))

FIELD_SETS = [
    ("fields_mnt4753_tests", "fields_mnt4753_fq", O.MNT4_FQ),
    ("fields_mnt6753_tests", "fields_mnt6753_fq", O.MNT6_FQ),
]


def vals(file_key, scope):
    return [(it["wrap"], int(it["value"], 16)) for it in KAT[file_key]["tests"][scope]]


def canon(field, item):
    """Canonical integer of a literal, whatever wrapper it was written with."""
    wrap, v = item
    return field.from_mont(v) if wrap == "mont" else v


def const(file_key, name, idx=0):
    it = PAR[file_key]["consts"][name][idx]
    return it["wrap"], int(it["value"], 16)


# ---------------------------------------------------------------- constants
@pytest.mark.parametrize("tests_key,par_key,F", FIELD_SETS)
def test_field_constants(tests_key, par_key, F):
    c = PAR[par_key]["consts"]
    ints = PAR[par_key]["ints"]
    assert int(c["MODULUS"][0]["value"], 16) == F.p
    assert int(c["R"][0]["value"], 16) == F.R
    assert int(c["R2"][0]["value"], 16) == F.R2
    assert ints["INV"] == F.inv64
    assert ints["MODULUS_BITS"] == F.bits == 753
    assert ints["TWO_ADICITY"] == F.two_adicity
    assert F.from_mont(int(c["GENERATOR"][0]["value"], 16)) == F.generator == 17
    assert F.from_mont(int(c["ROOT_OF_UNITY"][0]["value"], 16)) == F.root_of_unity
    assert int(c["T"][0]["value"], 16) == F.t
    # exact order 2^s
    assert pow(F.root_of_unity, 1 << F.two_adicity, F.p) == 1
    assert pow(F.root_of_unity, 1 << (F.two_adicity - 1), F.p) == F.p - 1


def test_curve_constants():
    for key, curve, F in (("curves_mnt4753_g1", O.MNT4_G1, O.MNT4_FQ), ("curves_mnt6753_g1", O.MNT6_G1, O.MNT6_FQ)):
        assert canon(F, const(key, "COEFF_A")) == curve.a[0]
        assert canon(F, const(key, "COEFF_B")) == curve.b[0]
        g = ((canon(F, const(key, "G1_GENERATOR_X")),), (canon(F, const(key, "G1_GENERATOR_Y")),))
        assert curve.on_curve(g)
        assert curve.mul(g, curve.r) is None
    F = O.MNT4_FQ
    assert canon(F, const("fields_mnt4753_fq2", "NONRESIDUE")) == 13
    assert canon(F, const("curves_mnt4753_g2", "MUL_BY_A_C0")) == 26
    assert canon(F, const("curves_mnt4753_g2", "MUL_BY_A_C1")) == 26
    assert tuple(canon(F, const("curves_mnt4753_g2", "COEFF_B", i)) for i in range(2)) == O.MNT4_G2.b
    g = (tuple(canon(F, const("curves_mnt4753_g2", "G2_GENERATOR_X_C%d" % i)) for i in range(2)),
         tuple(canon(F, const("curves_mnt4753_g2", "G2_GENERATOR_Y_C%d" % i)) for i in range(2)))
    assert O.MNT4_G2.on_curve(g)
    assert O.MNT4_G2.mul(g, O.MNT4_G2.r) is None
    F = O.MNT6_FQ
    assert canon(F, const("fields_mnt6753_fq3", "NONRESIDUE")) == 11
    assert canon(F, const("curves_mnt6753_g2", "MUL_BY_A_C0")) == 121
    assert canon(F, const("curves_mnt6753_g2", "MUL_BY_A_C1")) == 121
    assert canon(F, const("curves_mnt6753_g2", "MUL_BY_A_C2")) == 11
    assert tuple(canon(F, const("curves_mnt6753_g2", "COEFF_B", i)) for i in range(3)) == O.MNT6_G2.b
    g = (tuple(canon(F, const("curves_mnt6753_g2", "G2_GENERATOR_X_C%d" % i)) for i in range(3)),
         tuple(canon(F, const("curves_mnt6753_g2", "G2_GENERATOR_Y_C%d" % i)) for i in range(3)))
    assert O.MNT6_G2.on_curve(g)
    assert O.MNT6_G2.mul(g, O.MNT6_G2.r) is None


def test_byte_fixture():
    # fields/mnt{4,6}753/test_vec/*_tobyte: 96 little-endian bytes of a canonical element
    for name, F in (("mnt4753", O.MNT4_FQ), ("mnt6753", O.MNT6_FQ)):
        raw = bytes.fromhex(open(os.path.join(HERE, "golden", name + "_tobyte.hex")).read().strip())
        assert len(raw) == 96
        x = O.bytes96_to_int(raw)
        assert x < F.p
        assert O.int_to_bytes96(x) == raw


# ---------------------------------------------------------------- prime field KATs (raw Montgomery limbs)
@pytest.mark.parametrize("tests_key,par_key,F", FIELD_SETS)
def test_fq_mul_assign(tests_key, par_key, F):
    (wa, a), (wb, b), (wc, c) = vals(tests_key, "test_fq_mul_assign")
    assert wa == wb == wc == "mont"
    assert F.mont_mul(a, b) == c


@pytest.mark.parametrize("tests_key,par_key,F", FIELD_SETS)
def test_fq_squaring(tests_key, par_key, F):
    a, c = vals(tests_key, "test_fq_squaring")
    assert a[0] == "mont" and c[0] == "canon"
    assert F.from_mont(F.mont_mul(a[1], a[1])) == c[1]


@pytest.mark.parametrize("tests_key,par_key,F", FIELD_SETS)
def test_fq_add_assign(tests_key, par_key, F):
    v = [x for _, x in vals(tests_key, "test_fq_add_assign")]
    # shape: tmp, +0 -> e, +1 -> e, +rnd -> e, tmp=(q-1), +1 -> 0, tmp, +rnd -> e(q-1), +1 -> 0
    tmp, zero, e0, one, e1, rnd, e2, qm1, one2, t3, r3, e3, one3 = v
    assert zero == 0 and one == one2 == one3 == 1
    assert F.add(tmp, 0) == e0
    assert F.add(e0, 1) == e1
    assert F.add(e1, rnd) == e2
    assert qm1 == F.p - 1 and F.add(qm1, 1) == 0
    assert F.add(t3, r3) == e3 == F.p - 1


@pytest.mark.parametrize("tests_key,par_key,F", FIELD_SETS)
def test_fq_sub_assign(tests_key, par_key, F):
    v = [x for _, x in vals(tests_key, "test_fq_sub_assign")]
    a0, b0, e0, a1, b1, e1, z0, z1, a2, z2, e2 = v
    assert F.sub(a0, b0) == e0
    assert F.sub(a1, b1) == e1
    assert z0 == z1 == z2 == 0
    assert F.sub(a2, 0) == e2


# ---------------------------------------------------------------- Fq2 / Fq3 KATs (canonical)
def ext_vals(tests_key, scope, k):
    v = [x for _, x in vals(tests_key, scope)]
    assert len(v) % k == 0
    return [tuple(v[i:i + k]) for i in range(0, len(v), k)]


EXT_SETS = [("fields_mnt4753_tests", "fq2", O.FQ2_MNT4), ("fields_mnt6753_tests", "fq3", O.FQ3_MNT6)]


@pytest.mark.parametrize("tests_key,name,E", EXT_SETS)
def test_ext_mul(tests_key, name, E):
    a, b, c = ext_vals(tests_key, "test_%s_mul" % name, E.k)
    assert E.mul(a, b) == c


@pytest.mark.parametrize("tests_key,name,E", EXT_SETS)
def test_ext_squaring(tests_key, name, E):
    raw = [x for _, x in vals(tests_key, "test_%s_squaring" % name)]
    # the trailing 2k literals are (a, a^2); the leading ones are small-integer sanity cases
    a = tuple(raw[-2 * E.k:-E.k])
    c = tuple(raw[-E.k:])
    assert E.sqr(a) == c
    u = (0, 1) + (0,) * (E.k - 2)
    uk = u
    for _ in range(E.k - 1):
        uk = E.mul(uk, u)
    assert uk == E.from_int(E.nr)


@pytest.mark.parametrize("tests_key,name,E", EXT_SETS)
def test_ext_inverse(tests_key, name, E):
    a, c = ext_vals(tests_key, "test_%s_inverse" % name, E.k)
    assert E.inv(a) == c
    assert E.mul(a, c) == E.one()


@pytest.mark.parametrize("tests_key,name,E", EXT_SETS)
def test_ext_add_sub_neg_double(tests_key, name, E):
    a, b, c = ext_vals(tests_key, "test_%s_addition" % name, E.k)
    assert E.add(a, b) == c
    a, b, c = ext_vals(tests_key, "test_%s_subtraction" % name, E.k)
    assert E.sub(a, b) == c
    a, c = ext_vals(tests_key, "test_%s_negation" % name, E.k)
    assert E.neg(a) == c
    a, c = ext_vals(tests_key, "test_%s_doubling" % name, E.k)
    assert E.add(a, a) == c


# ---------------------------------------------------------------- curve KATs (canonical coordinates)
CURVE_SETS = [
    ("curves_mnt4753_tests", "g1", O.MNT4_G1),
    ("curves_mnt4753_tests", "g2", O.MNT4_G2),
    ("curves_mnt6753_tests", "g1", O.MNT6_G1),
    ("curves_mnt6753_tests", "g2", O.MNT6_G2),
]


def coords(tests_key, scope, k):
    v = [x for _, x in vals(tests_key, scope)]
    assert len(v) % k == 0
    return [tuple(v[i:i + k]) for i in range(0, len(v), k)]


@pytest.mark.parametrize("tests_key,g,C", CURVE_SETS)
def test_curve_addition(tests_key, g, C):
    x1, y1, z1, x2, y2, z2, ex, ey = coords(tests_key, "test_%s_addition_correctness" % g, C.F.k)
    P = C.from_projective(x1, y1, z1)
    Q = C.from_projective(x2, y2, z2)
    assert C.on_curve(P) and C.on_curve(Q)
    assert C.add(P, Q) == (ex, ey)


@pytest.mark.parametrize("tests_key,g,C", CURVE_SETS)
def test_curve_doubling(tests_key, g, C):
    x1, y1, z1, ex, ey = coords(tests_key, "test_%s_doubling_correctness" % g, C.F.k)
    P = C.from_projective(x1, y1, z1)
    assert C.double(P) == (ex, ey)
    assert C.add(P, P) == (ex, ey)


@pytest.mark.parametrize("tests_key,g,C", CURVE_SETS)
def test_curve_conversion(tests_key, g, C):
    x1, y1, z1, ex, ey = coords(tests_key, "test_%s_affine_projective_conversion" % g, C.F.k)
    assert C.from_projective(x1, y1, z1) == (ex, ey)


@pytest.mark.parametrize("tests_key,C", [("curves_mnt4753_tests", O.MNT4_G1), ("curves_mnt6753_tests", O.MNT6_G1)])
def test_g1_scalar_multiplication(tests_key, C):
    v = [x for _, x in vals(tests_key, "test_g1_scalar_multiplication")]
    x, y, s, ex, ey = v
    assert C.mul(((x,), (y,)), s) == ((ex,), (ey,))


# ---------------------------------------------------------------- L2 properties the reference tests state
def test_pippenger_semantics_equal_naive_sum():
    """variable_base.rs:114-151: MSM == naive sum, including bases.len() == scalars.len()+1,
    plus the special cases the reference code branches on (0, 1, infinity, duplicates)."""
    C = O.MNT4_G1
    rng = O.SplitMix64(0x5EED0001)
    pts = [O.random_g1_point(rng, C) for _ in range(6)]
    srng = O.SplitMix64(0x5EED0002)
    bases = [pts[0], pts[1], None, pts[2], pts[2], pts[3], pts[4], pts[5], pts[0]]
    scalars = [O.random_field_element(srng, O.MNT4_FR), 0, 5, 1, 1, C.r - 1, 7, O.random_field_element(srng, O.MNT4_FR)]
    assert len(bases) == len(scalars) + 1
    assert O.msm_pippenger_ref(C, bases, scalars) == O.msm_naive(C, bases, scalars)
    # >= 32 scalars switches the window rule (c = ceil(2/3 log2 n + 2))
    bases = [pts[i % 6] for i in range(40)]
    scalars = [O.random_field_element(srng, O.MNT4_FR) >> (13 * (i % 7)) for i in range(40)]
    assert O.ref_window_size(40) == 6
    assert O.msm_pippenger_ref(C, bases, scalars) == O.msm_naive(C, bases, scalars)
    assert O.msm_pippenger_ref(C, [], []) is None


def test_window_rule():
    # c = 13 / 16 / 17 at 2^16 / 2^20 / 2^22; < 32 scalars -> 3
    assert [O.ref_window_size(n) for n in (31, 32, 1 << 16, 1 << 20, 1 << 22)] == [3, 6, 13, 16, 17]


@pytest.mark.parametrize("F", [O.MNT4_FR, O.MNT6_FR])
def test_fft_is_dft_and_roundtrips(F):
    """fft/test.rs:9-43 round trips + serial_fft == the DFT definition, natural order."""
    rng = O.SplitMix64(0x5EED0003)
    for log_n in (0, 1, 3, 5):
        n = 1 << log_n
        d = O.EvaluationDomain(F, n)
        a = [O.random_field_element(rng, F) for _ in range(n)]
        assert d.fft(a) == O.dft_naive(a, d.group_gen, F.p)
        assert d.ifft(d.fft(a)) == a
        assert d.coset_ifft(d.coset_fft(a)) == a
        # coset_fft evaluates on g*<omega>
        ev = d.coset_fft(a)
        x = (d.generator * pow(d.group_gen, 3 % n, F.p)) % F.p
        assert ev[3 % n] == sum(c * pow(x, j, F.p) for j, c in enumerate(a)) % F.p
    # short input is zero-padded (domain.rs:121)
    d = O.EvaluationDomain(F, 8)
    a = [O.random_field_element(rng, F) for _ in range(5)]
    assert d.fft(a) == d.fft(a + [0, 0, 0])


def test_domain_limits():
    # domain.rs:70-72 and SURVEY.md F2: mnt6753::Fr tops out at 2^14, mnt4753::Fr at 2^29
    assert O.EvaluationDomain.try_new(O.MNT6_FR, 1 << 14) is not None
    assert O.EvaluationDomain.try_new(O.MNT6_FR, (1 << 14) + 1) is None
    assert O.EvaluationDomain.try_new(O.MNT4_FR, 1 << 29) is not None
    assert O.EvaluationDomain.try_new(O.MNT4_FR, (1 << 29) + 1) is None
