"""Parity of the CUDA path (through the C ABI of libg753.so) against the oracle on a real B200.

Small sizes are compared value-for-value with the Python oracle; the reference's own KATs are
replayed through the device code; larger sizes use size-independent properties (see
test_gpu_large.py for the 2^16+ cases checked against the C++ oracle).
"""
import json
import os

import numpy as np
import pytest

from oracle import g753 as O
from util753 import (FIELDS, G, GROUPS, array_field, array_to_ints, field_array, ffi, ints_to_array,
                     points_to_arrays, projective_to_point, sample_points, sample_scalars)

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "reference_kat.json")))


@pytest.fixture(scope="module")
def ctx():
    c = G.Context(0)           # raises if libg753.so or the GPU is missing: no CPU path
    yield c
    c.close()


def field_op(ctx, field, op, a, b=None):
    a = ints_to_array(a)
    bb = ints_to_array(b) if b is not None else None
    out = np.zeros_like(a)
    ctx.lib.check(ctx.lib.field_op(ctx.handle, field, op, ffi.ptr(a), ffi.ptr(bb), ffi.ptr(out), a.shape[0]))
    return array_to_ints(out)


@pytest.mark.parametrize("field,key,F", [(0, "fields_mnt4753_tests", O.MNT4_FQ), (1, "fields_mnt6753_tests", O.MNT6_FQ)])
def test_reference_field_kats_on_device(ctx, field, key, F):
    """fields/mnt4753/tests.rs:605-650 (mul), :697-730 (square), :271-600 (add/sub) and the
    mnt6753 twins, raw Montgomery limbs, through the device multiplier."""
    t = KAT[key]["tests"]
    a, b, c = [int(x["value"], 16) for x in t["test_fq_mul_assign"]]
    assert field_op(ctx, field, ffi.OP_MUL, [a], [b]) == [c]
    a, c = [int(x["value"], 16) for x in t["test_fq_squaring"]]
    assert field_op(ctx, field, ffi.OP_SQR, [a]) == [F.to_mont(c)]
    v = [int(x["value"], 16) for x in t["test_fq_add_assign"]]
    assert field_op(ctx, field, ffi.OP_ADD, [v[4], v[9]], [v[5], v[10]]) == [v[6], v[11]]
    v = [int(x["value"], 16) for x in t["test_fq_sub_assign"]]
    assert field_op(ctx, field, ffi.OP_SUB, [v[0], v[3]], [v[1], v[4]]) == [v[2], v[5]]


@pytest.mark.parametrize("field,F", [(0, O.MNT4_FQ), (1, O.MNT6_FQ)])
def test_field_ops_random(ctx, field, F):
    rng = O.SplitMix64(0x1234 + field)
    p = F.p
    edge = [0, 1, p - 1, p - 2, (p - 1) // 2, F.R, (1 << 752), p - (1 << 32)]
    a = edge + [O.random_field_element(rng, F) for _ in range(2000)]
    b = list(reversed(edge)) + [O.random_field_element(rng, F) for _ in range(2000)]
    assert field_op(ctx, field, ffi.OP_MUL, a, b) == [F.mont_mul(x, y) for x, y in zip(a, b)]
    assert field_op(ctx, field, ffi.OP_ADD, a, b) == [(x + y) % p for x, y in zip(a, b)]
    assert field_op(ctx, field, ffi.OP_SUB, a, b) == [(x - y) % p for x, y in zip(a, b)]
    assert field_op(ctx, field, ffi.OP_SQR, a) == [F.mont_mul(x, x) for x in a]
    assert field_op(ctx, field, ffi.OP_TO_MONT, a) == [F.to_mont(x) for x in a]
    assert field_op(ctx, field, ffi.OP_FROM_MONT, a) == [F.from_mont(x) for x in a]
    inv = field_op(ctx, field, ffi.OP_INV, a[1:40])
    assert inv == [F.to_mont(F.inv(F.from_mont(x))) for x in a[1:40]]


CURVE_KATS = [
    (ffi.MNT4_G1, "curves_mnt4753_tests", "g1"), (ffi.MNT4_G2, "curves_mnt4753_tests", "g2"),
    (ffi.MNT6_G1, "curves_mnt6753_tests", "g1"), (ffi.MNT6_G2, "curves_mnt6753_tests", "g2"),
]


def point_op(ctx, group, op, A, B=None, scalar=None):
    C = GROUPS[group]
    ca, _ = points_to_arrays(C, [A])
    if A is None:
        ca[:] = 0
    cb = None
    if op in (0, 3):
        cb, _ = points_to_arrays(C, [B])
        if B is None:
            cb[:] = 0
    elif op == 2:
        cb = ints_to_array([scalar])
    out = np.zeros((3, C.F.k * 12), dtype=np.uint64)
    ctx.lib.check(ctx.lib.point_op(ctx.handle, group, op, ffi.ptr(ca), ffi.ptr(cb), ffi.ptr(out)))
    return projective_to_point(C, out)


@pytest.mark.parametrize("group,key,g", CURVE_KATS)
def test_reference_curve_kats_on_device(ctx, group, key, g):
    """curves/mnt{4,6}753/tests.rs addition / doubling / scalar-multiplication KATs through the
    device group law."""
    C = GROUPS[group]
    k = C.F.k
    t = KAT[key]["tests"]

    def coords(scope):
        v = [int(x["value"], 16) for x in t[scope]]
        return [tuple(v[i:i + k]) for i in range(0, len(v), k)]

    x1, y1, z1, x2, y2, z2, ex, ey = coords("test_%s_addition_correctness" % g)
    P, Q = C.from_projective(x1, y1, z1), C.from_projective(x2, y2, z2)
    assert point_op(ctx, group, 0, P, Q) == (ex, ey)
    x1, y1, z1, ex, ey = coords("test_%s_doubling_correctness" % g)
    P = C.from_projective(x1, y1, z1)
    assert point_op(ctx, group, 1, P) == (ex, ey)
    assert point_op(ctx, group, 0, P, P) == (ex, ey)
    assert point_op(ctx, group, 0, P, C.neg(P)) is None
    assert point_op(ctx, group, 0, None, P) == P
    assert point_op(ctx, group, 0, P, None) == P
    # full (XYZZ + XYZZ) addition incl. its exceptional cases
    assert point_op(ctx, group, 3, P, Q) == C.add(C.double(P), C.double(Q))
    assert point_op(ctx, group, 3, P, P) == C.double(C.double(P))
    assert point_op(ctx, group, 3, P, C.neg(P)) is None
    assert point_op(ctx, group, 3, None, Q) == C.double(Q)
    if g == "g1":
        x, y, s, ex, ey = [int(v["value"], 16) for v in t["test_g1_scalar_multiplication"]]
        assert point_op(ctx, group, 2, ((x,), (y,)), scalar=s) == ((ex,), (ey,))


@pytest.mark.parametrize("group", sorted(GROUPS))
def test_msm_small_vs_oracle(ctx, group):
    C = GROUPS[group]
    n = 96 if C.F.k == 1 else 40
    pts = sample_points(C, n, 0xA0 + group)
    sc = sample_scalars(C, n, 0xB0 + group)
    pts[1] = None
    sc[2] = 0
    sc[3] = 1
    sc[4] = C.r - 1
    pts[6], sc[6] = pts[5], sc[5]
    pts[7], sc[7] = C.neg(pts[5]), sc[5]
    sc[8] = 3
    sc[9] = (1 << 752) - 1 if (1 << 752) - 1 < C.r else C.r - 2
    coords, inf = points_to_arrays(C, pts)
    bases = ctx.upload_bases(group, coords, inf)
    arr = ints_to_array(sc)
    got = G.VariableBaseMSM.multi_scalar_mul(bases, arr)
    assert projective_to_point(C, got) == O.msm_naive(C, pts, sc)
    got = G.VariableBaseMSM.multi_scalar_mul(bases, arr[:31])       # < 32 scalars (c = 3 in the reference)
    assert projective_to_point(C, got) == O.msm_naive(C, pts, sc[:31])
    got = G.VariableBaseMSM.multi_scalar_mul(bases, arr, first=5)   # bases.len() != scalars.len()
    assert projective_to_point(C, got) == O.msm_naive(C, pts[5:], sc)
    got = G.VariableBaseMSM.multi_scalar_mul(bases, arr[:0])
    assert projective_to_point(C, got) is None
    got = G.VariableBaseMSM.multi_scalar_mul(coords, arr, group=group, infinity=inf, ctx=ctx)
    assert projective_to_point(C, got) == O.msm_naive(C, pts, sc)
    bases.free()


@pytest.mark.parametrize("group", [ffi.MNT4_G1, ffi.MNT6_G1])
def test_msm_accumulate_exceptions(ctx, group):
    """the branches of the accumulation kernel's mixed addition (EcS::madd_acc_g): acc + (-acc) -> infinity
    and on from there, acc + acc -> doubling (also of a negated base), in one bucket each"""
    C = GROUPS[group]
    P, Q, R = sample_points(C, 3, 0xE0 + group)
    s, t = sample_scalars(C, 2, 0xE8 + group)
    for pts, sc in (([P, C.neg(P)], [s, s]),
                    ([P, C.neg(P), Q], [s, s, s]),
                    ([P, P], [s, s]),
                    ([P, P, C.neg(P), C.neg(P), R], [s, s, s, s, t]),
                    ([P, C.neg(P), C.neg(P)], [s, s, s]),
                    ([P, P], [C.r - s, C.r - s])):
        coords, inf = points_to_arrays(C, pts)
        got = G.VariableBaseMSM.multi_scalar_mul(coords, ints_to_array(sc), group=group, infinity=inf, ctx=ctx)
        assert projective_to_point(C, got) == O.msm_naive(C, pts, sc)


@pytest.mark.parametrize("form", ["1", "0"])
@pytest.mark.parametrize("group", sorted(GROUPS))
def test_msm_accumulation_forms(ctx, monkeypatch, group, form):
    """both accumulation forms, forced, on every group and on the device's lane-cooperative towers: the
    pairwise tree of affine additions with shared inversions (k_tree_round) and the XYZZ running sums
    (k_bucket_acc); doubling, cancellation, infinity operands, long runs, single entries"""
    import shared_checks
    monkeypatch.setenv("G753_MSM_AFFINE", form)
    monkeypatch.setenv("G753_MSM_C", "5")
    shared_checks.check_msm_accumulation_cases(ctx, group)
    if form == "1":      # long batches: one thread walks many buckets
        monkeypatch.setenv("G753_TREE_BATCH", "37")
        shared_checks.check_msm_accumulation_cases(ctx, group)


def test_msm_linearity_2e14(ctx):
    """size-independent property: msm(B, s) + msm(B, t) == msm(B, s + t mod r); bases are a
    short list repeated, so the oracle answer is also computable directly."""
    C = O.MNT4_G1
    base_pts = sample_points(C, 8, 0x51)
    n = 1 << 14
    reps = n // 8
    coords8, _ = points_to_arrays(C, base_pts)
    coords = np.tile(coords8, (reps, 1))
    rng = np.random.default_rng(7)
    s = [int.from_bytes(rng.bytes(94), "little") % C.r for _ in range(n)]
    t = [int.from_bytes(rng.bytes(94), "little") % C.r for _ in range(n)]
    bases = ctx.upload_bases(ffi.MNT4_G1, coords)
    r1 = projective_to_point(C, G.VariableBaseMSM.multi_scalar_mul(bases, ints_to_array(s)))
    r2 = projective_to_point(C, G.VariableBaseMSM.multi_scalar_mul(bases, ints_to_array(t)))
    r3 = projective_to_point(C, G.VariableBaseMSM.multi_scalar_mul(
        bases, ints_to_array([(x + y) % C.r for x, y in zip(s, t)])))
    assert C.add(r1, r2) == r3
    # direct: sum over the 8 distinct points of (sum of their scalars mod r)
    agg = [sum(s[j::8]) % C.r for j in range(8)]
    assert r1 == O.msm_naive(C, base_pts, agg)
    bases.free()


@pytest.mark.parametrize("field", sorted(FIELDS))
def test_ntt_vs_oracle(ctx, field):
    F = FIELDS[field]
    rng = O.SplitMix64(0xD0 + field)
    for log_n in (0, 1, 4, 9):
        n = 1 << log_n
        dom = G.EvaluationDomain.new(field, n, ctx=ctx)
        ref = O.EvaluationDomain(F, n)
        a = [O.random_field_element(rng, F) for _ in range(n)]
        arr = field_array(F, a)
        assert array_field(F, dom.fft(arr)) == ref.fft(a)
        assert array_field(F, dom.ifft(arr)) == ref.ifft(a)
        assert array_field(F, dom.coset_fft(arr)) == ref.coset_fft(a)
        assert array_field(F, dom.coset_ifft(arr)) == ref.coset_ifft(a)
    dom = G.EvaluationDomain.new(field, 100, ctx=ctx)      # ragged: pads to 128
    ref = O.EvaluationDomain(F, 100)
    a = [O.random_field_element(rng, F) for _ in range(100)]
    assert dom.size() == 128
    assert array_field(F, dom.fft(field_array(F, a))) == ref.fft(a)


@pytest.mark.parametrize("field", sorted(FIELDS))
def test_domain_public_fields_and_pointwise(ctx, field):
    """EvaluationDomain's public fields (domain.rs:24-39), mul_polynomials_in_evaluation_domain
    (:289-302) and divide_by_vanishing_poly_on_coset_in_place (:245-256) on the device"""
    F = FIELDS[field]
    rng = O.SplitMix64(0xD7 + field)
    for n in (1, 2, 1 << 10, 1 << (F.two_adicity - 1) if F.two_adicity < 20 else 1 << 22):
        dom = G.EvaluationDomain.new(field, n, ctx=ctx)
        ref = O.EvaluationDomain(F, n)
        assert array_field(F, dom.size_inv) == [ref.size_inv]
        assert array_field(F, dom.group_gen) == [ref.group_gen]
        assert array_field(F, dom.group_gen_inv) == [ref.group_gen_inv]
        assert array_field(F, dom.generator_inv) == [ref.generator_inv]
        assert array_field(F, dom.vanishing_on_coset_inv) == [pow(pow(ref.generator, n, F.p) - 1, -1, F.p)]
    n = 1 << 10
    dom = G.EvaluationDomain.new(field, n, ctx=ctx)
    a = [O.random_field_element(rng, F) for _ in range(n)]
    b = [O.random_field_element(rng, F) for _ in range(n)]
    got = dom.mul_polynomials_in_evaluation_domain(field_array(F, a), field_array(F, b))
    assert array_field(F, got) == [x * y % F.p for x, y in zip(a, b)]
    zinv = pow(pow(F.generator, n, F.p) - 1, -1, F.p)
    got = dom.divide_by_vanishing_poly_on_coset_in_place(field_array(F, a))
    assert array_field(F, got) == [x * zinv % F.p for x in a]


def test_ntt_max_size_mnt6_fr(ctx):
    """mnt6753::Fr tops out at 2^14: bit-exact round trip + spot evaluation"""
    F = O.MNT6_FR
    field = ffi.FIELD_MNT6_FR
    n = 1 << 14
    assert G.EvaluationDomain.new(field, n + 1, ctx=ctx) is None
    dom = G.EvaluationDomain.new(field, n, ctx=ctx)
    rng = np.random.default_rng(3)
    a = [int.from_bytes(rng.bytes(94), "little") % F.p for _ in range(n)]
    arr = field_array(F, a)
    ev = dom.fft(arr)
    assert np.array_equal(dom.ifft(ev), arr)
    ref = O.EvaluationDomain(F, n)
    evi = array_field(F, ev)
    for i in (0, 1, 5, n // 2, n - 1):
        x = pow(ref.group_gen, i, F.p)
        acc = 0
        for c in reversed(a):
            acc = (acc * x + c) % F.p
        assert evi[i] == acc
    cev = dom.coset_fft(arr)
    assert np.array_equal(dom.coset_ifft(cev), arr)


def test_ntt_large_properties(ctx):
    """2^20 on mnt4753::Fr (BASELINE config 2): round trips, linearity and Horner spot checks"""
    F = O.MNT4_FR
    field = ffi.FIELD_MNT4_FR
    n = 1 << 20
    dom = G.EvaluationDomain.new(field, n, ctx=ctx)
    rng = np.random.default_rng(11)
    raw = rng.integers(0, 1 << 63, size=(n, 12), dtype=np.uint64)
    raw[:, 11] &= np.uint64(0xFFFF)             # < 2^752 < p : valid Montgomery representations
    ev = dom.fft(raw)
    assert np.array_equal(dom.ifft(ev), raw)
    cev = dom.coset_fft(raw)
    assert np.array_equal(dom.coset_ifft(cev), raw)
    ref = O.EvaluationDomain(F, n)
    coeffs = array_field(F, raw)
    evi = array_field(F, ev[:4])
    cevi = array_field(F, cev[:4])
    for i in (0, 1, 3):
        for vals, shift in ((evi, 1), (cevi, ref.generator)):
            x = shift * pow(ref.group_gen, i, F.p) % F.p
            acc = 0
            for c in reversed(coeffs):
                acc = (acc * x + c) % F.p
            assert vals[i] == acc


@pytest.mark.parametrize("field,log_n", [(ffi.FIELD_MNT4_FR, 20), (ffi.FIELD_MNT6_FR, 14)])
def test_ntt_config2_vs_cpp_restatement(ctx, field, log_n):
    """BASELINE config 2 at its full sizes (2^20 on mnt4753::Fr, the 2^14 maximum of mnt6753::Fr): all four
    transforms bit-exact, every limb, against the C++ restatement of domain.rs:120-179 / 305-416"""
    from oracle import ref753
    n = 1 << log_n
    rng = np.random.default_rng(21 + log_n)
    raw = rng.integers(0, 1 << 63, size=(n, 12), dtype=np.uint64)
    raw[:, 11] &= np.uint64(0xFFFF)
    dom = G.EvaluationDomain.new(field, n, ctx=ctx)
    for mode, run in ((ffi.FFT, dom.fft), (ffi.IFFT, dom.ifft), (ffi.COSET_FFT, dom.coset_fft),
                      (ffi.COSET_IFFT, dom.coset_ifft)):
        assert np.array_equal(run(raw), ref753.fft(field, raw, mode)), mode


# ---- synthetic key generator + full-size properties -------------------------------------------
def _dot_mod(scalars, logs, r):
    return sum(int(s) * int(a) for s, a in zip(array_to_ints(scalars), logs)) % r


@pytest.mark.parametrize("group", sorted(GROUPS))
def test_generated_bases_vs_oracle(ctx, group):
    """g753_bases_generate: bases[i] = a_i * G, affine, for the documented a_i"""
    from util753 import g2_generator
    C = GROUPS[group]
    n = 5
    bases = ctx.generate_bases(group, n, 0xABC + group)
    logs = G.Bases.generated_logs(n, 0xABC + group)
    got = bases.download()
    k = C.F.k
    F = C.F.base
    params = __import__("importlib").import_module("ginger-lib_b200.params")
    gen_m = params.GENERATOR_MONT[group]
    gen = (tuple(F.from_mont(v) for v in gen_m[:k]), tuple(F.from_mont(v) for v in gen_m[k:]))
    if k > 1:
        assert gen == g2_generator(C)
    for i in range(n):
        vals = [F.from_mont(v) for v in array_to_ints(got[i].reshape(-1, 12))]
        assert (tuple(vals[:k]), tuple(vals[k:])) == C.mul(gen, int(logs[i]))
    # MSM over the generated key == (sum s_i a_i) * G
    sc = sample_scalars(C, n, 0x77)
    out = G.VariableBaseMSM.multi_scalar_mul(bases, ints_to_array(sc))
    assert projective_to_point(C, out) == C.mul(gen, _dot_mod(ints_to_array(sc), logs, C.r))
    bases.free()


@pytest.mark.parametrize("group,log_n", [(ffi.MNT4_G1, 20), (ffi.MNT6_G1, 16), (ffi.MNT4_G2, 15), (ffi.MNT6_G2, 14)])
def test_msm_large_discrete_log_property(ctx, group, log_n):
    """size-independent check at sizes no CPU oracle reaches: sum_i s_i (a_i G) == (sum_i s_i a_i mod r) G"""
    import bench
    C = GROUPS[group]
    n = 1 << log_n
    params = __import__("importlib").import_module("ginger-lib_b200.params")
    bases = ctx.generate_bases(group, n, 0x5EED + group)
    logs = G.Bases.generated_logs(n, 0x5EED + group)
    sc = bench.random_scalars(n, 0xF00 + group)
    sc[1] = 0                                     # zero scalar
    sc[2] = 0
    sc[2, 0] = 1                                  # scalar one
    sc[3] = ints_to_array([C.r - 1])[0]
    out = G.VariableBaseMSM.multi_scalar_mul(bases, sc)
    k = bench.dot_mod(sc, logs, C.r)
    F = C.F.base
    kk = C.F.k
    gen_m = params.GENERATOR_MONT[group]
    gen = (tuple(F.from_mont(v) for v in gen_m[:kk]), tuple(F.from_mont(v) for v in gen_m[kk:]))
    assert projective_to_point(C, out) == C.mul(gen, k)
    # zip truncation on a slice view of the resident key (groth16/mod.rs:318-350)
    out = G.VariableBaseMSM.multi_scalar_mul(bases, sc[:1000], first=n - 1500)
    k_slice = bench.dot_mod(sc[:1000], logs[n - 1500:n - 500], C.r)
    assert projective_to_point(C, out) == C.mul(gen, k_slice)
    # same key with the precomputed shifted copies (g753_bases_precompute): same group element,
    # for the whole key, a long slice (on the copies) and a short slice (plain pipeline)
    plain = out
    bases.precompute(8 if kk == 1 else 4)
    out = G.VariableBaseMSM.multi_scalar_mul(bases, sc)
    assert projective_to_point(C, out) == C.mul(gen, k)
    out = G.VariableBaseMSM.multi_scalar_mul(bases, sc[:1000], first=n - 1500)
    assert projective_to_point(C, out) == projective_to_point(C, plain)
    m = n // 2 + 3
    out = G.VariableBaseMSM.multi_scalar_mul(bases, sc[:m], first=5)
    assert projective_to_point(C, out) == C.mul(gen, bench.dot_mod(sc[:m], logs[5:5 + m], C.r))
    bases.free()


@pytest.mark.parametrize("group,log_n", [(ffi.MNT4_G1, 16), (ffi.MNT6_G1, 14), (ffi.MNT4_G2, 12), (ffi.MNT6_G2, 10)])
def test_msm_vs_cpp_restatement(ctx, group, log_n):
    """CUDA MSM == the C++ restatement of the reference's Pippenger (oracle/ref753.cpp) on the same
    bases and scalars, after normalisation (BASELINE config 1 shape, smaller n)"""
    import bench
    from oracle import ref753
    C = GROUPS[group]
    n = 1 << log_n
    bases = ctx.generate_bases(group, n, 0xC0FFEE + group)
    coords = bases.download()
    sc = bench.random_scalars(n, 0xBEEF + group)
    got = G.VariableBaseMSM.multi_scalar_mul(bases, sc)
    want = ref753.msm(group, coords, None, sc)
    assert projective_to_point(C, got) == projective_to_point(C, want)
    bases.precompute(0)
    got = G.VariableBaseMSM.multi_scalar_mul(bases, sc)
    assert projective_to_point(C, got) == projective_to_point(C, want)
    bases.free()


@pytest.mark.parametrize("field,n", [(ffi.FIELD_MNT6_FR, 40), (ffi.FIELD_MNT4_FR, 48 * 25)])
def test_mixed_radix_ntt_vs_definition(ctx, field, n):
    """BASELINE config 4 (mixed-radix FFT): the four transforms against the direct DFT definition;
    parity unpinned - the reference has no mixed-radix domain"""
    F = FIELDS[field]
    p = F.p
    dom = G.MixedRadixDomain.new(field, n, ctx=ctx)
    assert dom is not None
    w = O.mixed_radix_omega(F, n)
    rng = O.SplitMix64(0x3B + n)
    a = [O.random_field_element(rng, F) for _ in range(n)]
    ninv = pow(n, -1, p)
    arr = field_array(F, a)
    assert array_field(F, dom.fft(arr)) == O.dft_naive(a, w, p)
    inv = [x * ninv % p for x in O.dft_naive(a, pow(w, -1, p), p)]
    assert array_field(F, dom.ifft(arr)) == inv
    assert array_field(F, dom.coset_fft(arr)) == O.dft_naive(O.distribute_powers(a, F.generator, p), w, p)
    assert array_field(F, dom.coset_ifft(arr)) == O.distribute_powers(inv, pow(F.generator, -1, p), p)


@pytest.mark.parametrize("field,n", [(ffi.FIELD_MNT6_FR, (1 << 15) * 25), (ffi.FIELD_MNT4_FR, (1 << 18) * 5)])
def test_mixed_radix_ntt_full_size_properties(ctx, field, n):
    """SURVEY.md 8c candidate sizes (819 200 on mnt6753::Fr, 1 310 720 on mnt4753::Fr): round trips and
    evaluation at the points omega^k of a polynomial given by a few coefficients"""
    import bench
    F = FIELDS[field]
    p = F.p
    dom = G.MixedRadixDomain.new(field, n, ctx=ctx)
    raw = bench.random_scalars(n, 0x44)
    raw[:, 11] &= np.uint64(0xFFFF)
    assert (dom.ifft(dom.fft(raw)) == raw).all()
    assert (dom.coset_ifft(dom.coset_fft(raw)) == raw).all()
    coeffs = [3, 1, 4, 1, 5, 9, 2, 6]
    ev = array_field(F, dom.fft(field_array(F, coeffs)))
    w = O.mixed_radix_omega(F, n)
    for k in (0, 1, 2, 12345, n - 1):
        x = pow(w, k, p)
        assert ev[k] == sum(c * pow(x, i, p) for i, c in enumerate(coeffs)) % p


@pytest.mark.parametrize("group", sorted(GROUPS))
def test_bases_from_wire_format(ctx, group):
    """SURVEY.md 8f-2, proving-key loader: GroupAffine::write records (x || y || infinity byte, canonical
    little-endian) are converted to the resident Montgomery layout on the device"""
    from util753 import sample_points
    C = GROUPS[group]
    k = C.F.k
    pts = sample_points(C, 9, 0x3C0 + group)
    pts[4] = None
    wire = b""
    for P in pts:
        if P is None:
            wire += O.int_to_bytes96(0) * k + O.int_to_bytes96(1) + O.int_to_bytes96(0) * (k - 1) + b"\x01"
        else:
            wire += b"".join(O.int_to_bytes96(c) for c in P[0]) + b"".join(O.int_to_bytes96(c) for c in P[1]) + b"\x00"
    bases = G.Bases.from_wire(ctx, group, wire)
    coords, inf = points_to_arrays(C, pts)
    ref = ctx.upload_bases(group, coords, inf)
    assert (bases.download() == ref.download()).all()
    sc = sample_scalars(C, 9, 0x3D0 + group)
    got = G.VariableBaseMSM.multi_scalar_mul(bases, ints_to_array(sc))
    assert projective_to_point(C, got) == O.msm_naive(C, pts, sc)
    bases.free()
    ref.free()


@pytest.mark.parametrize("group", sorted(GROUPS))
def test_fixed_base_msm(ctx, group):
    """FixedBaseMSM::multi_scalar_mul + batch normalisation (fixed_base.rs:66-79,
    generator.rs:243-284) - scalars[i] * G, affine, against the oracle, then 2^12 scalars against the
    variable-base MSM (sum_i t_i (s_i G) == (sum_i t_i s_i) G)"""
    import bench
    from util753 import sample_points
    C = GROUPS[group]
    base = sample_points(C, 1, 0x4F0 + group)[0]
    sc = [0, 1, C.r - 1, 255, 256, (1 << 752) + 12345] + sample_scalars(C, 3, 0x4F1)
    coords, _ = points_to_arrays(C, [base])
    out, inf = G.FixedBaseMSM.multi_scalar_mul(group, coords[0], ints_to_array(sc), ctx=ctx)
    want_c, want_inf = points_to_arrays(C, [C.mul(base, s) for s in sc])
    assert (inf == want_inf).all() and (out == want_c).all()
    n = 1 << 12
    s_arr = bench.random_scalars(n, 0x4F2 + group)
    pts, inf = G.FixedBaseMSM.multi_scalar_mul(group, coords[0], s_arr, ctx=ctx)
    assert not inf.any()
    t_arr = bench.random_scalars(n, 0x4F3 + group)
    bases = ctx.upload_bases(group, pts)
    got = G.VariableBaseMSM.multi_scalar_mul(bases, t_arr)
    k = sum(int(a) * int(b) for a, b in zip(array_to_ints(s_arr), array_to_ints(t_arr))) % C.r
    assert projective_to_point(C, got) == C.mul(base, k)
    bases.free()


@pytest.mark.parametrize("group", sorted(GROUPS))
def test_bases_from_wire_rejects_malformed(ctx, group):
    import shared_checks
    shared_checks.check_wire_rejects_malformed(ctx, group)


@pytest.mark.parametrize("group", [ffi.MNT4_G2, ffi.MNT6_G2])
@pytest.mark.parametrize("lanes", [0, 1])
def test_reference_ext_kats_on_device_towers(ctx, group, lanes):
    """the reference's Fq2 / Fq3 KATs (fields/mnt4753/tests.rs:1071-1867, fields/mnt6753/tests.rs:1277-2378)
    and random elements through the LANE-COOPERATIVE towers Tw2C / Tw3C (slots.cuh) the G2 kernels run, in the
    lane split of the accumulation kernels (lanes = 0: 2 / 4 lanes per element) and of the reduction
    kernels (lanes = 1: 4 / 8 lanes)"""
    import shared_checks
    shared_checks.check_reference_ext_kats(ctx, group, lanes)
    shared_checks.check_ext_ops_random(ctx, group, lanes)


@pytest.mark.parametrize("fid", [0, 1])
def test_coop_field_arithmetic_vs_model(ctx, fid):
    """the warp-cooperative field arithmetic (csrc/coop.cuh: one product on 8 lanes, digit-serial Montgomery in
    base 2^96, ballot-resolved carries) limb for limb against its Python model (tools/gen_coop.py), on edge
    values (all-ones digits, values at the bounds, carries rippling across every lane) and random ones"""
    import importlib.util
    import random
    spec = importlib.util.spec_from_file_location("gen_coop", os.path.join(HERE, "..", "tools", "gen_coop.py"))
    GC = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(GC)
    p = GC.field_moduli()[fid]
    f = GC.Field(p)
    rng = random.Random(0xC0F + fid)
    M96 = (1 << 96) - 1
    edge = [0, 1, p - 1, p, p + 1, 2 * p - 1, M96, M96 << 96, (1 << 753) - 1, 100 * p, (1 << 192) - 1,
            sum(M96 << (192 * i) for i in range(4)), (1 << 767) - 1, sum(M96 << (96 * i) for i in range(7)),
            (1 << 96), (1 << 672)]
    vals = edge + [rng.randrange(0, 181 * p) for _ in range(80)]

    def arr(xs):
        return ints_to_array(xs)

    def run(op, xs, ys, k=0):
        a, b = arr(xs), arr(ys)
        out = np.zeros_like(a)
        ctx.lib.check(ctx.lib.coop_op(ctx.handle, fid, op, k, ffi.ptr(a), ffi.ptr(b), ffi.ptr(out), len(xs)))
        return array_to_ints(out)

    # products (input bounds x y <= 2^15)
    xs, ys = [], []
    for i, x in enumerate(vals):
        for y in vals[i % 5::5]:
            if (x // p + 1) * (y // p + 1) <= (1 << 15) and x < (1 << 768) and y < (1 << 768):
                xs.append(x)
                ys.append(y)
    got = run(0, xs, ys)
    want = [GC.join(f.mul(GC.split(x), GC.split(y))) for x, y in zip(xs, ys)]
    assert got == want
    # sums and differences
    xs, ys = [], []
    for x in vals:
        for y in vals[::4]:
            if x + y < (1 << 768):
                xs.append(x)
                ys.append(y)
    assert run(1, xs, ys) == [x + y for x, y in zip(xs, ys)]
    for k in range(GC.MAX_K_LOG + 1):
        xs, ys = [], []
        for x in vals:
            for y in vals[::6]:
                if y <= (p << k) and x + (p << k) - y < (1 << 768):
                    xs.append(x)
                    ys.append(y)
        assert run(2, xs, ys, k) == [x + (p << k) - y for x, y in zip(xs, ys)]
    # canonicalisation and the zero test of values below 2 p
    xs = [0, 1, p - 1, p, p + 1, 2 * p - 1] + [rng.randrange(0, 2 * p) for _ in range(30)]
    assert run(3, xs, xs) == [x % p for x in xs]
    assert run(4, xs, xs) == [1 if x % p == 0 else 0 for x in xs]
