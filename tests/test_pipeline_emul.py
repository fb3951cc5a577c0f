"""CPU-tier check of the kernel/orchestration logic through the C ABI, using the TEST-ONLY
host-emulation build (tests/host_emul/capi_emul.cpp): digit recoding, counting sort, bucket
bookkeeping, reduction levels, window fold and Stockham indexing versus the oracle.

The same assertions run against the real library on a B200 in tests/test_gpu_parity.py.
"""
import os
import subprocess

import numpy as np
import pytest

from oracle import g753 as O
from util753 import (FIELDS, G, GROUPS, array_field, field_array, ffi, ints_to_array, points_to_arrays,
                     projective_to_point, sample_points, sample_scalars)

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "capi_emul.cpp")
LIB = os.path.join(HERE, "host_emul", "libg753_emul.so")
CSRC = os.path.join(HERE, "..", "ginger-lib_b200", "csrc")


@pytest.fixture(scope="module")
def ctx():
    from util753 import build_emul
    lib = ffi.Library(build_emul())
    c = G.Context(0, library=lib)
    yield c
    c.close()


@pytest.mark.parametrize("group", sorted(GROUPS))
def test_msm_small(ctx, group):
    C = GROUPS[group]
    n = 24 if C.F.k == 1 else 10
    pts = sample_points(C, n, 0xA0 + group)
    sc = sample_scalars(C, n, 0xB0 + group)
    # the special cases the reference branches on (variable_base.rs:36-57) plus signed-digit edges
    pts[1] = None
    sc[2] = 0
    sc[3] = 1
    sc[4] = C.r - 1
    pts[6] = pts[5]
    sc[6] = sc[5]
    pts[7] = C.neg(pts[5])
    sc[7] = sc[5]
    sc[8] = (1 << 752) - 1 if (1 << 752) - 1 < C.r else C.r - 2
    coords, inf = points_to_arrays(C, pts)
    bases = ctx.upload_bases(group, coords, inf)
    got = G.VariableBaseMSM.multi_scalar_mul(bases, ints_to_array(sc))
    assert projective_to_point(C, got) == O.msm_naive(C, pts, sc)
    # zip-truncation both ways + slice views (groth16/mod.rs:318-350)
    got = G.VariableBaseMSM.multi_scalar_mul(bases, ints_to_array(sc[:5]))
    assert projective_to_point(C, got) == O.msm_naive(C, pts[:5], sc[:5])
    got = G.VariableBaseMSM.multi_scalar_mul(bases, ints_to_array(sc), first=3)
    assert projective_to_point(C, got) == O.msm_naive(C, pts[3:], sc)
    # empty -> zero
    got = G.VariableBaseMSM.multi_scalar_mul(bases, ints_to_array([]))
    assert projective_to_point(C, got) is None
    # one-shot host form
    got = G.VariableBaseMSM.multi_scalar_mul(coords, ints_to_array(sc + [5]), group=group, infinity=inf, ctx=ctx)
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


@pytest.fixture(scope="module")
def ctx_optin():
    """the opt-in kernel forms (six-slot mixed addition, rolled multiplier) in their own emulation build"""
    from util753 import build_emul
    lib = ffi.Library(build_emul(("-DG753_ACC6=1", "-DG753_ROLLED=1"), "_acc6_rolled"))
    c = G.Context(0, library=lib)
    yield c
    c.close()


@pytest.mark.parametrize("group", [ffi.MNT4_G1])     # (CPU suite time; the forms only differ on the prime-field curves)
def test_msm_optin_forms(ctx_optin, group):
    """-DG753_ACC6=1 -DG753_ROLLED=1: same results from the six-slot addition and the rolled multiplier"""
    test_msm_small(ctx_optin, group)
    if GROUPS[group].F.k == 1:
        test_msm_accumulate_exceptions(ctx_optin, group)


def test_msm_window_sizes(ctx, monkeypatch):
    """force several window widths through the same input (exercises multi-level reduction)"""
    C = O.MNT4_G1
    n = 40
    pts = sample_points(C, n, 0x77)
    sc = sample_scalars(C, n, 0x78)
    coords, inf = points_to_arrays(C, pts)
    want = O.msm_naive(C, pts, sc)
    for c in (3, 7, 11):
        monkeypatch.setenv("G753_MSM_C", str(c))
        cx = G.Context(0, library=ctx.lib)
        bases = cx.upload_bases(ffi.MNT4_G1, coords, inf)
        got = G.VariableBaseMSM.multi_scalar_mul(bases, ints_to_array(sc))
        assert projective_to_point(C, got) == want, c
        # g753_last_msm_plan: W * c covers the 753 scalar bits + the sign carry; a plain key has one row per window
        plan = cx.last_msm_plan()
        assert plan == {"c": c, "windows": -(-754 // c), "rows": -(-754 // c), "copies": 1,
                        "accumulation": "xyzz_running_sums"}      # short MSMs keep the running sums
        bases.free()
        cx.close()


@pytest.mark.parametrize("group", [ffi.MNT4_G1, ffi.MNT4_G2])
def test_msm_heavy_buckets(ctx, group):
    """skewed scalars (many equal / tiny ones, as boolean witnesses give): buckets longer than
    one work item are cut into several items and re-joined by k_bucket_fixup"""
    C = GROUPS[group]
    base = sample_points(C, 3, 0x91 + group)
    n = 150
    pts = [base[i % 3] for i in range(n)]
    sc = [5] * 140 + [1, 0, 2, 7, C.r - 5, 5, 5, 3, 1, 1]
    coords, inf = points_to_arrays(C, pts)
    bases = ctx.upload_bases(group, coords, inf)
    got = G.VariableBaseMSM.multi_scalar_mul(bases, ints_to_array(sc))
    assert projective_to_point(C, got) == O.msm_naive(C, pts, sc)
    bases.free()


def test_msm_very_heavy_bucket_tree_fixup(ctx):
    """one bucket holding 1300 points (11 work items of <= 128): the partial sums are joined by the
    in-place tree of k_bucket_fixup_level over two levels"""
    C = O.MNT4_G1
    base = sample_points(C, 5, 0x99)
    n = 1300
    pts = [base[i % 5] for i in range(n)]
    sc = [3] * n
    sc[10] = C.r - 3
    sc[11] = 0
    coords, inf = points_to_arrays(C, pts)
    bases = ctx.upload_bases(ffi.MNT4_G1, coords, inf)
    got = G.VariableBaseMSM.multi_scalar_mul(bases, ints_to_array(sc))
    cnt = [sum(1 for i in range(n) if i % 5 == j and i not in (10, 11)) for j in range(5)]
    want = None
    for j in range(5):
        want = C.add(want, C.mul(base[j], 3 * cnt[j] % C.r))
    want = C.add(want, C.mul(base[0], C.r - 3))
    assert projective_to_point(C, got) == want
    bases.free()


@pytest.mark.parametrize("field", sorted(FIELDS))
def test_ntt_small(ctx, field):
    F = FIELDS[field]
    rng = O.SplitMix64(0xD0 + field)
    for log_n in (0, 1, 2, 5, 7):
        n = 1 << log_n
        dom = G.EvaluationDomain.new(field, n, ctx=ctx)
        ref = O.EvaluationDomain(F, n)
        a = [O.random_field_element(rng, F) for _ in range(n)]
        arr = field_array(F, a)
        assert array_field(F, dom.fft(arr)) == ref.fft(a)
        assert array_field(F, dom.ifft(arr)) == ref.ifft(a)
        assert array_field(F, dom.coset_fft(arr)) == ref.coset_fft(a)
        assert array_field(F, dom.coset_ifft(arr)) == ref.coset_ifft(a)
    # short input zero-pads, long input truncates (domain.rs:121)
    dom = G.EvaluationDomain.new(field, 6, ctx=ctx)
    assert dom.size() == 8
    ref = O.EvaluationDomain(F, 6)
    a = [O.random_field_element(rng, F) for _ in range(11)]
    assert array_field(F, dom.coset_fft(field_array(F, a[:5]))) == ref.coset_fft(a[:5])
    assert array_field(F, dom.fft(field_array(F, a))) == ref.fft(a)


def test_domain_new_none(ctx):
    assert G.EvaluationDomain.new(ffi.FIELD_MNT6_FR, (1 << 14) + 1, ctx=ctx) is None
    assert G.EvaluationDomain.new(ffi.FIELD_MNT6_FR, 1 << 14, ctx=ctx) is not None
    assert G.EvaluationDomain.new(ffi.FIELD_MNT4_FR, (1 << 29) + 1, ctx=ctx) is None


@pytest.mark.parametrize("field", sorted(FIELDS))
def test_domain_public_fields_and_pointwise(ctx, field):
    """EvaluationDomain's public fields (domain.rs:24-39), mul_polynomials_in_evaluation_domain
    (:289-302) and divide_by_vanishing_poly_on_coset_in_place (:245-256)"""
    F = FIELDS[field]
    rng = O.SplitMix64(0xD7 + field)
    for n in (1, 2, 8, 1 << 13):
        dom = G.EvaluationDomain.new(field, n, ctx=ctx)
        ref = O.EvaluationDomain(F, n)
        assert array_field(F, dom.size_inv) == [ref.size_inv]
        assert array_field(F, dom.group_gen) == [ref.group_gen]
        assert array_field(F, dom.group_gen_inv) == [ref.group_gen_inv]
        assert array_field(F, dom.generator_inv) == [ref.generator_inv]
        zinv = pow(pow(ref.generator, n, F.p) - 1, -1, F.p)
        assert array_field(F, dom.vanishing_on_coset_inv) == [zinv]
    with pytest.raises(AttributeError):
        dom.no_such_field
    n = 8
    dom = G.EvaluationDomain.new(field, n, ctx=ctx)
    a = [O.random_field_element(rng, F) for _ in range(n)]
    b = [O.random_field_element(rng, F) for _ in range(n)]
    got = dom.mul_polynomials_in_evaluation_domain(field_array(F, a), field_array(F, b))
    assert array_field(F, got) == [x * y % F.p for x, y in zip(a, b)]
    zinv = pow(pow(F.generator, n, F.p) - 1, -1, F.p)
    got = dom.divide_by_vanishing_poly_on_coset_in_place(field_array(F, a))
    assert array_field(F, got) == [x * zinv % F.p for x in a]


@pytest.mark.parametrize("field", sorted(FIELDS))
def test_dense_polynomial_mul(ctx, field):
    """`&DensePolynomial * &DensePolynomial` through the transforms (dense.rs:342-357) = schoolbook product"""
    F = FIELDS[field]
    rng = O.SplitMix64(0xDE + field)
    for la, lb in ((1, 1), (3, 5), (8, 8), (17, 2)):
        a = [O.random_field_element(rng, F) for _ in range(la)]
        b = [O.random_field_element(rng, F) for _ in range(lb)]
        want = [0] * (la + lb - 1)
        for i, x in enumerate(a):
            for j, y in enumerate(b):
                want[i + j] = (want[i + j] + x * y) % F.p
        pa = G.DensePolynomial(field, field_array(F, a + [0, 0]), ctx)     # trailing zeros are dropped
        pb = G.DensePolynomial(field, field_array(F, b), ctx)
        assert pa.degree() == la - 1
        assert array_field(F, (pa * pb).coeffs) == want
    zero = G.DensePolynomial(field, field_array(F, [0, 0]), ctx)
    assert zero.is_zero() and (zero * pb).is_zero() and (pb * zero).is_zero()


@pytest.mark.parametrize("field", sorted(FIELDS))
def test_domain_elements_vanishing_lagrange(ctx, field):
    """elements, evaluate_vanishing_polynomial, evaluate_all_lagrange_coefficients, reindex_by_subdomain
    (domain.rs:183-284) against their definitions"""
    check_domain_methods(ctx, field, (1, 2, 8, 32))


def check_domain_methods(cx, field, sizes):
    F = FIELDS[field]
    p = F.p
    rng = O.SplitMix64(0xE1 + field)
    for n in sizes:
        dom = G.EvaluationDomain.new(field, n, ctx=cx)
        ref = O.EvaluationDomain(F, n)
        el = [pow(ref.group_gen, i, p) for i in range(n)]
        assert array_field(F, dom.elements()) == el
        tau = O.random_field_element(rng, F)
        assert array_field(F, dom.evaluate_vanishing_polynomial(field_array(F, [tau]))) == [(pow(tau, n, p) - 1) % p]
        (d0, c0), (d1, c1) = dom.vanishing_polynomial()
        assert (d0, d1) == (0, n) and array_field(F, c0) == [p - 1] and array_field(F, c1) == [1]
        # tau outside the domain: the Lagrange basis evaluated at tau (sums to 1, interpolates)
        z = (pow(tau, n, p) - 1) * ref.size_inv % p
        want = [z * w * pow((tau - w) % p, -1, p) % p for w in el]
        got = array_field(F, dom.evaluate_all_lagrange_coefficients(field_array(F, [tau])))
        assert got == want and sum(got) % p == 1
        # tau inside the domain: an indicator vector
        k = n // 2
        got = array_field(F, dom.evaluate_all_lagrange_coefficients(field_array(F, [el[k]])))
        assert got == [1 if i == k else 0 for i in range(n)]
    big, small = G.EvaluationDomain.new(field, 16, ctx=cx), G.EvaluationDomain.new(field, 4, ctx=cx)
    idx = [big.reindex_by_subdomain(small, i) for i in range(16)]
    assert idx[:4] == [0, 4, 8, 12] and sorted(idx) == list(range(16))
    assert idx[4:] == [1, 2, 3, 5, 6, 7, 9, 10, 11, 13, 14, 15]


def test_device_vector_chain(ctx):
    """ifft -> coset_fft -> pointwise -> coset_ifft chained on 'device' memory"""
    F = O.MNT4_FR
    field = ffi.FIELD_MNT4_FR
    rng = O.SplitMix64(99)
    n = 16
    a = [O.random_field_element(rng, F) for _ in range(n)]
    b = [O.random_field_element(rng, F) for _ in range(n)]
    ref = O.EvaluationDomain(F, n)
    va = G.DeviceVector(ctx, field, n, field_array(F, a))
    vb = G.DeviceVector(ctx, field, n, field_array(F, b))
    for v in (va, vb):
        v.ntt(ffi.IFFT)
        v.ntt(ffi.COSET_FFT)
    va.op(ffi.OP_MUL, vb)
    k = 12345
    va.scale(ints_to_array([F.to_mont(k)]))
    va.ntt(ffi.COSET_IFFT)
    ea = ref.coset_fft(ref.ifft(a))
    eb = ref.coset_fft(ref.ifft(b))
    want = ref.coset_ifft([x * y * k % F.p for x, y in zip(ea, eb)])
    assert array_field(F, va.download()) == want
    va.free()
    vb.free()


def test_bad_arguments(ctx):
    lib = ctx.lib
    assert lib.ntt(ctx.handle, 5, None, 3, 0) != 0
    assert lib.domain_check(0, 15) == ffi.ERR_DOMAIN
    assert lib.domain_check(1, 29) == ffi.OK
    assert lib.domain_check(1, 30) == ffi.ERR_DOMAIN
    assert lib.group_coord_limbs(ffi.MNT6_G2) == 36
    import ctypes
    h = ctypes.c_void_p()
    z = np.zeros((1, 24), dtype=np.uint64)
    assert lib.bases_upload(ctx.handle, 9, ffi.ptr(z), None, 1, ctypes.byref(h)) == ffi.ERR_BAD_ARG
    assert b"unknown group" in lib.last_error()
    out = np.zeros(12, dtype=np.uint64)
    assert lib.domain_constant(ctx.handle, ffi.FIELD_MNT6_FR, 15, 0, ffi.ptr(out)) == ffi.ERR_DOMAIN   # `None`
    assert lib.domain_constant(ctx.handle, ffi.FIELD_MNT6_FR, 3, 5, ffi.ptr(out)) == ffi.ERR_BAD_ARG
    assert lib.domain_constant(ctx.handle, ffi.FIELD_MNT6_FR, 3, 0, None) == ffi.ERR_BAD_ARG
    # a 2^29 domain's constants need no O(n) tables
    assert lib.domain_constant(ctx.handle, ffi.FIELD_MNT4_FR, 29, 1, ffi.ptr(out)) == ffi.OK
    F = O.MNT4_FR
    assert array_field(F, out) == [O.EvaluationDomain(F, 1 << 29).group_gen]


@pytest.mark.parametrize("group,c,copies,form", [(ffi.MNT4_G1, 7, 4, None), (ffi.MNT4_G1, 0, 8, None), (ffi.MNT6_G2, 9, 3, None),
                                                 (ffi.MNT4_G1, 5, 4, "1"), (ffi.MNT4_G2, 6, 3, "1")])
def test_msm_precomputed_key_copies(ctx, monkeypatch, group, c, copies, form):
    """g753_bases_precompute: copy j holds 2^(j*rows*c) * P_i; the MSM over the key (and over
    slices of it) must give the same group element as the plain pipeline / the naive sum.  form "1": the
    addition tree forced on the copies (several windows per bucket row, entries that point into different copies)"""
    C = GROUPS[group]
    if form is not None:
        monkeypatch.setenv("G753_MSM_AFFINE", form)
    n = 24 if C.F.k == 1 else 8
    pts = sample_points(C, n, 0x1A0 + group)
    sc = sample_scalars(C, n, 0x1B0 + group)
    pts[1] = None
    sc[2] = 0
    sc[3] = 1
    sc[4] = C.r - 1
    pts[6] = pts[5]
    pts[7] = C.neg(pts[5])
    sc[7] = sc[5]
    coords, inf = points_to_arrays(C, pts)
    if c:
        monkeypatch.setenv("G753_MSM_C", str(c))
    cx = G.Context(0, library=ctx.lib)
    bases = cx.upload_bases(group, coords, inf).precompute(copies)
    got = G.VariableBaseMSM.multi_scalar_mul(bases, ints_to_array(sc))
    assert projective_to_point(C, got) == O.msm_naive(C, pts, sc)
    got = G.VariableBaseMSM.multi_scalar_mul(bases, ints_to_array(sc), first=2)      # slice, still on the copies
    assert projective_to_point(C, got) == O.msm_naive(C, pts[2:], sc)
    got = G.VariableBaseMSM.multi_scalar_mul(bases, ints_to_array(sc[:3]), first=4)  # short slice: plain pipeline
    assert projective_to_point(C, got) == O.msm_naive(C, pts[4:7], sc[:3])
    keep = inf == 0                                        # infinite bases are zeroed on upload
    assert (bases.download(0, n)[keep] == coords[keep]).all()
    bases.free()
    cx.close()


@pytest.mark.parametrize("field,log_n", [(ffi.FIELD_MNT4_FR, 4), (ffi.FIELD_MNT6_FR, 5), (ffi.FIELD_MNT4_FR, 1), (ffi.FIELD_MNT4_FR, 0)])
def test_sharded_ntt_single_rank(ctx, field, log_n):
    """four-step decomposition (ntt_dist.cuh) with world = 1: column transforms + twiddles, the
    (identity) exchange, row transforms == the oracle's transform, all four modes"""
    D = __import__("importlib").import_module("ginger-lib_b200.distributed")
    F = FIELDS[field]
    n = 1 << log_n
    rng = O.SplitMix64(0x4D + log_n)
    a = [O.random_field_element(rng, F) for _ in range(n)]
    ref = O.EvaluationDomain(F, n)
    dom = D.ShardedEvaluationDomain(ctx, field, log_n)
    for mode, want in ((ffi.FFT, ref.fft(a)), (ffi.IFFT, ref.ifft(a)), (ffi.COSET_FFT, ref.coset_fft(a)),
                       (ffi.COSET_IFFT, ref.coset_ifft(a))):
        out = dom.gather(dom.transform(dom.scatter(field_array(F, a)), mode))
        assert array_field(F, out) == want, mode
    dom.close()


@pytest.mark.parametrize("field,n", [(ffi.FIELD_MNT4_FR, 12), (ffi.FIELD_MNT4_FR, 45), (ffi.FIELD_MNT6_FR, 40),
                                     (ffi.FIELD_MNT6_FR, 7), (ffi.FIELD_MNT4_FR, 16)])
def test_mixed_radix_ntt_vs_definition(ctx, field, n):
    """mixed-radix domain (n = 2^a m): the four transforms against the direct DFT definition
    (parity unpinned: the reference has no mixed-radix domain)"""
    F = FIELDS[field]
    p = F.p
    dom = G.MixedRadixDomain.new(field, n, ctx=ctx)
    assert dom is not None and dom.size() == n
    w = O.mixed_radix_omega(F, n)
    rng = O.SplitMix64(0x3A + n)
    a = [O.random_field_element(rng, F) for _ in range(n)]
    g, ninv = F.generator, pow(n, -1, p)
    arr = field_array(F, a)
    assert array_field(F, dom.fft(arr)) == O.dft_naive(a, w, p)
    assert array_field(F, dom.ifft(arr)) == [x * ninv % p for x in O.dft_naive(a, pow(w, -1, p), p)]
    assert array_field(F, dom.coset_fft(arr)) == O.dft_naive(O.distribute_powers(a, g, p), w, p)
    inv = [x * ninv % p for x in O.dft_naive(a, pow(w, -1, p), p)]
    assert array_field(F, dom.coset_ifft(arr)) == O.distribute_powers(inv, pow(g, -1, p), p)


def test_mixed_radix_domain_none(ctx):
    assert G.MixedRadixDomain.new(ffi.FIELD_MNT6_FR, 13 * 4, ctx=ctx) is None          # 13 does not divide p - 1
    assert G.MixedRadixDomain.new(ffi.FIELD_MNT6_FR, (1 << 16) * 3, ctx=ctx) is None   # beyond the two-adicity
    assert G.MixedRadixDomain.new(ffi.FIELD_MNT6_FR, (1 << 15) * 25, ctx=ctx) is not None
    assert G.MixedRadixDomain.new(ffi.FIELD_MNT4_FR, (1 << 18) * 5, ctx=ctx) is not None


@pytest.mark.parametrize("group,form", [(g, "1") for g in sorted(GROUPS)] + [(ffi.MNT4_G1, "0"), (ffi.MNT4_G2, "0")])
def test_msm_accumulation_forms(ctx, monkeypatch, group, form):
    """both accumulation forms, forced: the pairwise tree of affine additions with shared inversions
    (k_tree_round) on every group and the XYZZ running sums (k_bucket_acc; the form every short MSM of this
    suite runs anyway) on one curve of each kind; the GPU tier runs all eight combinations"""
    import shared_checks
    monkeypatch.setenv("G753_MSM_AFFINE", form)
    monkeypatch.setenv("G753_MSM_C", "5")
    shared_checks.check_msm_accumulation_cases(ctx, group)
    if form == "1" and group in (ffi.MNT4_G1, ffi.MNT4_G2):      # long batches: one thread walks many buckets
        monkeypatch.setenv("G753_TREE_BATCH", "37")
        shared_checks.check_msm_accumulation_cases(ctx, group)


@pytest.mark.parametrize("group", sorted(GROUPS))
def test_bases_from_wire_format(ctx, group):
    """proving-key loader: the reference's serialisation of affine points (x || y || infinity byte,
    canonical little-endian 96-byte elements; byte order pinned by the reference's *_tobyte fixtures
    through oracle.int_to_bytes96) lands in the same resident layout as the limb upload"""
    C = GROUPS[group]
    k = C.F.k
    pts = sample_points(C, 6, 0x3C0 + group)
    pts[2] = None
    wire = b""
    for P in pts:
        if P is None:
            wire += O.int_to_bytes96(0) * k + O.int_to_bytes96(1) + O.int_to_bytes96(0) * (k - 1) + b"\x01"   # zero() = (0, 1, true)
        else:
            wire += b"".join(O.int_to_bytes96(c) for c in P[0]) + b"".join(O.int_to_bytes96(c) for c in P[1]) + b"\x00"
    bases = G.Bases.from_wire(ctx, group, wire)
    coords, inf = points_to_arrays(C, pts)
    ref = ctx.upload_bases(group, coords, inf)
    assert len(bases) == 6
    assert (bases.download() == ref.download()).all()
    sc = sample_scalars(C, 6, 0x3D0 + group)
    got = G.VariableBaseMSM.multi_scalar_mul(bases, ints_to_array(sc))
    assert projective_to_point(C, got) == O.msm_naive(C, pts, sc)
    bases.free()
    ref.free()


@pytest.mark.parametrize("group", [ffi.MNT4_G1, ffi.MNT6_G2])
def test_fixed_base_msm(ctx, group):
    """FixedBaseMSM::multi_scalar_mul + normalisation (fixed_base.rs:66-79, generator.rs:243-284):
    scalars[i] * G for 0, 1, r - 1, a full-width and some small scalars"""
    C = GROUPS[group]
    base = sample_points(C, 1, 0x4F0 + group)[0]
    sc = [0, 1, C.r - 1, 255, 256, (1 << 752) + 12345] + sample_scalars(C, 2, 0x4F1)
    coords, _ = points_to_arrays(C, [base])
    out, inf = G.FixedBaseMSM.multi_scalar_mul(group, coords[0], ints_to_array(sc), ctx=ctx)
    want_c, want_inf = points_to_arrays(C, [C.mul(base, s) for s in sc])
    assert (inf == want_inf).all()
    assert (out == want_c).all()


@pytest.mark.parametrize("group", sorted(GROUPS))
def test_bases_from_wire_rejects_malformed(ctx, group):
    import shared_checks
    shared_checks.check_wire_rejects_malformed(ctx, group)


@pytest.mark.parametrize("group", [ffi.MNT4_G2, ffi.MNT6_G2])
def test_ext_op_entry_point(ctx, group):
    """g753_ext_op: the reference's Fq2 / Fq3 KATs + random elements through the C ABI (this tier runs the
    one-thread towers; the GPU tier runs the lane-cooperative ones through the same entry point)"""
    import shared_checks
    shared_checks.check_reference_ext_kats(ctx, group, 0)
    shared_checks.check_ext_ops_random(ctx, group, 1, count=12)
