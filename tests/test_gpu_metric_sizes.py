"""Parity at the sizes BASELINE.json's metric is quoted on, through the C ABI on a real B200:
config 3 (2^22-point MNT4-753 G1 MSM), the 2^22 FFT (all four transforms, every limb), config 4
(2^20-point MNT6-753 G2 / Fq3 MSM), config 5 (2^20 Groth16 proof), the log_n sweep of the transforms
(fft/test.rs:9-43 walks every size) and MSMs over the reference benchmark circuit's own witness.
The checker is the C++ restatement of the reference (oracle/ref753.cpp) wherever a CPU finishes in
seconds, a size-independent discrete-log property at full size otherwise."""
import importlib

import numpy as np
import pytest

from oracle import g753 as O
from oracle import ref753
from util753 import FIELDS, G, GROUPS, ffi, ints_to_array, projective_to_point

import bench
import bench_groth16

pytestmark = pytest.mark.gpu
params = importlib.import_module("ginger-lib_b200.params")


@pytest.fixture(scope="module")
def ctx():
    c = G.Context(0)           # raises if libg753.so or the GPU is missing: no CPU path
    yield c
    c.close()


def generator(group):
    C = GROUPS[group]
    F, k = C.F.base, C.F.k
    gen_m = params.GENERATOR_MONT[group]
    return (tuple(F.from_mont(v) for v in gen_m[:k]), tuple(F.from_mont(v) for v in gen_m[k:]))


def same_point(ctx, group, xyz_a, xyz_b):
    """two projective results name the same group element: normalised on the device, compared limb by limb"""
    k = ffi.GROUP_K[group]
    both = np.ascontiguousarray(np.stack([np.asarray(xyz_a).reshape(-1), np.asarray(xyz_b).reshape(-1)]))
    xy = np.zeros((2, 2 * k * 12), dtype=np.uint64)
    inf = np.zeros(2, dtype=np.uint8)
    ctx.lib.check(ctx.lib.batch_normalize(ctx.handle, group, ffi.ptr(both), 2, ffi.ptr(xy), ffi.ptr(inf)))
    return bool((xy[0] == xy[1]).all() and inf[0] == inf[1])


def test_msm_g1_2e22_config3(ctx):
    """BASELINE config 3 at full size: plain key and key with precomputed copies against the discrete-log
    prediction (sum s_i a_i mod r) * G, a 2^18-point prefix against the C++ restatement of
    variable_base.rs:10-83, and the host-scalar entry point against the device-scalar one"""
    group, log_n = ffi.MNT4_G1, 22
    C = GROUPS[group]
    n = 1 << log_n
    bases = ctx.generate_bases(group, n, 0x22C3)
    logs = G.Bases.generated_logs(n, 0x22C3)
    sc = bench.random_scalars(n, 0x22C4)
    sc[5] = 0
    sc[6] = 0
    sc[6, 0] = 1
    sc[7] = ints_to_array([C.r - 1])[0]
    want = C.mul(generator(group), bench.dot_mod(sc, logs, C.r))
    plain = G.VariableBaseMSM.multi_scalar_mul(bases, sc)
    assert projective_to_point(C, plain) == want
    m = 1 << 18
    coords = bases.download(0, m)
    prefix = G.VariableBaseMSM.multi_scalar_mul(bases, sc[:m])
    assert same_point(ctx, group, prefix, ref753.msm(group, coords, None, sc[:m]))
    bases.precompute(0)
    pre = G.VariableBaseMSM.multi_scalar_mul(bases, sc)
    assert ctx.last_msm_plan()["copies"] > 1
    assert projective_to_point(C, pre) == want
    prefix2 = G.VariableBaseMSM.multi_scalar_mul(bases, sc[:m])     # short slice of the precomputed key
    assert same_point(ctx, group, prefix, prefix2)
    bases.free()


def test_fft_2e22_all_modes_vs_cpp(ctx):
    """the 2^22 transform of the metric, all four modes, EVERY limb against the C++ restatement of
    domain.rs:120-179 / 305-416 (best_fft with the box's threads)"""
    field, log_n = ffi.FIELD_MNT4_FR, 22
    n = 1 << log_n
    rng = np.random.default_rng(2222)
    raw = rng.integers(0, 1 << 63, size=(n, 12), dtype=np.uint64)
    raw[:, 11] &= np.uint64(0xFFFF)
    dom = G.EvaluationDomain.new(field, n, ctx=ctx)
    for mode, run in ((ffi.FFT, dom.fft), (ffi.IFFT, dom.ifft), (ffi.COSET_FFT, dom.coset_fft),
                      (ffi.COSET_IFFT, dom.coset_ifft)):
        assert np.array_equal(run(raw), ref753.fft(field, raw, mode)), mode


@pytest.mark.parametrize("field", sorted(FIELDS))
def test_ntt_size_sweep_vs_cpp(ctx, field):
    """every domain size (the reference's fft/test.rs:9-43 walks log_n = 0..9; the pass splitter takes a
    different path for almost every log_n): 0..14 on mnt6753::Fr (its maximum), 0..24 on mnt4753::Fr;
    all four transforms up to 2^21 (2^22 has its own test), fft alone at 2^23 and 2^24"""
    sizes = list(range(15)) if field == ffi.FIELD_MNT6_FR else list(range(22)) + [23, 24]
    rng = np.random.default_rng(900 + field)
    for log_n in sizes:
        n = 1 << log_n
        raw = rng.integers(0, 1 << 63, size=(n, 12), dtype=np.uint64)
        raw[:, 11] &= np.uint64(0xFFFF)
        dom = G.EvaluationDomain.new(field, n, ctx=ctx)
        assert dom is not None and dom.size() == n
        modes = ((ffi.FFT, dom.fft), (ffi.IFFT, dom.ifft), (ffi.COSET_FFT, dom.coset_fft), (ffi.COSET_IFFT, dom.coset_ifft))
        if log_n > 21:
            modes = modes[:1]
        for mode, run in modes:
            assert np.array_equal(run(raw), ref753.fft(field, raw, mode)), (log_n, mode)


def test_msm_fq3_2e20_config4(ctx):
    """BASELINE config 4 at full size: MNT6-753 G2 (over Fq3), 2^20 points, against the discrete-log
    prediction, on the plain key and on the key with precomputed copies"""
    group, log_n = ffi.MNT6_G2, 20
    C = GROUPS[group]
    n = 1 << log_n
    bases = ctx.generate_bases(group, n, 0x20C4)
    logs = G.Bases.generated_logs(n, 0x20C4)
    sc = bench.random_scalars(n, 0x20C5)
    sc[9] = 0
    want = C.mul(generator(group), bench.dot_mod(sc, logs, C.r))
    assert projective_to_point(C, G.VariableBaseMSM.multi_scalar_mul(bases, sc)) == want
    bases.precompute(0)
    assert projective_to_point(C, G.VariableBaseMSM.multi_scalar_mul(bases, sc)) == want
    bases.free()


def test_groth16_2e20_config5(ctx):
    """BASELINE config 5 at full size: one create_proof over a 2^20 domain, A / B / C against the generator
    multiples prover.rs:270-337 prescribes for a key with known discrete logs (bench_groth16.verify_proof)"""
    res = bench_groth16.run(ctx, 20, steps=1, warmup=0, verify=True)
    assert res["verified"]


def skewed_scalars(n, r, seed):
    """half of the scalars 0 / 1 / small - the histogram real witnesses have (booleans, counters):
    zero scalars are skipped, ones take the reference's fast path (variable_base.rs:37-42), small values
    pile into a handful of buckets of window 0 and leave every other window empty"""
    sc = bench.random_scalars(n, seed)
    rng = np.random.default_rng(seed + 1)
    kind = rng.integers(0, 8, size=n)
    small = rng.integers(0, 1 << 16, size=n, dtype=np.uint64)
    for k, val in ((0, 0), (1, 1), (2, None), (3, 2)):
        idx = np.flatnonzero(kind == k)
        sc[idx] = 0
        sc[idx, 0] = small[idx] if val is None else np.uint64(val)
    return sc


@pytest.mark.parametrize("group,log_n", [(ffi.MNT4_G1, 20), (ffi.MNT4_G2, 16)])
def test_msm_skewed_witness(ctx, group, log_n):
    """0 / 1 / small-heavy scalars at size: against the discrete-log prediction (plain and precomputed key)
    and, on a 2^16 prefix, against the C++ restatement (which takes the reference's scalar == 1 path)"""
    C = GROUPS[group]
    n = 1 << log_n
    bases = ctx.generate_bases(group, n, 0x5E0 + group)
    logs = G.Bases.generated_logs(n, 0x5E0 + group)
    sc = skewed_scalars(n, C.r, 0x5E8 + group)
    want = C.mul(generator(group), bench.dot_mod(sc, logs, C.r))
    assert projective_to_point(C, G.VariableBaseMSM.multi_scalar_mul(bases, sc)) == want
    m = min(n, 1 << (16 if C.F.k == 1 else 12))
    got = G.VariableBaseMSM.multi_scalar_mul(bases, sc[:m])
    assert same_point(ctx, group, got, ref753.msm(group, bases.download(0, m), None, sc[:m]))
    bases.precompute(0)
    assert projective_to_point(C, G.VariableBaseMSM.multi_scalar_mul(bases, sc)) == want
    bases.free()


def test_msm_and_witness_map_reference_circuit(ctx):
    """the reference's own benchmark circuit (groth16/examples/snark-scalability/constraints.rs:19-91) at a
    2^18 domain: the witness map of its evaluation vectors every limb against the C++ restatement's
    transforms, and the MSM of its assignment (what the A / B / L queries are multiplied by) against the
    discrete-log prediction.  (Its values are 1, 1, 2, 2, 4, 8, 12, 96, ... for the first ~25 variables
    and fill the field from there on; b is 1 on every other row.)"""
    groth16 = importlib.import_module("ginger-lib_b200.groth16")
    F, field, group = O.MNT4_FR, ffi.FIELD_MNT4_FR, ffi.MNT4_G1
    C = GROUPS[group]
    log_n = 18
    n = 1 << log_n
    z, ea, eb, ec = bench_groth16.benchmark_circuit(n - 3, F.p)
    assert len(ea) == n and len(z) == n
    # ---- witness map vs the restated transforms ---------------------------------------------------
    a, b, c = (bench_groth16.to_mont(ctx, field, bench_groth16.ints_to_limbs(v)) for v in (ea, eb, ec))
    h = groth16.witness_map(ctx, field, a, b, c, 0, 0, 0)
    fa = ref753.fft(field, ref753.fft(field, a, ffi.IFFT), ffi.COSET_FFT)
    fb = ref753.fft(field, ref753.fft(field, b, ffi.IFFT), ffi.COSET_FFT)
    fc = ref753.fft(field, ref753.fft(field, c, ffi.IFFT), ffi.COSET_FFT)
    ab = ref753.field_op(field, 2, ref753.field_op(field, 0, fa, fb), fc)          # a * b - c
    zinv = G.EvaluationDomain.new(field, n, ctx=ctx).vanishing_on_coset_inv
    ab = ref753.field_op(field, 0, ab, np.tile(zinv.reshape(1, 12), (n, 1)))
    want_h = ref753.fft(field, ab, ffi.COSET_IFFT)
    assert np.array_equal(h[:n - 1], want_h[:n - 1])
    assert not h[n - 1:].any()                      # d1 = d2 = d3 = 0 (r1cs_to_qap.rs:125-132)
    # the satisfied circuit's quotient is a polynomial: a * b - c vanishes on the domain, so the
    # coefficient the reference drops (ab[n - 1], :163-166) is zero
    assert not want_h[n - 1].any()
    # ---- MSM over the assignment --------------------------------------------------------------------
    sc = bench_groth16.ints_to_limbs(z)
    bases = ctx.generate_bases(group, n, 0xC1C)
    logs = G.Bases.generated_logs(n, 0xC1C)
    want = C.mul(generator(group), bench.dot_mod(sc, logs, C.r))
    assert projective_to_point(C, G.VariableBaseMSM.multi_scalar_mul(bases, sc)) == want
    bases.precompute(0)
    assert projective_to_point(C, G.VariableBaseMSM.multi_scalar_mul(bases, sc)) == want
    m = 1 << 15
    got = G.VariableBaseMSM.multi_scalar_mul(bases, sc[:m])
    assert same_point(ctx, group, got, ref753.msm(group, bases.download(0, m), None, sc[:m]))
    bases.free()
