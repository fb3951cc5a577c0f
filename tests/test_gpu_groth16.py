"""Groth16 prover path on the GPU (BASELINE config 5 shape at reduced size): witness_map and
create_proof through the C ABI against (a) the Python oracle at tiny sizes, bit-exact proofs, (b) the
C++ restatement's transforms at 2^14 (witness map, every limb), (c) the discrete-log property of a
synthetic key at 2^16 (the whole proof)."""
import importlib

import numpy as np
import pytest

from oracle import g753 as O
from test_groth16_emul import affine_of, check_instance
from util753 import G, GROUPS, array_to_ints, ffi, ints_to_array, projective_to_point

groth16 = importlib.import_module("ginger-lib_b200.groth16")
params_mod = importlib.import_module("ginger-lib_b200.params")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = G.Context(0)
    yield c
    c.close()


def test_proof_tiny_vs_oracle(ctx):
    F = O.MNT4_FR
    check_instance(ctx, 0x6107, 8, 3, 6, 0, 0, 0, 0x1234567 << 600, F.p - 3)
    check_instance(ctx, 0x6108, 4, 2, 3, 5, 7, 11, 3, 4)
    check_instance(ctx, 0x6109, 32, 3, 40, 1, 2, 3, F.p - 1, 12345)   # more variables than constraints
    check_instance(ctx, 0x6601, 8, 2, 9, 1, 2, 3, O.MNT6_FR.p - 5, 0x77 << 700, engine="mnt6")   # G2 over Fq3


def _mont_random(n, seed):
    import bench
    raw = bench.random_scalars(n, seed)
    raw[:, 11] &= np.uint64(0xFFFF)        # < p: a valid Montgomery representation
    return raw


def test_witness_map_2e14_vs_cpp_restatement(ctx):
    """every limb of h against the C++ restatement's ifft / coset_fft / coset_ifft chain"""
    from oracle import ref753
    F = O.MNT4_FR
    field = ffi.FIELD_MNT4_FR
    n = 1 << 14
    a, b, c = (_mont_random(n, 0x100 + i) for i in range(3))
    d1, d2, d3 = 9, 8, 7
    h = groth16.witness_map(ctx, field, a, b, c, d1, d2, d3)
    fa = ref753.fft(field, ref753.fft(field, a, ffi.IFFT), ffi.COSET_FFT)
    fb = ref753.fft(field, ref753.fft(field, b, ffi.IFFT), ffi.COSET_FFT)
    fc = ref753.fft(field, ref753.fft(field, c, ffi.IFFT), ffi.COSET_FFT)
    ab = ref753.field_op(field, ffi.OP_SUB, ref753.field_op(field, ffi.OP_MUL, fa, fb), fc)
    zinv = F.to_mont(pow((pow(F.generator, n, F.p) - 1) % F.p, -1, F.p))
    ab = ref753.field_op(field, ffi.OP_MUL, ab, np.tile(ints_to_array([zinv]), (n, 1)))
    q = ref753.fft(field, ab, ffi.COSET_IFFT)
    want = np.zeros((n + 1, 12), dtype=np.uint64)
    want[:n - 1] = q[:n - 1]
    d1d2 = d1 * d2 % F.p
    want[n] = ints_to_array([F.to_mont(d1d2)])[0]
    h0 = (F.from_mont(array_to_ints(q[:1])[0]) - d3 - d1d2) % F.p
    want[0] = ints_to_array([F.to_mont(h0)])[0]
    assert (h == want).all()


def test_proof_2e16_discrete_log_property(ctx):
    """synthetic key whose every base is a known multiple of the generator: the proof's A, B, C must be
    the generator multiples the prover equations give (prover.rs:270-337)"""
    import bench
    F = O.MNT4_FR
    rmod = F.p
    field = ffi.FIELD_MNT4_FR
    g1, g2 = ffi.MNT4_G1, ffi.MNT4_G2
    log_n = 16
    n = 1 << log_n
    ni = 3
    n_aux = n - ni                        # num_aux = num_constraints (snark-scalability constraints.rs:19-91)
    n_vars = ni + n_aux
    seeds = {"a": 11, "b1": 12, "b2": 13, "h": 14, "l": 15}
    Ba = ctx.generate_bases(g1, n_vars, seeds["a"])
    Bb1 = ctx.generate_bases(g1, n_vars, seeds["b1"])
    Bb2 = ctx.generate_bases(g2, n_vars, seeds["b2"])
    Bh = ctx.generate_bases(g1, n - 1, seeds["h"])
    Bl = ctx.generate_bases(g1, n_aux, seeds["l"])
    vk = ctx.generate_bases(g1, 3, 21).download()           # alpha, beta, delta in G1
    vk2 = ctx.generate_bases(g2, 2, 22).download()          # beta, delta in G2
    vlog, vlog2 = G.Bases.generated_logs(3, 21), G.Bases.generated_logs(2, 22)
    P = groth16.Parameters(ctx, g1, g2, field, vk[0], vk[1], vk2[0], vk[2], vk2[1], Ba, Bb1, Bb2, Bh, Bl, ni, precompute=4)
    z = _mont_random(n_vars, 0x200)
    z[0] = ints_to_array([F.to_mont(1)])[0]
    a, b, c = (_mont_random(n, 0x300 + i) for i in range(3))
    r, s = 0xABCDEF << 500, F.p - 12345
    d1 = d2 = d3 = 0
    t = {}
    proof = groth16.create_proof(P, z, a, b, c, d1, d2, d3, r, s, timings=t)
    h = groth16.witness_map(ctx, field, a, b, c, d1, d2, d3)
    # canonical scalars on the host (test-side arithmetic): into_repr of z and h
    zc = ctx_from_mont(ctx, field, z)
    hc = ctx_from_mont(ctx, field, h)
    la, lb1, lb2 = (G.Bases.generated_logs(n_vars, seeds[k]) for k in ("a", "b1", "b2"))
    lh, ll = G.Bases.generated_logs(n - 1, seeds["h"]), G.Bases.generated_logs(n_aux, seeds["l"])
    zc1 = zc.copy()
    zc1[0] = ints_to_array([1])[0]
    alpha, beta, delta = (int(v) for v in vlog)
    beta2, delta2 = (int(v) for v in vlog2)
    A = (r * delta + bench.dot_mod(zc1, la, rmod) + alpha) % rmod
    B1 = (s * delta + bench.dot_mod(zc1, lb1, rmod) + beta) % rmod
    B2 = (s * delta2 + bench.dot_mod(zc1, lb2, rmod) + beta2) % rmod
    C = (s * A + r * B1 - r * s * delta + bench.dot_mod(zc[ni:], ll, rmod) + bench.dot_mod(hc[:n - 1], lh, rmod)) % rmod

    def gen_mul(group, k):
        C_ = GROUPS[group]
        Fb, kk = C_.F.base, C_.F.k
        gen_m = params_mod.GENERATOR_MONT[group]
        gen = np.stack([bench.int_to_limbs(v) for v in gen_m]).reshape(-1)
        out = np.zeros((3, kk * 12), dtype=np.uint64)
        ctx.lib.check(ctx.lib.point_op(ctx.handle, group, 2, ffi.ptr(gen), ffi.ptr(bench.int_to_limbs(k)), ffi.ptr(out)))
        return projective_to_point(C_, out)

    assert affine_of(O.MNT4_G1, proof.a, proof.infinity[0]) == gen_mul(g1, A)
    assert affine_of(O.MNT4_G2, proof.b, proof.infinity[1]) == gen_mul(g2, B2)
    assert affine_of(O.MNT4_G1, proof.c, proof.infinity[2]) == gen_mul(g1, C)
    P.free()


def ctx_from_mont(ctx, field, arr):
    """into_repr on the host (independent of the library): Montgomery limbs -> canonical limbs"""
    F = O.MNT4_FR if field == ffi.FIELD_MNT4_FR else O.MNT6_FR
    return ints_to_array([F.from_mont(v) for v in array_to_ints(arr)])


@pytest.mark.parametrize("engine", ["mnt4", "mnt6"])
def test_proof_verifies_with_pairing(ctx, engine):
    """generate (oracle) -> prove (this library, on the GPU) -> verify (pairing, oracle): the reference's
    acceptance test groth16/test.rs:216-301 with an arbiter independent of the prover"""
    import shared_checks
    shared_checks.check_proof_verifies_with_pairing(ctx, engine)
    shared_checks.check_proof_verifies_with_pairing(ctx, engine, precompute=0)
