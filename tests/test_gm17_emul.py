"""CPU-tier check of the GM17 prover sequencing (ginger-lib_b200/gm17.py over the C ABI) against the
oracle's restatement of R1CStoSAP::witness_map / gm17 create_proof, using the TEST-ONLY host-emulation
build of the kernels.  The same assertions run on the real library in tests/test_gpu_zz_gm17.py."""
import numpy as np
import pytest

from oracle import g753 as O
from test_groth16_emul import ENGINES, affine_of
from test_pipeline_emul import ctx  # noqa: F401  (fixture: emulation library context)
from util753 import G, array_field, ffi, field_array, points_to_arrays, sample_points

gm17 = __import__("importlib").import_module("ginger-lib_b200.gm17")


def tiny_instance(seed, n=8, ni=2, n_aux=3, nc=2, engine="mnt4", inf_head=False):
    """a synthetic GM17 proving key + SAP witness of the shapes gm17/generator.rs:222-343 produces:
    a / b / c_2 queries over all SAP variables, c_query_1 over the non-input ones, g_gamma2_z_t of
    domain size; domain >= 2 nc + 2 (ni - 1) + 1 (r1cs_to_sap.rs:153-157)"""
    C1, C2, F = ENGINES[engine][:3]
    assert n >= 2 * nc + 2 * (ni - 1) + 1
    rng = O.SplitMix64(seed)
    n_vars = ni + n_aux + nc + (ni - 1)                   # sap_num_variables
    g1pts = sample_points(C1, 3 + 2 * n_vars + (n_vars - ni) + n, seed + 1)
    g2pts = sample_points(C2, 1 + n_vars, seed + 2)
    it1, it2 = iter(g1pts), iter(g2pts)
    take = lambda it, m: [next(it) for _ in range(m)]
    key = O.GM17Key(C1, C2, take(it1, n_vars), take(it2, n_vars), take(it1, n_vars - ni), take(it1, n_vars),
                    next(it1), next(it2), next(it1), next(it1), take(it1, n))
    key.b_query[n_vars - 1] = None      # an infinity base inside a query
    key.c_query_2[ni] = None
    if inf_head:                        # query[0] at infinity: the reference adds it like any other point
        key.a_query[0] = None
        key.c_query_2[0] = None
        key.g_gamma2_z_t[0] = None
    z = [1] + [O.random_field_element(rng, F) for _ in range(n_vars - 1)]
    z[ni] = 0                           # zero / one / p-1 witness values
    z[ni + 1] = 1
    z[ni + 2] = F.p - 1
    used = 2 * nc + 2 * (ni - 1) + 1
    a = [O.random_field_element(rng, F) for _ in range(used)] + [0] * (n - used)
    c = [O.random_field_element(rng, F) for _ in range(used)] + [0] * (n - used)
    return key, z, a, c


def upload(cx, key, ni, engine="mnt4"):
    C1, C2, _, g1, g2, field = ENGINES[engine]
    one = lambda C, P: points_to_arrays(C, [P])[0][0]
    q = lambda C, pts: points_to_arrays(C, pts)
    return gm17.Parameters(cx, g1, g2, field, q(C1, key.a_query), q(C2, key.b_query), q(C1, key.c_query_1),
                           q(C1, key.c_query_2), one(C1, key.g_gamma_z), one(C2, key.h_gamma_z),
                           one(C1, key.g_ab_gamma_z), one(C1, key.g_gamma2_z2), q(C1, key.g_gamma2_z_t), ni)


def check_instance(cx, seed, n, ni, n_aux, nc, d1, d2, r, engine="mnt4", inf_head=False):
    C1, C2, F, _, _, field = ENGINES[engine]
    key, z, a, c = tiny_instance(seed, n, ni, n_aux, nc, engine, inf_head)
    h_ref = O.sap_witness_map(F, a, c, d1, d2)
    h = gm17.witness_map(cx, field, field_array(F, a), field_array(F, c), d1, d2)
    assert array_field(F, h) == h_ref
    want = O.gm17_create_proof(key, ni, z, h_ref, d1, d2, r, F)
    params = upload(cx, key, ni, engine=engine)
    proof = gm17.create_proof(params, field_array(F, z), field_array(F, a), field_array(F, c), d1, d2, r)
    got = (affine_of(C1, proof.a, proof.infinity[0]), affine_of(C2, proof.b, proof.infinity[1]),
           affine_of(C1, proof.c, proof.infinity[2]))
    assert got == want
    params.free()


def test_sap_witness_map_and_proof_tiny(ctx):  # noqa: F811
    F = O.MNT4_FR
    check_instance(ctx, 0x1707, 8, 2, 3, 2, 5, 7, 0x1234567 << 600)
    check_instance(ctx, 0x1708, 16, 3, 4, 4, 0, 0, F.p - 3)
    check_instance(ctx, 0x1709, 4, 1, 3, 1, F.p - 1, 3, 0)
    check_instance(ctx, 0x170A, 8, 2, 3, 2, 5, 7, 11, inf_head=True)


def test_proof_tiny_mnt6(ctx):  # noqa: F811
    """the other engine of the cycle: G2 over Fq3, scalar field of two-adicity 15"""
    check_instance(ctx, 0x1761, 8, 2, 3, 2, 1, 2, O.MNT6_FR.p - 5, engine="mnt6")
