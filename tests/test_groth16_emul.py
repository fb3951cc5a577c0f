"""CPU-tier check of the Groth16 prover sequencing (ginger-lib_b200/groth16.py over the C ABI)
against the oracle's restatement of witness_map / create_proof, using the TEST-ONLY host-emulation
build of the kernels.  The same assertions run on the real library in tests/test_gpu_groth16.py."""
import numpy as np
import pytest

from oracle import g753 as O
from test_pipeline_emul import ctx  # noqa: F401  (fixture: emulation library context)
from util753 import G, array_field, array_to_ints, ffi, field_array, ints_to_array, points_to_arrays, sample_points

groth16 = __import__("importlib").import_module("ginger-lib_b200.groth16")


# the two pairing engines of the cycle: (G1 curve, G2 curve, scalar field, C-ABI ids)
ENGINES = {
    "mnt4": (O.MNT4_G1, O.MNT4_G2, O.MNT4_FR, ffi.MNT4_G1, ffi.MNT4_G2, ffi.FIELD_MNT4_FR),
    "mnt6": (O.MNT6_G1, O.MNT6_G2, O.MNT6_FR, ffi.MNT6_G1, ffi.MNT6_G2, ffi.FIELD_MNT6_FR),
}


def tiny_instance(seed, n=8, ni=3, n_aux=6, engine="mnt4"):
    """a synthetic proving key + witness of the shapes generator.rs:225-319 produces:
    a/b queries over all variables, h_query of domain_size - 1, l_query over the aux variables"""
    C1, C2, F = ENGINES[engine][:3]
    rng = O.SplitMix64(seed)
    n_vars = ni + n_aux
    g1pts = sample_points(C1, 3 + 2 * n_vars + (n - 1) + n_aux, seed + 1)
    g2pts = sample_points(C2, 2 + n_vars, seed + 2)
    it1, it2 = iter(g1pts), iter(g2pts)
    take = lambda it, m: [next(it) for _ in range(m)]
    key = O.Groth16Key(C1, C2, next(it1), next(it1), next(it2), next(it1), next(it2),
                       take(it1, n_vars), take(it1, n_vars), take(it2, n_vars), take(it1, n - 1), take(it1, n_aux))
    key.b_g1_query[n_vars - 1] = None   # an infinity base inside a query (zero column of B)
    key.b_g2_query[n_vars - 1] = None
    z = [1] + [O.random_field_element(rng, F) for _ in range(n_vars - 1)]
    if n_vars > 7:
        z[5] = 0                        # zero / one / p-1 witness values
        z[6] = 1
        z[7] = F.p - 1
    nc = n - ni                         # num_constraints + num_inputs = domain size (r1cs_to_qap.rs:105-107)
    a = [O.random_field_element(rng, F) for _ in range(nc)] + [1] + z[1:ni]   # :111-119
    b = [O.random_field_element(rng, F) for _ in range(nc)] + [0] * ni
    c = [x * y % F.p for x, y in zip(a[:nc], b[:nc])] + [0] * ni
    return key, z, a, b, c


def upload(cx, key, ni, precompute=0, engine="mnt4"):
    C1, C2, _, g1, g2, field = ENGINES[engine]
    one = lambda C, P: points_to_arrays(C, [P])[0][0]
    q = lambda C, pts: points_to_arrays(C, pts)
    return groth16.Parameters(cx, g1, g2, field, one(C1, key.alpha_g1), one(C1, key.beta_g1),
                              one(C2, key.beta_g2), one(C1, key.delta_g1), one(C2, key.delta_g2),
                              q(C1, key.a_query), q(C1, key.b_g1_query), q(C2, key.b_g2_query),
                              q(C1, key.h_query), q(C1, key.l_query), ni, precompute=precompute)


def affine_of(curve, xy, inf):
    if inf:
        return None
    F, k = curve.F.base, curve.F.k
    vals = array_to_ints(np.asarray(xy).reshape(-1, 12))
    assert all(v < F.p for v in vals), "non-canonical proof coordinate"
    c = [F.from_mont(v) for v in vals]
    return (tuple(c[:k]), tuple(c[k:]))


def check_instance(cx, seed, n, ni, n_aux, d1, d2, d3, r, s, engine="mnt4"):
    C1, C2, F, _, _, field = ENGINES[engine]
    key, z, a, b, c = tiny_instance(seed, n, ni, n_aux, engine)
    h_ref = O.witness_map(F, a, b, c, d1, d2, d3)
    h = groth16.witness_map(cx, field, field_array(F, a), field_array(F, b), field_array(F, c), d1, d2, d3)
    assert array_field(F, h) == h_ref
    want = O.groth16_create_proof(key, ni, z, h_ref, r, s)
    params = upload(cx, key, ni, engine=engine)
    for _ in range(2):       # the second proof reuses the key's device workspace
        proof = groth16.create_proof(params, field_array(F, z), field_array(F, a), field_array(F, b), field_array(F, c),
                                     d1, d2, d3, r, s)
        got = (affine_of(C1, proof.a, proof.infinity[0]), affine_of(C2, proof.b, proof.infinity[1]),
               affine_of(C1, proof.c, proof.infinity[2]))
        assert got == want
    params.free()


def test_witness_map_and_proof_tiny(ctx):  # noqa: F811
    F = O.MNT4_FR
    check_instance(ctx, 0x6107, 8, 3, 6, 0, 0, 0, 0x1234567 << 600, (F.p - 3))
    # non-zero d's exercise the reference's h initialisation quirk (r1cs_to_qap.rs:124-134)
    check_instance(ctx, 0x6108, 4, 2, 3, 5, 7, 11, 3, 4)


def test_proof_tiny_mnt6(ctx):  # noqa: F811
    """the other engine of the cycle: G2 over Fq3, scalar field of two-adicity 15"""
    check_instance(ctx, 0x6601, 4, 2, 3, 1, 2, 3, O.MNT6_FR.p - 5, 0x77 << 700, engine="mnt6")


def test_proof_tiny_with_the_addition_tree(ctx, monkeypatch):  # noqa: F811
    """the form every MSM of a full-size proof runs (msm.cuh k_tree_round), forced on the tiny instance: G1 and
    G2 (Fq2) queries, the duplicate / zero assignments of tiny_instance included"""
    monkeypatch.setenv("G753_MSM_AFFINE", "1")
    cx = G.Context(0, library=ctx.lib)
    check_instance(cx, 0x6109, 4, 2, 3, 5, 7, 11, 3, 4)
    cx.close()


@pytest.mark.parametrize("engine", ["mnt6"])     # MNT4: test_oracle_pairing.py (oracle) and test_gpu_groth16.py (GPU, both engines)
def test_proof_verifies_with_pairing(ctx, engine):  # noqa: F811
    import shared_checks
    shared_checks.check_proof_verifies_with_pairing(ctx, engine)
