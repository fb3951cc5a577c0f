"""The pairing / key-generation / verification restatement (oracle/pairing753.py) pinned: its derived
parameters against the reference's literals, the pairing's defining properties on the reference's
generators, and generate -> prove -> verify on the oracle alone (the reference's acceptance test,
proof-systems/src/groth16/test.rs:216-301).  The GPU prover is put through the same verifier in
tests/test_gpu_groth16.py and, on the host-emulation build, in tests/test_groth16_emul.py."""
import json
import os

import pytest

from oracle import g753 as O
from oracle import pairing753 as PR

HERE = os.path.dirname(os.path.abspath(__file__))
PAIRING = json.load(open(os.path.join(HERE, "golden", "reference_pairing.json")))
PARAMS = json.load(open(os.path.join(HERE, "golden", "reference_params.json")))


def generators(eng):
    """the reference's prime-subgroup generators of G1 and G2 (curves/mnt{4,6}753/g1.rs, g2.rs)"""
    import importlib
    from util753 import ffi, g2_generator
    params = importlib.import_module("ginger-lib_b200.params")
    F = eng.g1.F.base
    gm = params.GENERATOR_MONT[ffi.MNT4_G1 if eng.name == "mnt4" else ffi.MNT6_G1]
    P = ((F.from_mont(gm[0]),), (F.from_mont(gm[1]),))
    assert eng.g1.on_curve(P) and eng.g1.mul(P, eng.r) is None
    return P, g2_generator(eng.g2)


@pytest.mark.parametrize("name,key", [("mnt4", "curves_mnt4753_mod"), ("mnt6", "curves_mnt6753_mod")])
def test_derived_parameters_equal_the_reference_literals(name, key):
    eng = PR.engine(name)
    ref = PAIRING[key]
    assert eng.loop_count == int(ref["ATE_LOOP_COUNT"], 16)
    assert eng.wnaf == ref["WNAF"]
    assert eng.loop_count_neg == ref["ATE_IS_LOOP_COUNT_NEG"]
    consts = PARAMS[key]["consts"]
    assert eng.m1 == int(consts["FINAL_EXPONENT_LAST_CHUNK_1"][0]["value"], 16)
    assert abs(eng.m0) == int(consts["FINAL_EXPONENT_LAST_CHUNK_ABS_OF_W0"][0]["value"], 16)
    assert (eng.m0 < 0) == ref["FINAL_EXPONENT_LAST_CHUNK_W0_IS_NEG"]
    # TWIST_COEFF_A = a * twist^2 is the coefficient of the G2 curve the oracle uses
    F = eng.g1.F.base
    ca = F.from_mont(int(consts["TWIST_COEFF_A"][0]["value"], 16))
    assert ca in eng.g2.a and sum(1 for v in eng.g2.a if v) == 1
    assert tuple(eng.Fe.mul(eng.twist_sq, tuple([eng.g1.a[0]] + [0] * (eng.ke - 1)))) == tuple(eng.g2.a)


@pytest.mark.parametrize("name", ["mnt4", "mnt6"])
def test_pairing_properties(name):
    eng = PR.engine(name)
    Fk = eng.Fk
    P, Q = generators(eng)
    e = eng.pairing(P, Q)
    assert e != Fk.one                                   # non-degenerate
    assert Fk.pow(e, eng.r) == Fk.one                    # lands in the order-r subgroup
    a, b = 0x1234567 << 300 | 5, 0xABCDEF << 500 | 3
    assert eng.pairing(eng.g1.mul(P, a), eng.g2.mul(Q, b)) == Fk.pow(e, a * b % eng.r)   # bilinear
    assert eng.pairing(eng.g1.neg(P), Q) == Fk.inv(e)
    # the product form the verifier uses
    assert eng.multi_pairing([(P, Q), (eng.g1.neg(P), Q)]) == Fk.one
    # Frobenius is the q-th power
    x = [(7 * i + 3) % eng.p for i in range(Fk.k)]
    assert Fk.frobenius(x, 1) == Fk.pow(x, eng.p)


def tiny_groth16(name, num_constraints=5, seed=0x7e57):
    """the reference benchmark circuit, a CRS from fixed toxic waste, a satisfying assignment"""
    import bench_groth16
    eng = PR.engine(name)
    F = eng.fr
    at, bt, ct, ni, n_aux = PR.benchmark_circuit_matrices(num_constraints)
    z, ea, eb, ec = bench_groth16.benchmark_circuit(num_constraints, F.p)
    assert len(z) == ni + n_aux
    for row_a, row_b, row_c, va, vb, vc in zip(at, bt, ct, ea, eb, ec):     # the matrices and the evaluations agree
        dot = lambda row: sum(cf * z[v] for v, cf in row.items()) % F.p
        assert (dot(row_a), dot(row_b), dot(row_c)) == (va, vb, vc)
        assert va * vb % F.p == vc
    rng = O.SplitMix64(seed)
    alpha, beta, gamma, delta, tau = (O.random_field_element(rng, F) for _ in range(5))
    P, Q = generators(eng)
    g1 = eng.g1.mul(P, O.random_field_element(rng, F))
    g2 = eng.g2.mul(Q, O.random_field_element(rng, F))
    key, vk, n = PR.generate_parameters(eng, at, bt, ct, ni, n_aux, alpha, beta, gamma, delta, tau, g1, g2)
    pad = n - len(ea)
    return eng, key, vk, ni, z, ea + [0] * pad, eb + [0] * pad, ec + [0] * pad


@pytest.mark.parametrize("name", ["mnt4"])      # MNT6 goes through the same code in test_groth16_emul.py (CPU) / test_gpu_groth16.py
def test_generate_prove_verify_on_the_oracle(name):
    eng, key, vk, ni, z, a, b, c = tiny_groth16(name)
    F = eng.fr
    rng = O.SplitMix64(0x51)
    r, s = O.random_field_element(rng, F), O.random_field_element(rng, F)
    h = O.witness_map(F, a, b, c, 0, 0, 0)
    proof = O.groth16_create_proof(key, ni, z, h, r, s)
    assert PR.verify_proof(eng, vk, proof, z[1:ni])
    assert not PR.verify_proof(eng, vk, proof, [z[1], (z[2] + 1) % F.p])        # wrong public input
    bad = (proof[0], proof[1], eng.g1.add(proof[2], key.delta_g1))
    assert not PR.verify_proof(eng, vk, bad, z[1:ni])                            # tampered proof
