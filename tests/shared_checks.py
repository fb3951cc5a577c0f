"""Checks that run on BOTH tiers: against the TEST-ONLY host-emulation build of the C ABI in the CPU
suite (tests/test_pipeline_emul.py) and against the real libg753.so on a B200 (tests/test_gpu_parity.py).
Each takes the tier's context."""
import json
import os

import numpy as np
import pytest

from oracle import g753 as O
from util753 import (G, GROUPS, array_to_ints, ffi, ints_to_array, points_to_arrays, projective_to_point, sample_points,
                     sample_scalars)

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "reference_kat.json")))

# group id -> (KAT file key, scope prefix, oracle tower)
EXT_OF_GROUP = {ffi.MNT4_G2: ("fields_mnt4753_tests", "fq2", O.FQ2_MNT4),
                ffi.MNT6_G2: ("fields_mnt6753_tests", "fq3", O.FQ3_MNT6)}


def ext_op(ctx, group, lanes, op, E, a, b=None):
    """elements as tuples of canonical ints -> g753_ext_op -> tuples of canonical ints"""
    F, k = E.base, E.k

    def pack(elems):
        return ints_to_array([F.to_mont(c) for e in elems for c in e])

    aa = pack(a)
    bb = pack(b) if b is not None else None
    out = np.zeros_like(aa)
    ctx.lib.check(ctx.lib.ext_op(ctx.handle, group, lanes, op, ffi.ptr(aa), ffi.ptr(bb), ffi.ptr(out), len(a)))
    vals = array_to_ints(out)
    assert all(v < F.p for v in vals), "non-canonical coordinate returned"
    vals = [F.from_mont(v) for v in vals]
    return [tuple(vals[i:i + k]) for i in range(0, len(vals), k)]


def check_reference_ext_kats(ctx, group, lanes):
    """the reference's own Fq2 / Fq3 known answers (fields/mnt4753/tests.rs:1071-1867,
    fields/mnt6753/tests.rs:1277-2378) through the tower code the kernels run"""
    key, name, E = EXT_OF_GROUP[group]
    k = E.k
    t = KAT[key]["tests"]

    def vals(scope):
        v = [int(x["value"], 16) for x in t["test_%s_%s" % (name, scope)]]
        return [tuple(v[i:i + k]) for i in range(0, len(v), k)]

    a, b, c = vals("mul")
    assert ext_op(ctx, group, lanes, 0, E, [a], [b]) == [c]
    assert ext_op(ctx, group, lanes, 20, E, [a], [b]) == [c]
    assert ext_op(ctx, group, lanes, 21, E, [a], [b]) == [c]
    raw = [int(x["value"], 16) for x in t["test_%s_squaring" % name]]
    a, c = tuple(raw[-2 * k:-k]), tuple(raw[-k:])
    assert ext_op(ctx, group, lanes, 3, E, [a]) == [c]
    assert ext_op(ctx, group, lanes, 23, E, [a]) == [c]
    a, c = vals("inverse")
    assert ext_op(ctx, group, lanes, 5, E, [a]) == [c]
    a, b, c = vals("addition")
    assert ext_op(ctx, group, lanes, 1, E, [a], [b]) == [c]
    a, b, c = vals("subtraction")
    assert ext_op(ctx, group, lanes, 2, E, [a], [b]) == [c]
    a, c = vals("negation")
    assert ext_op(ctx, group, lanes, 4, E, [a]) == [c]
    a, c = vals("doubling")
    assert ext_op(ctx, group, lanes, 12, E, [a]) == [c]


def check_ext_ops_random(ctx, group, lanes, count=70):
    """batched random + edge elements against the oracle tower (one column per pair, several blocks)"""
    _, _, E = EXT_OF_GROUP[group]
    F, k = E.base, E.k
    rng = O.SplitMix64(0xE87 + group + 16 * lanes)
    edge = [E.zero(), E.one(), tuple([F.p - 1] * k), (0,) * (k - 1) + (1,), (F.p - 1,) + (0,) * (k - 1)]
    a = edge + [tuple(O.random_field_element(rng, F) for _ in range(k)) for _ in range(count)]
    b = list(reversed(edge)) + [tuple(O.random_field_element(rng, F) for _ in range(k)) for _ in range(count)]
    assert ext_op(ctx, group, lanes, 0, E, a, b) == [E.mul(x, y) for x, y in zip(a, b)]
    assert ext_op(ctx, group, lanes, 20, E, a, b) == [E.mul(x, y) for x, y in zip(a, b)]
    assert ext_op(ctx, group, lanes, 21, E, a, b) == [E.mul(x, y) for x, y in zip(a, b)]
    assert ext_op(ctx, group, lanes, 1, E, a, b) == [E.add(x, y) for x, y in zip(a, b)]
    assert ext_op(ctx, group, lanes, 2, E, a, b) == [E.sub(x, y) for x, y in zip(a, b)]
    assert ext_op(ctx, group, lanes, 3, E, a) == [E.sqr(x) for x in a]
    assert ext_op(ctx, group, lanes, 23, E, a) == [E.sqr(x) for x in a]
    assert ext_op(ctx, group, lanes, 4, E, a) == [E.neg(x) for x in a]
    assert ext_op(ctx, group, lanes, 12, E, a) == [E.add(x, x) for x in a]
    nz = [x for x in a if not E.is_zero(x)][:12]
    assert ext_op(ctx, group, lanes, 5, E, nz) == [E.inv(x) for x in nz]


def wire_of(C, pts):
    k = C.F.k
    wire = b""
    for P in pts:
        if P is None:
            wire += O.int_to_bytes96(0) * k + O.int_to_bytes96(1) + O.int_to_bytes96(0) * (k - 1) + b"\x01"
        else:
            wire += b"".join(O.int_to_bytes96(c) for c in P[0]) + b"".join(O.int_to_bytes96(c) for c in P[1]) + b"\x00"
    return wire


def check_wire_rejects_malformed(ctx, group):
    """GroupAffine::read fails on a coordinate >= the modulus (Fp768::read, fp_768.rs:791-803) and on a
    flag byte other than 0 / 1 (bool::read, bytes.rs:227-237); so does the loader - and a failed load
    leaves no handle behind"""
    C = GROUPS[group]
    k = C.F.k
    p = C.F.base.p
    pts = sample_points(C, 5, 0x3E0 + group)
    good = wire_of(C, pts)
    rec = 2 * k * 96 + 1
    b = G.Bases.from_wire(ctx, group, good)          # the well-formed key loads
    b.free()
    # a coordinate equal to the modulus / above it / all ones
    for bad_val in (p, p + 5, (1 << 768) - 1):
        w = bytearray(good)
        off = 3 * rec + 96 * (2 * k - 1)             # last coordinate element of point 3
        w[off:off + 96] = bad_val.to_bytes(96, "little")
        with pytest.raises(ffi.G753Error) as ei:
            G.Bases.from_wire(ctx, group, bytes(w))
        assert ei.value.code == ffi.ERR_BAD_ARG and "modulus" in str(ei.value)
    # an infinity flag that is neither 0 nor 1
    w = bytearray(good)
    w[2 * rec - 1] = 2
    with pytest.raises(ffi.G753Error) as ei:
        G.Bases.from_wire(ctx, group, bytes(w))
    assert ei.value.code == ffi.ERR_BAD_ARG and "flag" in str(ei.value)
    # p - 1 is the largest legal coordinate value
    w = bytearray(good)
    w[0:96] = (p - 1).to_bytes(96, "little")
    G.Bases.from_wire(ctx, group, bytes(w)).free()


def check_proof_verifies_with_pairing(ctx, name, precompute=1):
    """the reference's acceptance test (proof-systems/src/groth16/test.rs:216-301): a CRS generated for the
    reference's benchmark circuit (generator.rs restated in oracle/pairing753.py), the proof made by THIS
    library, accepted by the pairing check of verifier.rs:18-44 - an arbiter that shares no code with the
    prover - and rejected for another public input"""
    import importlib
    import test_groth16_emul as T16
    import test_oracle_pairing as TP
    from oracle import pairing753 as PR
    from util753 import field_array
    groth16 = importlib.import_module("ginger-lib_b200.groth16")
    eng, key, vk, ni, z, a, b, c = TP.tiny_groth16(name, num_constraints=5, seed=0x7e58)
    C1, C2, F = T16.ENGINES[name][:3]
    rng = O.SplitMix64(0x52)
    r, s = O.random_field_element(rng, F), O.random_field_element(rng, F)
    params = T16.upload(ctx, key, ni, precompute=precompute, engine=name)
    proof = groth16.create_proof(params, field_array(F, z), field_array(F, a), field_array(F, b), field_array(F, c),
                                 0, 0, 0, r, s)
    got = (T16.affine_of(C1, proof.a, proof.infinity[0]), T16.affine_of(C2, proof.b, proof.infinity[1]),
           T16.affine_of(C1, proof.c, proof.infinity[2]))
    params.free()
    assert PR.verify_proof(eng, vk, got, z[1:ni])
    assert not PR.verify_proof(eng, vk, got, [z[1], (z[2] + 1) % F.p])
    assert got == O.groth16_create_proof(key, ni, z, O.witness_map(F, a, b, c, 0, 0, 0), r, s)


def check_msm_accumulation_cases(ctx, group):
    """the exceptional cases of the bucket accumulation, whichever form the environment forces
    (G753_MSM_AFFINE is read when a context is created): duplicates (doubling), P and -P in one bucket
    (cancellation, then infinity as an operand one level up the tree), an infinity base, repeated small
    scalars (long runs in one bucket: several rounds, batches that span buckets), buckets of one entry"""
    C = GROUPS[group]
    cx = G.Context(0, library=ctx.lib)
    n = 60
    pts = sample_points(C, n, 0x2A0 + group)
    sc = sample_scalars(C, n, 0x2B0 + group)
    pts[1] = None
    pts[6] = pts[5]
    sc[6] = sc[5]                        # same point twice in the same buckets: doubling branch
    pts[8] = C.neg(pts[7])
    sc[8] = sc[7]                        # P and -P with the same scalar: cancellation
    pts[10] = pts[9]
    sc[10] = C.r - sc[9]                 # P with s and -s: cancels through the sign bit
    for i in range(20, 50):
        sc[i] = 7                        # 30 points in one bucket of window 0
    pts[30] = pts[29]                    # ... two of them equal
    coords, inf = points_to_arrays(C, pts)
    bases = cx.upload_bases(group, coords, inf)
    got = G.VariableBaseMSM.multi_scalar_mul(bases, ints_to_array(sc))
    assert projective_to_point(C, got) == O.msm_naive(C, pts, sc)
    # all points equal with one scalar: every bucket step is a doubling or a plain chain
    pts2 = [pts[3]] * 40
    sc2 = [5] * 40
    coords, inf = points_to_arrays(C, pts2)
    b2 = cx.upload_bases(group, coords, inf)
    got = G.VariableBaseMSM.multi_scalar_mul(b2, ints_to_array(sc2))
    assert projective_to_point(C, got) == C.mul(pts[3], 200)
    # infinity as an operand further up the tree: (P - P), (Q + Q'), (A - A), (B - B) pair up as
    # (inf + QQ') and (inf + inf), then QQ' + inf; the same bucket again through the sign bit
    P_, Q_, Q2, A_, B_ = pts[11], pts[12], pts[13], pts[14], pts[15]
    pts3 = [P_, C.neg(P_), Q_, Q2, A_, C.neg(A_), B_, C.neg(B_)]
    for s3 in (3, C.r - 3):
        coords, inf = points_to_arrays(C, pts3)
        b3 = cx.upload_bases(group, coords, inf)
        got = G.VariableBaseMSM.multi_scalar_mul(b3, ints_to_array([s3] * 8))
        assert projective_to_point(C, got) == C.mul(C.add(Q_, Q2), s3)
        b3.free()
    # no bucket with two entries: no round runs, the finish reads the key itself (sign applied)
    for s1 in (1, C.r - 1, 0):
        coords, inf = points_to_arrays(C, [P_])
        b1 = cx.upload_bases(group, coords, inf)
        got = G.VariableBaseMSM.multi_scalar_mul(b1, ints_to_array([s1]))
        assert projective_to_point(C, got) == C.mul(P_, s1)
        b1.free()
    bases.free()
    b2.free()
    cx.close()
