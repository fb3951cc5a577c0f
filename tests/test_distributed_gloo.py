"""N > 1 path on CPU: two gloo ranks, each with its own point-range shard, run the sharded MSM
(local MSM -> all-gather of partial points -> fold) and must agree with the single-rank result and
the oracle.  The compute inside each rank is the TEST-ONLY host-emulation build of the C ABI
(tests/host_emul), standing in for the GPU; the sharding / exchange / fold logic under test is the
product's ginger-lib_b200/distributed.py."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

WORKER = r'''
import importlib, os, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {here!r})
import torch.distributed as dist
from oracle import g753 as O
from util753 import G, GROUPS, ffi, ints_to_array, points_to_arrays, projective_to_point, sample_points, sample_scalars
D = importlib.import_module("ginger-lib_b200.distributed")
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size={world})
rank, world = dist.get_rank(), dist.get_world_size()
lib = ffi.Library(os.path.join({here!r}, "host_emul", "libg753_emul.so"))
ctx = G.Context(0, library=lib)
for group, n in (((ffi.MNT4_G1, 23), (ffi.MNT4_G2, 7)) if world == 2 else ((ffi.MNT4_G1, 23),)):   # (CPU suite time)
    C = GROUPS[group]
    pts = sample_points(C, n, 0x600 + group)          # same seeded inputs on every rank
    sc = sample_scalars(C, n, 0x700 + group)
    pts[2] = None
    sc[4] = 0
    lo, hi = D.shard_range(n, rank, world)
    coords, inf = points_to_arrays(C, pts[lo:hi])
    bases = ctx.upload_bases(group, coords, inf)
    got = D.ShardedMSM(ctx, bases).multi_scalar_mul(ints_to_array(sc[lo:hi]))
    assert projective_to_point(C, got) == O.msm_naive(C, pts, sc), (rank, group)
    bases.free()
assert [D.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
# sharded four-step NTT: scatter -> step1 -> all-to-all -> step2 -> gather == the oracle's transform
from util753 import FIELDS, array_field, field_array
for field, log_n in ((ffi.FIELD_MNT4_FR, 4), (ffi.FIELD_MNT6_FR, 5), (ffi.FIELD_MNT4_FR, 2)):
    if world & (world - 1):
        break                  # the four-step transform shards over a power-of-two number of ranks
    F = FIELDS[field]
    n = 1 << log_n
    rng = O.SplitMix64(0x800 + log_n)
    a = [O.random_field_element(rng, F) for _ in range(n)]
    ref = O.EvaluationDomain(F, n)
    dom = D.ShardedEvaluationDomain(ctx, field, log_n)
    for mode, want in ((ffi.FFT, ref.fft(a)), (ffi.IFFT, ref.ifft(a)), (ffi.COSET_FFT, ref.coset_fft(a)),
                       (ffi.COSET_IFFT, ref.coset_ifft(a))):
        out = dom.gather(dom.transform(dom.scatter(field_array(F, a)), mode))
        assert array_field(F, out) == want, (rank, field, log_n, mode)
    if log_n % 2 == 0:   # n1 == n2: shards chain without re-layout (ifft then coset_fft, as the witness map does)
        sh = dom.transform(dom.transform(dom.scatter(field_array(F, a)), ffi.IFFT), ffi.COSET_FFT)
        assert array_field(F, dom.gather(sh)) == ref.coset_fft(ref.ifft(a)), (rank, "chain")
    dom.close()
# sharded Groth16 prover: heads everywhere, long queries split by point range, partial sums exchanged
import test_groth16_emul as T16
groth16 = importlib.import_module("ginger-lib_b200.groth16")
n, ni, n_aux = 8, 3, 7
key, z, a, b, c = T16.tiny_instance(0x6200, n, ni, n_aux)
F = O.MNT4_FR
r_, s_ = 0x1234567 << 500, F.p - 77
want = O.groth16_create_proof(key, ni, z, O.witness_map(F, a, b, c, 1, 2, 3), r_, s_)
one = lambda C, P: points_to_arrays(C, [P])[0][0]
heads = {{"a": points_to_arrays(O.MNT4_G1, key.a_query[:ni]), "b1": points_to_arrays(O.MNT4_G1, key.b_g1_query[:ni]),
         "b2": points_to_arrays(O.MNT4_G2, key.b_g2_query[:ni]), "h": points_to_arrays(O.MNT4_G1, key.h_query[:ni])}}
shards = {{}}
for name, C, grp, pts in ((("a", O.MNT4_G1, ffi.MNT4_G1, key.a_query[ni:]), ("b1", O.MNT4_G1, ffi.MNT4_G1, key.b_g1_query[ni:]),
                           ("b2", O.MNT4_G2, ffi.MNT4_G2, key.b_g2_query[ni:]), ("h", O.MNT4_G1, ffi.MNT4_G1, key.h_query[ni:]),
                           ("l", O.MNT4_G1, ffi.MNT4_G1, key.l_query)) if world == 2 else ()):
    lo, hi = groth16.shard_range(len(pts), rank, world)
    co, inf = points_to_arrays(C, pts[lo:hi])
    shards[name] = (ctx.upload_bases(grp, co, inf), lo)
if world == 2:     # (three ranks: the placed prover below only - CPU suite time)
    P = groth16.ShardedParameters(ctx, ffi.MNT4_G1, ffi.MNT4_G2, ffi.FIELD_MNT4_FR, one(O.MNT4_G1, key.alpha_g1),
                                  one(O.MNT4_G1, key.beta_g1), one(O.MNT4_G2, key.beta_g2), one(O.MNT4_G1, key.delta_g1),
                                  one(O.MNT4_G2, key.delta_g2), heads, shards, ni)
    proof = groth16.create_proof(P, field_array(F, z), field_array(F, a), field_array(F, b), field_array(F, c), 1, 2, 3, r_, s_)
    got = (T16.affine_of(O.MNT4_G1, proof.a, proof.infinity[0]), T16.affine_of(O.MNT4_G2, proof.b, proof.infinity[1]),
           T16.affine_of(O.MNT4_G1, proof.c, proof.infinity[2]))
    assert got == want, (rank, "sharded groth16")
    P.free()
# the same proof with cost-weighted PLACEMENT: whole MSMs / witness-map chains on different ranks,
# point-to-point exchange of the chains and of h, all-gather of the partial sums
placed = importlib.import_module("ginger-lib_b200.groth16_placed")
totals = {{"a": n_aux, "b1": n_aux, "b2": n_aux, "h": n - 1 - ni, "l": n_aux}}
plan = placed.ProofPlan(world, totals, k2=2, domain=n)
full = {{"a": (O.MNT4_G1, ffi.MNT4_G1, key.a_query[ni:]), "b1": (O.MNT4_G1, ffi.MNT4_G1, key.b_g1_query[ni:]),
        "b2": (O.MNT4_G2, ffi.MNT4_G2, key.b_g2_query[ni:]), "h": (O.MNT4_G1, ffi.MNT4_G1, key.h_query[ni:]),
        "l": (O.MNT4_G1, ffi.MNT4_G1, key.l_query)}}
mine = {{}}
for name, (lo, hi) in plan.shards_of(rank).items():
    C, grp, pts = full[name]
    co, inf = points_to_arrays(C, pts[lo:hi])
    mine[name] = ctx.upload_bases(grp, co, inf)
P = placed.PlacedParameters(ctx, ffi.MNT4_G1, ffi.MNT4_G2, ffi.FIELD_MNT4_FR, one(O.MNT4_G1, key.alpha_g1),
                            one(O.MNT4_G1, key.beta_g1), one(O.MNT4_G2, key.beta_g2), one(O.MNT4_G1, key.delta_g1),
                            one(O.MNT4_G2, key.delta_g2), heads, mine, ni, plan)
for _ in range(2):      # the second proof reuses the placed workspace
    proof = groth16.create_proof(P, field_array(F, z), field_array(F, a), field_array(F, b), field_array(F, c), 1, 2, 3, r_, s_)
    got = (T16.affine_of(O.MNT4_G1, proof.a, proof.infinity[0]), T16.affine_of(O.MNT4_G2, proof.b, proof.infinity[1]),
           T16.affine_of(O.MNT4_G1, proof.c, proof.infinity[2]))
    assert got == want, (rank, "placed groth16", plan.describe())
P.free()
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_proof_plan_tiles_every_msm():
    """the placement gives every point of every MSM to exactly one rank, at most one shard of an MSM per
    rank, and spreads the load: on 8 ranks B2 (Fq2) runs on four of them and A, B1, L on one each, H on one
    or two"""
    placed = importlib.import_module("ginger-lib_b200.groth16_placed")
    for world in (1, 2, 3, 4, 5, 8):
        for n_aux, n_h, k2 in ((1 << 20, (1 << 20) - 4, 2), (7, 4, 2), (1000, 1020, 3), (3, 0, 2)):
            totals = {"a": n_aux, "b1": n_aux, "b2": n_aux, "h": n_h, "l": n_aux}
            plan = placed.ProofPlan(world, totals, k2=k2)
            for nm, parts in plan.shards.items():
                assert parts[0][1] == 0 and parts[-1][2] == totals[nm]
                assert all(parts[i][2] == parts[i + 1][1] for i in range(len(parts) - 1))
                assert len({r for r, _, _ in parts}) == len(parts)
            covered = set()
            for r in range(world):
                covered |= set(plan.shards_of(r))
            assert covered == set(totals)
            assert 0 <= plan.finisher < world and set(plan.chain_owner) == {"a", "b", "c"}
    plan = placed.ProofPlan(8, {"a": 1 << 20, "b1": 1 << 20, "b2": 1 << 20, "h": 1 << 20, "l": 1 << 20}, k2=2)
    assert len(plan.shards["b2"]) == 4 and all(len(plan.shards[nm]) == 1 for nm in ("a", "b1", "l"))
    # H waits for the witness map: its tail may move to the least loaded rank
    assert len(plan.shards["h"]) <= 2 and plan.shards["h"][0][0] == plan.finisher
    assert max(plan.load.values()) < 1.25 * sum(plan.load.values()) / 8


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_msm_ntt_and_provers_gloo(tmp_path, world):
    sys.path.insert(0, HERE)
    from util753 import build_emul
    build_emul()
    script = tmp_path / "worker.py"
    port = 29500 + (os.getpid() % 2000) + world
    script.write_text(WORKER.format(root=ROOT, here=HERE, port=port, world=world))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=900)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, o[-3000:])
        assert "rank %d ok" % r in o
