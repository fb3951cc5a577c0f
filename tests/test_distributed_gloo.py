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
for group, n in ((ffi.MNT4_G1, 23), (ffi.MNT4_G2, 7)):
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
for name, C, grp, pts in (("a", O.MNT4_G1, ffi.MNT4_G1, key.a_query[ni:]), ("b1", O.MNT4_G1, ffi.MNT4_G1, key.b_g1_query[ni:]),
                          ("b2", O.MNT4_G2, ffi.MNT4_G2, key.b_g2_query[ni:]), ("h", O.MNT4_G1, ffi.MNT4_G1, key.h_query[ni:]),
                          ("l", O.MNT4_G1, ffi.MNT4_G1, key.l_query)):
    lo, hi = groth16.shard_range(len(pts), rank, world)
    co, inf = points_to_arrays(C, pts[lo:hi])
    shards[name] = (ctx.upload_bases(grp, co, inf), lo)
P = groth16.ShardedParameters(ctx, ffi.MNT4_G1, ffi.MNT4_G2, ffi.FIELD_MNT4_FR, one(O.MNT4_G1, key.alpha_g1),
                              one(O.MNT4_G1, key.beta_g1), one(O.MNT4_G2, key.beta_g2), one(O.MNT4_G1, key.delta_g1),
                              one(O.MNT4_G2, key.delta_g2), heads, shards, ni)
proof = groth16.create_proof(P, field_array(F, z), field_array(F, a), field_array(F, b), field_array(F, c), 1, 2, 3, r_, s_)
got = (T16.affine_of(O.MNT4_G1, proof.a, proof.infinity[0]), T16.affine_of(O.MNT4_G2, proof.b, proof.infinity[1]),
       T16.affine_of(O.MNT4_G1, proof.c, proof.infinity[2]))
assert got == want, (rank, "sharded groth16")
P.free()
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_sharded_msm_two_ranks_gloo(tmp_path):
    sys.path.insert(0, HERE)
    from util753 import build_emul
    build_emul()
    script = tmp_path / "worker.py"
    port = 29500 + (os.getpid() % 2000)
    script.write_text(WORKER.format(root=ROOT, here=HERE, port=port, world=2))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, o[-3000:])
        assert "rank %d ok" % r in o
