"""Host-side mirror of the Groth16 prover's hot path: `R1CStoQAP::witness_map` from the evaluated
constraints onwards (proof-systems/src/groth16/r1cs_to_qap.rs:121-166) and `create_proof` from the
witness map onwards (proof-systems/src/groth16/prover.rs:241-345).

Same inputs, outputs and semantics as the reference:
  * `Parameters` holds what `groth16::Parameters<E>` holds for the prover (mod.rs:313-371): the five
    queries (resident on the GPU as `Bases`) and the vk points alpha_g1, beta_g1, beta_g2, delta_g1,
    delta_g2;
  * `create_proof(params, full_assignment, a, b, c, d1, d2, d3, r, s)` returns the affine
    proof (A in G1, B in G2, C in G1), every coordinate fully reduced - the bits the reference's
    `Proof { a: g_a.into_affine(), b: g2_b.into_affine(), c: g_c.into_affine() }` holds.
Constraint synthesis / evaluation (prover.rs:215-237, r1cs_to_qap.rs:84-119) is R1CS code outside the
hot path: the caller passes the evaluation vectors a, b, c and the assignment.

All arithmetic runs in libg753.so on the GPU: the witness map is chained on the device
(g753_witness_map_dev), `into_repr` (prover.rs:241-267) is a device pass, the scalars of the nine
MSMs never visit the host, and the fixed small terms (delta * r, query[0], alpha, ...) ride along in
the input-query MSMs.  This file only shapes buffers and sequences C-ABI calls.
"""
import ctypes

import numpy as np

from . import ffi
from .algebra import Bases

LIMBS = 12
ONE = np.array([[1] + [0] * 11], dtype=np.uint64)


def _limbs(v):
    return np.array([[(int(v) >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(LIMBS)]], dtype=np.uint64)


class Parameters:
    """`groth16::Parameters<E>` as the prover reads it.  Queries: (coords (n, 2*k*12) Montgomery
    uint64, infinity (n,) uint8 or None), or an already resident `Bases` (no infinite entries among
    its first num_inputs).  vk points: (2*k*12,) Montgomery affine limbs."""

    def __init__(self, ctx, g1, g2, field, alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2, a_query, b_g1_query,
                 b_g2_query, h_query, l_query, num_inputs, precompute=1):
        self.ctx, self.g1, self.g2, self.field, self.num_inputs = ctx, g1, g2, field, num_inputs
        k2 = ffi.GROUP_K[g2]
        ni = num_inputs

        def up(group, q):
            b = q if isinstance(q, Bases) else Bases(ctx, group, q[0], q[1])
            if precompute is not None and precompute != 1 and len(b) >= 1 << 12:
                b.precompute(precompute)
            return b


        # small resident keys: query[0..ni] followed by the vk points their proof element adds
        # (prover.rs:274-283, 290-299, 306-315), so one short MSM covers all fixed small terms
        def small(group, q, extra, k):
            coords, inf = (q.download(0, ni), None) if isinstance(q, Bases) else q
            c = np.concatenate([ffi.as_u64(coords).reshape(-1, 2 * k * LIMBS)[:ni]] +
                               [ffi.as_u64(e).reshape(1, 2 * k * LIMBS) for e in extra])
            i = np.zeros(c.shape[0], dtype=np.uint8)
            if inf is not None:
                i[:ni] = np.asarray(inf, dtype=np.uint8)[:ni]
            return Bases(ctx, group, c, i)

        self.a_small = small(g1, a_query, [delta_g1, alpha_g1], 1)
        self.b1_small = small(g1, b_g1_query, [delta_g1, beta_g1], 1)
        self.b2_small = small(g2, b_g2_query, [delta_g2, beta_g2], k2)
        self.a_query, self.b_g1_query, self.h_query, self.l_query = (up(g1, q) for q in (a_query, b_g1_query, h_query, l_query))
        self.b_g2_query = up(g2, b_g2_query)
        # the five long MSMs: name -> (bases, first base, offset of the first scalar inside its vector)
        self.aux = {"a": (self.a_query, ni, 0), "b1": (self.b_g1_query, ni, 0), "b2": (self.b_g2_query, ni, 0),
                    "h": (self.h_query, ni, 0), "l": (self.l_query, 0, 0)}
        self.h_head = self.h_query          # h_query[0..ni] (prover.rs:320)
        self.sharder = None
        for b in (self.a_small, self.b1_small, self.b2_small):
            b.precompute(64)        # fixed tiny keys: ~12 doublings left in the serial window fold
        self.delta_g1 = ffi.as_u64(delta_g1).reshape(1, 2 * LIMBS)
        # second context of the same device for the per-proof latency chain, with its pre-allocated
        # 3-point key [g_a, g1_b, delta_g1] (the first two are overwritten every proof)
        from .algebra import Context
        self.ctx2 = Context(ctx.device, library=ctx.lib)
        self.fresh = Bases(self.ctx2, g1, np.concatenate([self.delta_g1] * 3), np.zeros(3, dtype=np.uint8))

    def free(self):
        for b in (self.a_query, self.b_g1_query, self.b_g2_query, self.h_query, self.l_query, self.a_small,
                  self.b1_small, self.b2_small, self.fresh):
            b.free()
        self._free_ws()
        self.ctx2.close()

    def _free_ws(self):
        ws = getattr(self, "_ws", None)
        if ws is not None:
            for dv in ws["bufs"]:
                dv.free()
            self._ws = None


def shard_range(n, rank, world):
    """contiguous range of `rank` among `world` (sizes differ by at most one)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedParameters(Parameters):
    """A proving key sharded over the ranks of a process group: every rank keeps the
    short head of each query (query[0..num_inputs] and the vk points) and one contiguous point range
    of each long query.  All ranks run the witness map on the same inputs, the five long MSMs run on
    the local ranges, their partial sums are all-gathered (N x 288 / 576 bytes over NCCL) and folded
    on every rank, and every rank finishes the same proof.

    heads: {"a", "b1", "b2", "h"} -> (coords of the first num_inputs bases, infinity or None);
    shards: {"a", "b1", "b2", "h", "l"} -> (Bases over this rank's range, lo), lo = index of the range's
    first base counted from the start of the long part (query[num_inputs + lo], l_query[lo])."""

    def __init__(self, ctx, g1, g2, field, alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2, heads, shards, num_inputs,
                 process_group=None, precompute=1):
        import torch.distributed as dist
        self.ctx, self.g1, self.g2, self.field, self.num_inputs = ctx, g1, g2, field, num_inputs
        k2 = ffi.GROUP_K[g2]
        ni = num_inputs

        def small(group, q, extra, k):
            coords, inf = q
            c = np.concatenate([ffi.as_u64(coords).reshape(-1, 2 * k * LIMBS)[:ni]] +
                               [ffi.as_u64(e).reshape(1, 2 * k * LIMBS) for e in extra])
            i = np.zeros(c.shape[0], dtype=np.uint8)
            if inf is not None:
                i[:ni] = np.asarray(inf, dtype=np.uint8)[:ni]
            return Bases(ctx, group, c, i)

        self.a_small = small(g1, heads["a"], [delta_g1, alpha_g1], 1)
        self.b1_small = small(g1, heads["b1"], [delta_g1, beta_g1], 1)
        self.b2_small = small(g2, heads["b2"], [delta_g2, beta_g2], k2)
        self.h_head = small(g1, heads["h"], [], 1)
        for b in (self.a_small, self.b1_small, self.b2_small):
            b.precompute(64)
        self.aux = {}
        for name, (bases, lo) in shards.items():
            if precompute is not None and precompute != 1 and len(bases) >= 1 << 12:
                bases.precompute(precompute)
            self.aux[name] = (bases, 0, lo)
        self.delta_g1 = ffi.as_u64(delta_g1).reshape(1, 2 * LIMBS)
        from .algebra import Context
        self.ctx2 = Context(ctx.device, library=ctx.lib)
        self.fresh = Bases(self.ctx2, g1, np.concatenate([self.delta_g1] * 3), np.zeros(3, dtype=np.uint8))
        self.sharder = _Sharder(ctx, dist, process_group)

    def free(self):
        for b in [v[0] for v in self.aux.values()] + [self.a_small, self.b1_small, self.b2_small, self.h_head, self.fresh]:
            b.free()
        self._free_ws()
        self.ctx2.close()


class _Sharder:
    """all-gather + fold of per-rank partial points, in place, on the context's stream"""

    def __init__(self, ctx, dist, pg):
        self.ctx, self.dist, self.pg = ctx, dist, pg
        self.world = dist.get_world_size(pg) if dist.is_initialized() else 1
        self.backend = dist.get_backend(pg) if dist.is_initialized() else None

    def reduce(self, group, d_slots):
        if self.world == 1:
            return
        import torch
        lib, ctx = self.ctx.lib, self.ctx
        k = ffi.GROUP_K[group]
        nb = 3 * k * 96
        for d in d_slots:
            if self.backend == "nccl":
                dev = torch.device("cuda", ctx.device)
                mine = torch.empty(nb // 8, dtype=torch.int64, device=dev)
                allp = torch.empty(self.world * nb // 8, dtype=torch.int64, device=dev)
                lib.check(lib.d2d(ctx.handle, ctypes.c_void_p(mine.data_ptr()), d, nb))
                lib.check(lib.sync(ctx.handle))
                self.dist.all_gather_into_tensor(allp, mine, group=self.pg)
                torch.cuda.synchronize(dev)
                lib.check(lib.points_sum_dev(ctx.handle, group, ctypes.c_void_p(allp.data_ptr()), self.world, d))
                lib.check(lib.sync(ctx.handle))
            else:
                host = np.empty(nb // 8, dtype=np.uint64)
                lib.check(lib.d2h(ctx.handle, ffi.ptr(host), d, nb))
                mine = torch.from_numpy(host.view(np.int64))
                parts = [torch.empty_like(mine) for _ in range(self.world)]
                self.dist.all_gather(parts, mine, group=self.pg)
                allp = np.ascontiguousarray(np.concatenate([p.numpy().view(np.uint64) for p in parts]))
                tmp = ctypes.c_void_p()
                lib.check(lib.dev_alloc(ctx.handle, allp.nbytes, ctypes.byref(tmp)))
                lib.check(lib.h2d(ctx.handle, tmp, ffi.ptr(allp), allp.nbytes))
                lib.check(lib.points_sum_dev(ctx.handle, group, tmp, self.world, d))
                lib.check(lib.sync(ctx.handle))
                lib.dev_free(ctx.handle, tmp)


class Proof:
    def __init__(self, a, b, c, inf):
        self.a, self.b, self.c, self.infinity = a, b, c, inf


class _Dev:
    """a device buffer of `n` field elements / scalars"""

    def __init__(self, ctx, n):
        self.ctx, self.n = ctx, n
        self.p = ctypes.c_void_p()
        ctx.lib.check(ctx.lib.dev_alloc(ctx.handle, max(n, 1) * 96, ctypes.byref(self.p)))

    def at(self, i):
        return ctypes.c_void_p(self.p.value + 96 * i)

    def put(self, host, first=0):
        host = ffi.as_u64(host).reshape(-1, LIMBS)
        self.ctx.lib.check(self.ctx.lib.h2d(self.ctx.handle, self.at(first), ffi.ptr(host), host.shape[0] * 96))

    def get(self, first=0, count=None):
        count = self.n - first if count is None else count
        out = np.empty((count, LIMBS), dtype=np.uint64)
        self.ctx.lib.check(self.ctx.lib.d2h(self.ctx.handle, ffi.ptr(out), self.at(first), count * 96))
        return out

    def free(self):
        if self.p:
            self.ctx.lib.dev_free(self.ctx.handle, self.p)
            self.p = None


def witness_map(ctx, field, a, b, c, d1, d2, d3):
    """host-buffer form: a, b, c (n, 12) Montgomery evaluations, d1..d3 canonical ints -> h (n+1, 12)
    Montgomery (r1cs_to_qap.rs:121-166)"""
    lib = ctx.lib
    a, b, c = (ffi.as_u64(v).reshape(-1, LIMBS) for v in (a, b, c))
    n = a.shape[0]
    log_n = n.bit_length() - 1
    if 1 << log_n != n or b.shape[0] != n or c.shape[0] != n:
        raise ValueError("a, b, c must have the domain's size (a power of two)")
    d = np.concatenate([_limbs(d1), _limbs(d2), _limbs(d3)])
    dm = np.zeros_like(d)
    lib.check(lib.field_op(ctx.handle, field, ffi.OP_TO_MONT, ffi.ptr(d), None, ffi.ptr(dm), 3))
    h = np.zeros((n + 1, LIMBS), dtype=np.uint64)
    lib.check(lib.witness_map(ctx.handle, field, ffi.ptr(a), ffi.ptr(b), ffi.ptr(c), log_n, ffi.ptr(dm), ffi.ptr(h)))
    return h


def create_proof(params, full_assignment, a, b, c, d1, d2, d3, r, s, timings=None, profile=None):
    """prover.rs:201-345 after constraint synthesis.

    full_assignment: (num_inputs + num_aux, 12) Montgomery, inputs first, index 0 = the constant one;
    a, b, c: (domain_size, 12) Montgomery evaluation vectors; d1, d2, d3, r, s: canonical ints (E::Fr).
    Returns Proof with affine Montgomery limbs: a (2, 12), b (2, k2*12), c (2, 12).

    Scheduling: the large (throughput-bound) MSMs are queued on the main context's stream; the one MSM
    over per-proof bases (s*g_a + r*g1_b - rs*delta, a 753-doubling latency chain) runs on a second
    context of the same device as soon as g_a and g1_b exist, overlapping the remaining large MSMs.

    profile: optional dict; when given, the stream is drained after each of the five long MSMs and its
    device phase times (digits / sort / accumulate / reduce / combine, ms) and plan are stored under the
    MSM's name - a diagnostic pass (the extra synchronisations cost overlap), never the timed one."""
    import time
    if getattr(params, "plan", None) is not None:      # cost-weighted placement over several GPUs
        from .groth16_placed import create_proof_placed
        return create_proof_placed(params, full_assignment, a, b, c, d1, d2, d3, r, s, timings=timings, profile=profile)
    ctx, ctx2, lib, field = params.ctx, params.ctx2, params.ctx.lib, params.field
    g1, g2, ni = params.g1, params.g2, params.num_inputs
    k2 = ffi.GROUP_K[g2]
    t0 = time.perf_counter()
    a, b, c = (ffi.as_u64(v).reshape(-1, LIMBS) for v in (a, b, c))
    z = ffi.as_u64(full_assignment).reshape(-1, LIMBS)
    n = a.shape[0]
    log_n = n.bit_length() - 1
    n_vars = z.shape[0]
    n_aux = n_vars - ni

    # every device buffer of the proof is allocated before anything is queued and kept with the key
    # for the next proof of the same shape (cudaMalloc / cudaFree synchronise the device)
    ws = getattr(params, "_ws", None)
    if ws is None or ws["shape"] != (n, n_vars, ni):
        if ws is not None:
            for dv in ws["bufs"]:
                dv.free()
        bufs = [_Dev(ctx, n) for _ in range(3)] + [_Dev(ctx, n + 1), _Dev(ctx, n_vars + 1), _Dev(ctx, ni + 2),
                                                    _Dev(ctx, ni + 2), _Dev(ctx, 3), _Dev(ctx, 3 * 16), _Dev(ctx, 3 * k2 * 4)]
        ws = params._ws = {"shape": (n, n_vars, ni), "bufs": bufs}
    bufs = ws["bufs"]
    dev, d_h, d_z, sr, ss, sc3, out1, out2 = bufs[:3], bufs[3], bufs[4], bufs[5], bufs[6], bufs[7], bufs[8], bufs[9]
    slot = lambda o, i, k=1: o.at(3 * k * i)
    try:
        def msm(cx, bases, first, count, d_scalars, d_out):
            count = max(0, min(count, len(bases) - first))
            lib.check(lib.msm_dev(cx.handle, bases.handle, first, count, d_scalars, d_out))

        def aux_msm(name, d_vec, base_index, total, d_out, cx=ctx):
            """one of the five long MSMs (or this rank's shard of it, see ShardedParameters)"""
            bases, first, off = params.aux[name]
            msm(cx, bases, first, max(0, total - off), d_vec.at(base_index + off), d_out)
            if profile is not None:
                cx.sync()
                profile[name] = dict(cx.last_msm_phases(), points=max(0, min(total - off, len(bases) - first)),
                                     plan=cx.last_msm_plan())

        sharder = params.sharder
        # one GPU: the G2 MSM (the longest, and independent of the witness map) runs on the SECOND context from
        # the moment the assignment is on the device, beside the witness map and the four G1 MSMs - the blocks of
        # its accumulation fill the SMs the other stream's low-occupancy phases (sort, reduction, fold) leave idle
        concurrent = sharder is None
        # ---- into_repr of the assignment (prover.rs:241-267) and the short scalar vectors ----------------
        dd = np.concatenate([_limbs(d1), _limbs(d2), _limbs(d3), _limbs(r), _limbs(s)])
        dm = np.zeros_like(dd)
        lib.check(lib.field_op(ctx.handle, field, ffi.OP_TO_MONT, ffi.ptr(dd), None, ffi.ptr(dm), 5))
        # -rs for the C term (prover.rs:327): field arithmetic on the device, canonical result
        rs = np.zeros((1, LIMBS), dtype=np.uint64)
        lib.check(lib.field_op(ctx.handle, field, ffi.OP_MUL, ffi.ptr(dm[3:4]), ffi.ptr(dm[4:5]), ffi.ptr(rs), 1))
        zero = np.zeros((1, LIMBS), dtype=np.uint64)
        lib.check(lib.field_op(ctx.handle, field, ffi.OP_SUB, ffi.ptr(zero), ffi.ptr(rs), ffi.ptr(rs), 1))
        lib.check(lib.field_op(ctx.handle, field, ffi.OP_FROM_MONT, ffi.ptr(rs), None, ffi.ptr(rs), 1))
        sc3.put(np.concatenate([_limbs(s), _limbs(r), rs]))
        d_z.put(z)
        lib.check(lib.vec_op_dev(ctx.handle, field, ffi.OP_FROM_MONT, d_z.p, None, n_vars))
        # small scalar vectors [1, inputs..., r|s, 1]: the constant-one slot of the assignment is
        # replaced by a literal 1 (the reference adds query[0] unconditionally, prover.rs:277)
        for dv, blind in ((sr, r), (ss, s)):
            dv.put(ONE, 0)
            if ni > 1:
                lib.check(lib.d2d(ctx.handle, dv.at(1), d_z.at(1), (ni - 1) * 96))
            dv.put(_limbs(blind), ni)
            dv.put(ONE, ni + 1)
        if concurrent:
            # B in G2 (:302-315) on the second context
            lib.check(lib.ctx_wait(ctx2.handle, ctx.handle))
            msm(ctx2, params.b2_small, 0, ni + 2, ss.p, slot(out2, 0, k2))
            aux_msm("b2", d_z, ni, n_aux, slot(out2, 1, k2), cx=ctx2)
            lib.check(lib.points_sum_dev(ctx2.handle, g2, slot(out2, 0, k2), 2, slot(out2, 2, k2)))
        # ---- witness map on the device, then into_repr of h ---------------------------------------------
        for dv, host in zip(dev, (a, b, c)):
            dv.put(host)
        lib.check(lib.witness_map_dev(ctx.handle, field, dev[0].p, dev[1].p, dev[2].p, log_n, ffi.ptr(dm), d_h.p))
        lib.check(lib.vec_op_dev(ctx.handle, field, ffi.OP_FROM_MONT, d_h.p, None, n + 1))
        if timings is not None:
            ctx.sync()
            timings["witness_map"] = time.perf_counter() - t0
            t0 = time.perf_counter()

        # A (prover.rs:270-283) -> slots 0, 1; B in G1 (:286-299) -> slots 2, 3; g_a, g1_b -> slots 8, 9
        aux_msm("a", d_z, ni, n_aux, slot(out1, 1))
        aux_msm("b1", d_z, ni, n_aux, slot(out1, 3))
        if sharder is not None:
            sharder.reduce(g1, [slot(out1, 1), slot(out1, 3)])
        msm(ctx, params.a_small, 0, ni + 2, sr.p, slot(out1, 0))
        msm(ctx, params.b1_small, 0, ni + 2, ss.p, slot(out1, 2))
        lib.check(lib.points_sum_dev(ctx.handle, g1, slot(out1, 0), 2, slot(out1, 8)))
        lib.check(lib.points_sum_dev(ctx.handle, g1, slot(out1, 2), 2, slot(out1, 9)))
        if not concurrent:
            lib.check(lib.ctx_wait(ctx2.handle, ctx.handle))
        # C (:318-337): L, H (zip-truncated against h: n + 1 scalars vs n - 1 bases) -> slots 4..6
        aux_msm("l", d_z, ni, n_aux, slot(out1, 6))
        aux_msm("h", d_h, ni, n + 1 - ni, slot(out1, 5))
        msm(ctx, params.h_head, 0, min(ni, n + 1), d_h.at(0), slot(out1, 4))
        if not concurrent:
            # B in G2 (:302-315)
            msm(ctx, params.b2_small, 0, ni + 2, ss.p, slot(out2, 0, k2))
            aux_msm("b2", d_z, ni, n_aux, slot(out2, 1, k2))
            sharder.reduce(g1, [slot(out1, 6), slot(out1, 5)])
            sharder.reduce(g2, [slot(out2, 1, k2)])
            lib.check(lib.points_sum_dev(ctx.handle, g2, slot(out2, 0, k2), 2, slot(out2, 2, k2)))

        # second context: s * g_a + r * g1_b - rs * delta_g1 (:322-329) as one MSM over fresh bases
        # (sharded keys: on the second context, overlapping the remaining large MSMs; one GPU: on the first,
        # the second is busy with the G2 MSM)
        cf = ctx if concurrent else ctx2
        ga_b1 = np.empty((2, 3 * LIMBS), dtype=np.uint64)
        lib.check(lib.d2h(cf.handle, ffi.ptr(ga_b1), slot(out1, 8), 2 * 3 * 96))
        xy = np.zeros((2, 2 * LIMBS), dtype=np.uint64)
        inf = np.zeros(2, dtype=np.uint8)
        lib.check(lib.batch_normalize(cf.handle, g1, ffi.ptr(ga_b1), 2, ffi.ptr(xy), ffi.ptr(inf)))
        lib.check(lib.bases_update(cf.handle, params.fresh.handle, 0, 2, ffi.ptr(xy), ffi.ptr(inf)))
        msm(cf, params.fresh, 0, 3, sc3.p, slot(out1, 7))
        lib.check(lib.ctx_wait(ctx.handle, ctx2.handle))
        lib.check(lib.points_sum_dev(ctx.handle, g1, slot(out1, 4), 4, slot(out1, 10)))
        gc = out1.get(3 * 10, 3).reshape(1, 3 * LIMBS)
        gb2 = out2.get(3 * k2 * 2, 3 * k2).reshape(1, 3 * k2 * LIMBS)
        xyc = np.zeros((1, 2 * LIMBS), dtype=np.uint64)
        infc = np.zeros(1, dtype=np.uint8)
        lib.check(lib.batch_normalize(ctx.handle, g1, ffi.ptr(gc), 1, ffi.ptr(xyc), ffi.ptr(infc)))
        xyb = np.zeros((1, 2 * k2 * LIMBS), dtype=np.uint64)
        infb = np.zeros(1, dtype=np.uint8)
        lib.check(lib.batch_normalize(ctx.handle, g2, ffi.ptr(gb2), 1, ffi.ptr(xyb), ffi.ptr(infb)))
        if timings is not None:
            ctx.sync()
            timings["msm_and_assembly"] = time.perf_counter() - t0
        return Proof(xy[0].reshape(2, LIMBS), xyb[0].reshape(2, k2 * LIMBS), xyc[0].reshape(2, LIMBS),
                     (bool(inf[0]), bool(infb[0]), bool(infc[0])))
    finally:
        ctx2.sync()
