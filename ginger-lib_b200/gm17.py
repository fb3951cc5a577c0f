"""Host-side mirror of the GM17 prover's hot path: `R1CStoSAP::witness_map` from the evaluated
constraints onwards (proof-systems/src/gm17/r1cs_to_sap.rs:191-245) and `create_proof` from the witness
map onwards (proof-systems/src/gm17/prover.rs:235-354) - SURVEY.md 8f-4.  No new kernels: the same
NTT, element-wise and MSM entry points of libg753.so as the Groth16 mirror (groth16.py).

  * `Parameters` holds what `gm17::Parameters<E>` holds for the prover (gm17/mod.rs:138-149): the five
    queries (resident on the GPU as `Bases`) and the points g_gamma_z, h_gamma_z, g_ab_gamma_z,
    g_gamma2_z2;
  * `create_proof(params, full_assignment, a, c, d1, d2, r)` returns the affine proof (A in G1, B in
    G2, C in G1), every coordinate fully reduced - the bits of the reference's
    `Proof { a: g_a.into_affine(), b: g_b.into_affine(), c: g_c.into_affine() }`.
Constraint synthesis and evaluation (prover.rs:205-230, r1cs_to_sap.rs:121-189, 206-232) is R1CS code
outside the hot path: the caller passes the SAP assignment (inputs, aux and the extra variables of
r1cs_to_sap.rs:127-151) and the evaluation vectors a and c of the domain's size.

The reference multiplies the point c2_acc by r (prover.rs:333-334); here the scalars of that MSM are
multiplied by r on the device instead (r * sum s_i P_i = sum (r s_i) P_i): the same group element, so
the same affine proof, without a per-proof key.  The fixed small terms (r * g_gamma_z, query[0], ...)
ride along in the short input-query MSMs, as in groth16.py.
"""
import ctypes

import numpy as np

from . import ffi
from .algebra import Bases
from .groth16 import LIMBS, ONE, Proof, _Dev, _limbs


class Parameters:
    """`gm17::Parameters<E>` as the prover reads it.  Queries: (coords (n, 2*k*12) Montgomery uint64,
    infinity (n,) uint8 or None); single points: (2*k*12,) Montgomery affine limbs."""

    def __init__(self, ctx, g1, g2, field, a_query, b_query, c_query_1, c_query_2, g_gamma_z, h_gamma_z,
                 g_ab_gamma_z, g_gamma2_z2, g_gamma2_z_t, num_inputs, precompute=1):
        self.ctx, self.g1, self.g2, self.field, self.num_inputs = ctx, g1, g2, field, num_inputs
        k2 = ffi.GROUP_K[g2]
        ni = num_inputs

        def up(group, q):
            b = Bases(ctx, group, q[0], q[1])
            if precompute is not None and precompute != 1 and len(b) >= 1 << 12:
                b.precompute(precompute)
            return b

        def small(group, q, pre, post, k):
            """pre + query[1..ni] + post as one short resident key; an entry of pre / post is either
            affine limbs or (limbs, infinity flag)"""
            coords = ffi.as_u64(q[0]).reshape(-1, 2 * k * LIMBS)
            split = lambda e: e if isinstance(e, tuple) else (e, False)
            rows = [ffi.as_u64(split(e)[0]).reshape(1, 2 * k * LIMBS) for e in pre] + [coords[1:ni]] + \
                   [ffi.as_u64(split(e)[0]).reshape(1, 2 * k * LIMBS) for e in post]
            c = np.concatenate(rows)
            i = np.zeros(c.shape[0], dtype=np.uint8)
            i[:len(pre)] = [split(e)[1] for e in pre]
            if post:
                i[-len(post):] = [split(e)[1] for e in post]
            if q[1] is not None:
                i[len(pre):len(pre) + ni - 1] = np.asarray(q[1], dtype=np.uint8)[1:ni]
            return Bases(ctx, group, c, i)

        def head(q, k):
            """query[0] with its infinity flag: the reference adds it whatever it is (prover.rs:268, 286,
            333-336), and an infinite base contributes nothing to the short MSM it rides in"""
            coords = ffi.as_u64(q[0]).reshape(-1, 2 * k * LIMBS)
            return (coords[0], bool(q[1] is not None and np.asarray(q[1])[0]))

        # A: [g_gamma_z, g_gamma_z, a_query[0], a_query[1..ni]] . [r, d1, 1, inputs]  (prover.rs:264-278)
        self.a_small = small(g1, a_query, [g_gamma_z, g_gamma_z, head(a_query, 1)], [], 1)
        # B: the same over G2 with h_gamma_z (:280-296)
        self.b_small = small(g2, b_query, [h_gamma_z, h_gamma_z, head(b_query, k2)], [], k2)
        # C: [g_gamma2_z2, g_ab_gamma_z, g_ab_gamma_z, c_query_2[0], g_gamma2_z2, g_gamma2_z_t[0], c_query_2[1..ni]]
        #    . [r^2, r, d1, r, 2 d1 r, d2, r * inputs]  (:327-343)
        self.c_small = small(g1, c_query_2, [g_gamma2_z2, g_ab_gamma_z, g_ab_gamma_z, head(c_query_2, 1), g_gamma2_z2,
                                             head(g_gamma2_z_t, 1)], [], 1)
        self.a_query, self.c_query_1, self.c_query_2, self.g_gamma2_z_t = (
            up(g1, q) for q in (a_query, c_query_1, c_query_2, g_gamma2_z_t))
        self.b_query = up(g2, b_query)

    def free(self):
        for b in (self.a_small, self.b_small, self.c_small, self.a_query, self.b_query, self.c_query_1,
                  self.c_query_2, self.g_gamma2_z_t):
            b.free()


def _to_mont(ctx, field, vals):
    """canonical ints -> (len, 12) Montgomery limbs, converted on the device"""
    d = np.concatenate([_limbs(v) for v in vals])
    out = np.zeros_like(d)
    ctx.lib.check(ctx.lib.field_op(ctx.handle, field, ffi.OP_TO_MONT, ffi.ptr(d), None, ffi.ptr(out), d.shape[0]))
    return out


def _fop(ctx, field, op, a, b=None):
    out = np.zeros((1, LIMBS), dtype=np.uint64)
    ctx.lib.check(ctx.lib.field_op(ctx.handle, field, op, ffi.ptr(np.ascontiguousarray(a)),
                                   ffi.ptr(np.ascontiguousarray(b)) if b is not None else None, ffi.ptr(out), 1))
    return out


def _witness_map_dev(ctx, field, d_a, d_c, d_h, n, d1m, d2m):
    """r1cs_to_sap.rs:191-245 chained on the device: d_a, d_c hold the evaluation vectors (both are
    overwritten), d_h receives the n + 1 Montgomery coefficients of h"""
    lib = ctx.lib
    log_n = n.bit_length() - 1
    ntt = lambda d, mode: lib.check(lib.ntt_dev(ctx.handle, field, d.p, log_n, mode))
    ntt(d_a, ffi.IFFT)                                                              # :191
    d1_double = _fop(ctx, field, ffi.OP_ADD, d1m, d1m)                              # :193
    lib.check(lib.d2d(ctx.handle, d_h.p, d_a.p, n * 96))                            # :194-195  h_i = 2 d1 * a_i
    lib.check(lib.vec_scale_dev(ctx.handle, field, d_h.p, ffi.ptr(d1_double), n))
    d1d1 = _fop(ctx, field, ffi.OP_MUL, d1m, d1m)
    h0 = _fop(ctx, field, ffi.OP_SUB, _fop(ctx, field, ffi.OP_SUB, d_h.get(0, 1), d2m), d1d1)   # :196-198
    d_h.put(h0, 0)
    d_h.put(d1d1, n)                                                                # :199
    ntt(d_a, ffi.COSET_FFT)                                                         # :201
    lib.check(lib.vec_op_dev(ctx.handle, field, ffi.OP_SQR, d_a.p, None, n))        # :203  aa = a . a
    ntt(d_c, ffi.IFFT)                                                              # :234
    ntt(d_c, ffi.COSET_FFT)                                                         # :235
    lib.check(lib.vec_op_dev(ctx.handle, field, ffi.OP_SUB, d_a.p, d_c.p, n))       # :237
    zinv = np.zeros((1, LIMBS), dtype=np.uint64)
    lib.check(lib.domain_constant(ctx.handle, field, log_n, 4, ffi.ptr(zinv)))
    lib.check(lib.vec_scale_dev(ctx.handle, field, d_a.p, ffi.ptr(zinv), n))        # :239
    ntt(d_a, ffi.COSET_IFFT)                                                        # :240
    if n > 1:
        lib.check(lib.vec_op_dev(ctx.handle, field, ffi.OP_ADD, d_h.p, d_a.p, n - 1))   # :242-245


def witness_map(ctx, field, a, c, d1, d2):
    """host-buffer form: a, c (n, 12) Montgomery evaluations, d1, d2 canonical ints -> h (n + 1, 12)
    Montgomery"""
    a, c = (ffi.as_u64(v).reshape(-1, LIMBS) for v in (a, c))
    n = a.shape[0]
    if 1 << (n.bit_length() - 1) != n or c.shape[0] != n:
        raise ValueError("a, c must have the domain's size (a power of two)")
    d_a, d_c, d_h = _Dev(ctx, n), _Dev(ctx, n), _Dev(ctx, n + 1)
    try:
        d_a.put(a)
        d_c.put(c)
        dm = _to_mont(ctx, field, [d1, d2])
        _witness_map_dev(ctx, field, d_a, d_c, d_h, n, dm[0:1], dm[1:2])
        return d_h.get()
    finally:
        for d in (d_a, d_c, d_h):
            d.free()


def create_proof(params, full_assignment, a, c, d1, d2, r):
    """gm17/prover.rs:198-354 after constraint synthesis.

    full_assignment: (sap_num_variables, 12) Montgomery, inputs first, index 0 = the constant one;
    a, c: (domain_size, 12) Montgomery evaluation vectors; d1, d2, r: canonical ints (E::Fr).
    Returns Proof with affine Montgomery limbs: a (2, 12), b (2, k2*12), c (2, 12)."""
    ctx, lib, field = params.ctx, params.ctx.lib, params.field
    g1, g2, ni = params.g1, params.g2, params.num_inputs
    k2 = ffi.GROUP_K[g2]
    a, c = (ffi.as_u64(v).reshape(-1, LIMBS) for v in (a, c))
    z = ffi.as_u64(full_assignment).reshape(-1, LIMBS)
    n, n_vars = a.shape[0], z.shape[0]
    n_aux = n_vars - ni
    d_a, d_c, d_h = _Dev(ctx, n), _Dev(ctx, n), _Dev(ctx, n + 1)
    d_z, d_rz = _Dev(ctx, n_vars), _Dev(ctx, n_vars)
    s_ab, s_c = _Dev(ctx, ni + 2), _Dev(ctx, ni + 5)
    out1, out2 = _Dev(ctx, 3 * 8), _Dev(ctx, 3 * k2 * 4)
    slot = lambda o, i, k=1: o.at(3 * k * i)
    bufs = (d_a, d_c, d_h, d_z, d_rz, s_ab, s_c, out1, out2)
    try:
        d_a.put(a)
        d_c.put(c)
        dm = _to_mont(ctx, field, [d1, d2, r])
        d1m, d2m, rm = dm[0:1], dm[1:2], dm[2:3]
        _witness_map_dev(ctx, field, d_a, d_c, d_h, n, d1m, d2m)                    # prover.rs:232
        # into_repr of h and of the assignment (:235-261); r * assignment for the c_query_2 terms
        lib.check(lib.vec_op_dev(ctx.handle, field, ffi.OP_FROM_MONT, d_h.p, None, n + 1))
        d_z.put(z)
        lib.check(lib.d2d(ctx.handle, d_rz.p, d_z.p, n_vars * 96))
        lib.check(lib.vec_scale_dev(ctx.handle, field, d_rz.p, ffi.ptr(rm), n_vars))
        lib.check(lib.vec_op_dev(ctx.handle, field, ffi.OP_FROM_MONT, d_z.p, None, n_vars))
        lib.check(lib.vec_op_dev(ctx.handle, field, ffi.OP_FROM_MONT, d_rz.p, None, n_vars))
        # r^2 and 2 d1 r (:299-301): field products on the device, canonical for the MSM
        canon = lambda m: _fop(ctx, field, ffi.OP_FROM_MONT, m)
        r2 = canon(_fop(ctx, field, ffi.OP_MUL, rm, rm))
        d1_r_2 = canon(_fop(ctx, field, ffi.OP_MUL, d1m, _fop(ctx, field, ffi.OP_ADD, rm, rm)))
        # short scalar vectors: [r, d1, 1, inputs] and [r^2, r, d1, r, 2 d1 r, d2, r * inputs]
        s_ab.put(np.concatenate([_limbs(r), _limbs(d1), ONE]), 0)
        s_c.put(np.concatenate([r2, _limbs(r), _limbs(d1), _limbs(r), d1_r_2, _limbs(d2)]), 0)
        if ni > 1:
            lib.check(lib.d2d(ctx.handle, s_ab.at(3), d_z.at(1), (ni - 1) * 96))
            lib.check(lib.d2d(ctx.handle, s_c.at(6), d_rz.at(1), (ni - 1) * 96))

        def msm(bases, first, count, d_scalars, d_out):
            count = max(0, min(count, len(bases) - first))                # zip-truncation, variable_base.rs:36
            lib.check(lib.msm_dev(ctx.handle, bases.handle, first, count, d_scalars, d_out))

        # A (:264-278) -> out1 slots 0, 1
        msm(params.a_small, 0, ni + 2, s_ab.p, slot(out1, 0))
        msm(params.a_query, ni, n_aux, d_z.at(ni), slot(out1, 1))
        lib.check(lib.points_sum_dev(ctx.handle, g1, slot(out1, 0), 2, slot(out1, 6)))
        # B (:280-296) -> out2 slots 0, 1
        msm(params.b_small, 0, ni + 2, s_ab.p, slot(out2, 0, k2))
        msm(params.b_query, ni, n_aux, d_z.at(ni), slot(out2, 1, k2))
        lib.check(lib.points_sum_dev(ctx.handle, g2, slot(out2, 0, k2), 2, slot(out2, 2, k2)))
        # C (:298-344): c1, the fixed terms with r * c2_inputs, r * c2_aux, g_acc -> out1 slots 2..5
        msm(params.c_query_1, 0, n_aux, d_z.at(ni), slot(out1, 2))        # get_c_query_1(0): the whole query
        msm(params.c_small, 0, ni + 5, s_c.p, slot(out1, 3))
        msm(params.c_query_2, ni, n_aux, d_rz.at(ni), slot(out1, 4))
        msm(params.g_gamma2_z_t, 0, n + 1, d_h.p, slot(out1, 5))          # h_input and h_aux parts in one
        lib.check(lib.points_sum_dev(ctx.handle, g1, slot(out1, 2), 4, slot(out1, 7)))
        ga = out1.get(3 * 6, 3).reshape(1, 3 * LIMBS)
        gc = out1.get(3 * 7, 3).reshape(1, 3 * LIMBS)
        gb = out2.get(3 * k2 * 2, 3 * k2).reshape(1, 3 * k2 * LIMBS)
        xy1 = np.zeros((2, 2 * LIMBS), dtype=np.uint64)
        inf1 = np.zeros(2, dtype=np.uint8)
        lib.check(lib.batch_normalize(ctx.handle, g1, ffi.ptr(np.concatenate([ga, gc])), 2, ffi.ptr(xy1), ffi.ptr(inf1)))
        xyb = np.zeros((1, 2 * k2 * LIMBS), dtype=np.uint64)
        infb = np.zeros(1, dtype=np.uint8)
        lib.check(lib.batch_normalize(ctx.handle, g2, ffi.ptr(gb), 1, ffi.ptr(xyb), ffi.ptr(infb)))
        return Proof(xy1[0].reshape(2, LIMBS), xyb[0].reshape(2, k2 * LIMBS), xy1[1].reshape(2, LIMBS),
                     (bool(inf1[0]), bool(infb[0]), bool(inf1[1])))
    finally:
        ctx.sync()
        for d in bufs:
            d.free()
