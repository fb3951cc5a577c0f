"""Groth16 `create_proof` over several GPUs with COST-WEIGHTED PLACEMENT ("the independent
A / B / C / H / L MSMs of a proof are spread across GPUs").

The five long MSMs of prover.rs:270-337 are independent of each other and only H depends on the witness
map; splitting every one of them over all GPUs (groth16.ShardedParameters) leaves each GPU with five small
MSMs whose fixed costs (bucket reduction, window fold) do not shrink with the shard, and has every GPU
repeat the seven transforms.  Here a static plan, fixed when the key is loaded, gives

  * every MSM `round(cost / average load)` point-range shards (at least one), costs in units of one G1 point:
    the G2 query weighs 3.4 (Fq2) / 9.5 (Fq3) per point; shards go to the least loaded rank (longest
    processing time first), so on 8 GPUs B2 runs on four of them and A, B1, L, H on one each;
  * the three independent chains ifft -> coset_fft of the witness map (r1cs_to_qap.rs:121-161) to up to
    three ranks; the owner of the first H shard collects them over NCCL point-to-point, finishes h
    (g753_witness_map_tail_dev) and forwards slices of it to other owners of H shards, if any.

Every rank uploads only the slices of the assignment / evaluation vectors its own work needs, runs its
shards, the partial sums (4 G1 points + 1 G2 point per rank) are all-gathered and folded on every rank,
and every rank finishes the same proof (same bits as the single-GPU prover: the fold order differs, the
affine result does not).  NCCL when the process group is NCCL (device tensors on the context's stream);
any other backend (gloo in the CPU tests) is staged through host memory.
"""
import ctypes
import heapq

import numpy as np

from . import ffi
from .algebra import Bases
from .groth16 import LIMBS, ONE, Parameters, Proof, _limbs

G2_COST = {2: 3.4, 3: 9.5}        # measured accumulate + reduce time per point relative to G1 (Fq2, Fq3)
CHAIN_COST = 0.06                 # one ifft + coset_fft chain of the witness map, in G1-MSM units per domain point
H_DELAY = 0.18                    # what the H MSM waits for (upload, chains, exchange, tail of the witness map), same units
SHARD_FIXED = 0.08                # fixed cost of one more shard (bucket reduction, fold) in units of the full MSM
MSM_NAMES = ("b2", "a", "b1", "l", "h")


class ProofPlan:
    """Which rank runs which point range of which MSM, and which ranks run the witness-map chains.

    totals: name -> number of points of the long part of each query (a, b1, b2, l: num_aux; h: domain - 1 -
    num_inputs); shards[name] = [(rank, lo, hi), ...] tiles [0, totals[name])."""

    def __init__(self, world, totals, k2=2, domain=None):
        self.world, self.totals = world, dict(totals)
        cost = {nm: float(totals[nm]) * (G2_COST[k2] if nm == "b2" else 1.0) for nm in MSM_NAMES}
        target = sum(cost.values()) / world
        load = [(0.0, r) for r in range(world)]
        heapq.heapify(load)
        self.shards = {}
        for nm in sorted(MSM_NAMES, key=lambda x: -cost[x]):
            parts = int(min(world, max(1, round(cost[nm] / target)))) if totals[nm] > 0 else 1
            bounds = [totals[nm] * i // parts for i in range(parts + 1)]
            out = []
            picked = [heapq.heappop(load) for _ in range(parts)]      # `parts` DISTINCT least-loaded ranks
            for i, (ld, r) in enumerate(picked):
                out.append((r, bounds[i], bounds[i + 1]))
                heapq.heappush(load, (ld + cost[nm] / parts, r))
            # shards of one MSM that landed on the same rank are merged when adjacent
            out.sort(key=lambda t: t[1])
            merged = []
            for r, lo, hi in out:
                if merged and merged[-1][0] == r and merged[-1][2] == lo:
                    merged[-1] = (r, merged[-1][1], hi)
                else:
                    merged.append((r, lo, hi))
            self.shards[nm] = merged
        # H starts late (it waits for the witness map): when its owner ends up the most loaded rank, the tail of
        # H moves to the least loaded rank that holds no H shard yet
        dom_pts = float(domain if domain is not None else totals["h"])
        loads = dict((r, ld) for ld, r in load)
        if world > 1 and len(self.shards["h"]) == 1 and totals["h"] > 1:
            owner = self.shards["h"][0][0]
            late = loads[owner] + H_DELAY * dom_pts
            other = min((r for r in range(world) if r != owner), key=lambda r: loads[r])
            move = (late - loads[other] - SHARD_FIXED * totals["h"]) / 2.0
            if late >= max(loads.values()) and move > 0.05 * totals["h"]:
                cut = totals["h"] - int(min(move, 0.5 * totals["h"]))
                self.shards["h"] = [(owner, 0, cut), (other, cut, totals["h"])]
                loads[owner] -= totals["h"] - cut
                loads[other] += totals["h"] - cut + SHARD_FIXED * totals["h"]
                load = [(ld, r) for r, ld in loads.items()]
        self.finisher = self.shards["h"][0][0]
        # the chains go to the least loaded ranks, the finisher first among equals
        dom = dom_pts
        loads = dict((r, ld) for ld, r in load)
        self.chain_owner = {}
        for ch in ("a", "b", "c"):
            r = min(range(world), key=lambda x: (loads[x], x != self.finisher, x))
            self.chain_owner[ch] = r
            loads[r] += CHAIN_COST * dom
        self.load = loads

    def shards_of(self, rank):
        """name -> (lo, hi) of the shard this rank holds (at most one per MSM), for key loading"""
        out = {}
        for nm, parts in self.shards.items():
            mine = [(lo, hi) for r, lo, hi in parts if r == rank]
            assert len(mine) <= 1
            if mine:
                out[nm] = mine[0]
        return out

    def describe(self):
        return {"shards": {nm: [list(t) for t in parts] for nm, parts in self.shards.items()},
                "chains": dict(self.chain_owner), "finisher": self.finisher,
                "load_in_g1_points": {str(r): round(v) for r, v in sorted(self.load.items())}}


class _Comm:
    """point-to-point / all-gather of device buffers: NCCL on the context's stream, or host-staged"""

    def __init__(self, ctx, pg):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.ctx, self.pg = torch, dist, ctx, pg
        self.world = dist.get_world_size(pg)
        self.rank = dist.get_rank(pg)
        self.nccl = dist.get_backend(pg) == "nccl"
        self._inflight = []
        if self.nccl:
            self.dev = torch.device("cuda", ctx.device)
            self.stream = torch.cuda.ExternalStream(int(ctx.lib.stream(ctx.handle)), device=self.dev)

    def buffer(self, nbytes):
        """a communicable device buffer: (torch tensor or None, device pointer)"""
        if self.nccl:
            t = self.torch.empty(max(nbytes, 8) // 8, dtype=self.torch.int64, device=self.dev)
            return t, ctypes.c_void_p(t.data_ptr())
        p = ctypes.c_void_p()
        self.ctx.lib.check(self.ctx.lib.dev_alloc(self.ctx.handle, max(nbytes, 8), ctypes.byref(p)))
        return None, p

    def free(self, buf):
        t, p = buf
        if t is None and p:
            self.ctx.lib.dev_free(self.ctx.handle, p)

    def _host(self, p, first, nbytes):
        out = np.empty(nbytes // 8, dtype=np.uint64)
        self.ctx.lib.check(self.ctx.lib.d2h(self.ctx.handle, ffi.ptr(out), ctypes.c_void_p(p.value + first), nbytes))
        return out

    # NCCL: isend / irecv are ordered after the work queued so far on the context's stream and do NOT
    # block it; wait() makes the context's stream (not the host) wait for the transfer.  All operations
    # of a rank go through one NCCL queue in issue order, so a rank always issues the sends others wait
    # for before the receives it waits for itself.
    def send(self, buf, first, nbytes, dst):
        t, p = buf
        if self.nccl:
            with self.torch.cuda.stream(self.stream):
                self._inflight.append(self.dist.isend(t[first // 8:(first + nbytes) // 8], dst, group=self.pg))
        else:
            self.dist.send(self.torch.from_numpy(self._host(p, first, nbytes).view(np.int64)), dst, group=self.pg)

    def recv(self, buf, first, nbytes, src):
        """returns a handle for wait()"""
        t, p = buf
        if self.nccl:
            with self.torch.cuda.stream(self.stream):
                return self.dist.irecv(t[first // 8:(first + nbytes) // 8], src, group=self.pg)
        h = self.torch.empty(nbytes // 8, dtype=self.torch.int64)
        self.dist.recv(h, src, group=self.pg)
        arr = np.ascontiguousarray(h.numpy().view(np.uint64))
        self.ctx.lib.check(self.ctx.lib.h2d(self.ctx.handle, ctypes.c_void_p(p.value + first), ffi.ptr(arr), nbytes))
        return None

    def wait(self, handles):
        if self.nccl:
            with self.torch.cuda.stream(self.stream):
                for h in handles:
                    if h is not None:
                        h.wait()

    def done(self):
        """end of a proof: the sends issued during it have been consumed (the all-gather came after them)"""
        self._inflight = []

    def all_gather(self, buf, nbytes, out_buf):
        """out[r * nbytes ...] = rank r's buf"""
        t, p = buf
        to, po = out_buf
        if self.nccl:
            with self.torch.cuda.stream(self.stream):
                self.dist.all_gather_into_tensor(to[:self.world * nbytes // 8], t[:nbytes // 8], group=self.pg)
        else:
            mine = self.torch.from_numpy(self._host(p, 0, nbytes).view(np.int64))
            parts = [self.torch.empty_like(mine) for _ in range(self.world)]
            self.dist.all_gather(parts, mine, group=self.pg)
            arr = np.ascontiguousarray(np.concatenate([x.numpy().view(np.uint64) for x in parts]))
            self.ctx.lib.check(self.ctx.lib.h2d(self.ctx.handle, po, ffi.ptr(arr), self.world * nbytes))


class PlacedParameters(Parameters):
    """A proving key placed over the ranks of a process group by a ProofPlan: every rank keeps the short
    head of each query (query[0..num_inputs] and the vk points) and the shards the plan gives it.

    heads: {"a", "b1", "b2", "h"} -> (coords of the first num_inputs bases, infinity or None);
    shards: name -> Bases over plan.shards_of(rank)[name] (the long part: query[num_inputs + lo ..],
    l_query[lo ..])."""

    def __init__(self, ctx, g1, g2, field, alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2, heads, shards, num_inputs,
                 plan, process_group=None, precompute=1):
        import torch.distributed as dist
        self.ctx, self.g1, self.g2, self.field, self.num_inputs = ctx, g1, g2, field, num_inputs
        self.plan = plan
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
        pg = process_group if process_group is not None else dist.group.WORLD
        self.comm = _Comm(ctx, pg)
        mine = plan.shards_of(self.comm.rank)
        assert set(mine) == set(shards), "key shards do not match the plan"
        self.shards = {}
        for name, bases in shards.items():
            lo, hi = mine[name]
            assert len(bases) == hi - lo
            if precompute is not None and precompute != 1 and len(bases) >= 1 << 12:
                bases.precompute(precompute)
            self.shards[name] = (bases, lo, hi)
        self.delta_g1 = ffi.as_u64(delta_g1).reshape(1, 2 * LIMBS)
        from .algebra import Context
        self.ctx2 = Context(ctx.device, library=ctx.lib)
        self.fresh = Bases(self.ctx2, g1, np.concatenate([self.delta_g1] * 3), np.zeros(3, dtype=np.uint8))
        self.sharder = None
        self._pw = None

    def free(self):
        for b in [v[0] for v in self.shards.values()] + [self.a_small, self.b1_small, self.b2_small, self.h_head, self.fresh]:
            b.free()
        if self._pw is not None:
            for buf in self._pw["bufs"]:
                self.comm.free(buf)
            self._pw = None
        self.ctx2.close()


def create_proof_placed(params, full_assignment, a, b, c, d1, d2, d3, r, s, timings=None, profile=None):
    """prover.rs:201-345 after constraint synthesis, over the ranks of params' process group; arguments
    and result as groth16.create_proof (every rank passes the same inputs and returns the same proof)"""
    import time
    ctx, ctx2, lib, field = params.ctx, params.ctx2, params.ctx.lib, params.field
    g1, g2, ni, plan, comm = params.g1, params.g2, params.num_inputs, params.plan, params.comm
    rank, world = comm.rank, comm.world
    k2 = ffi.GROUP_K[g2]
    t0 = time.perf_counter()
    a, b, c = (ffi.as_u64(v).reshape(-1, LIMBS) for v in (a, b, c))
    z = ffi.as_u64(full_assignment).reshape(-1, LIMBS)
    n = a.shape[0]
    log_n = n.bit_length() - 1
    n_vars = z.shape[0]
    my_chains = [ch for ch in ("a", "b", "c") if plan.chain_owner[ch] == rank]
    finisher = plan.finisher
    # G1 result slots: a, b1, l, h; one G2 slot: b2
    G1_SLOTS = {"a": 0, "b1": 1, "l": 2, "h": 3}
    part_bytes = 4 * 288 + 3 * k2 * 96

    pw = params._pw
    if pw is None or pw["shape"] != (n, n_vars):
        if pw is not None:
            for buf in pw["bufs"]:
                comm.free(buf)
        chains = {ch: comm.buffer(n * 96) for ch in ("a", "b", "c") if ch in my_chains or rank == finisher}
        d_h = comm.buffer((n + 1) * 96)
        d_z = comm.buffer((n_vars + 1) * 96)
        part = comm.buffer(part_bytes)
        gathered = comm.buffer(world * part_bytes)
        small = comm.buffer((2 * (ni + 2) + 3 + 3 * 16 + 3 * k2 * 4 + 8 * 3 * k2 + 8) * 96)
        pw = params._pw = {"shape": (n, n_vars), "chains": chains, "h": d_h, "z": d_z, "part": part, "gathered": gathered,
                           "small": small, "bufs": list(chains.values()) + [d_h, d_z, part, gathered, small]}
    chains, d_h, d_z, part, gathered = pw["chains"], pw["h"], pw["z"], pw["part"], pw["gathered"]
    at = lambda buf, i: ctypes.c_void_p(buf[1].value + 96 * i)
    sb = pw["small"]
    off = [0]

    def carve(count):
        p = at(sb, off[0])
        off[0] += count
        return p
    sr, ss, sc3 = carve(ni + 2), carve(ni + 2), carve(3)
    out1 = carve(3 * 16)
    out2 = carve(3 * k2 * 4)
    fold1 = carve(8 * 3 * k2)       # the partial points of one slot, contiguous (world <= 8)
    slot1 = lambda i: ctypes.c_void_p(out1.value + 288 * i)
    slot2 = lambda i: ctypes.c_void_p(out2.value + 3 * k2 * 96 * i)

    def put(dst, host):
        host = np.ascontiguousarray(ffi.as_u64(host).reshape(-1, LIMBS))
        lib.check(lib.h2d(ctx.handle, dst, ffi.ptr(host), host.shape[0] * 96))

    try:
        dd = np.concatenate([_limbs(d1), _limbs(d2), _limbs(d3), _limbs(r), _limbs(s)])
        dm = np.zeros_like(dd)
        lib.check(lib.field_op(ctx.handle, field, ffi.OP_TO_MONT, ffi.ptr(dd), None, ffi.ptr(dm), 5))
        rs = np.zeros((1, LIMBS), dtype=np.uint64)
        lib.check(lib.field_op(ctx.handle, field, ffi.OP_MUL, ffi.ptr(dm[3:4]), ffi.ptr(dm[4:5]), ffi.ptr(rs), 1))
        zero = np.zeros((1, LIMBS), dtype=np.uint64)
        lib.check(lib.field_op(ctx.handle, field, ffi.OP_SUB, ffi.ptr(zero), ffi.ptr(rs), ffi.ptr(rs), 1))
        lib.check(lib.field_op(ctx.handle, field, ffi.OP_FROM_MONT, ffi.ptr(rs), None, ffi.ptr(rs), 1))
        put(sc3, np.concatenate([_limbs(s), _limbs(r), rs]))
        # ---- A. the finisher posts the receives of the other ranks' chains ------------------------------
        chain_recvs = []
        if rank == finisher:
            for ch in ("a", "b", "c"):
                if plan.chain_owner[ch] != finisher:
                    chain_recvs.append(comm.recv(chains[ch], 0, n * 96, plan.chain_owner[ch]))
        # ---- B. my chains ifft -> coset_fft (r1cs_to_qap.rs:121-161), sent to the finisher -------------
        for ch, host in (("a", a), ("b", b), ("c", c)):
            if ch in my_chains:
                put(chains[ch][1], host)
                lib.check(lib.ntt_dev(ctx.handle, field, chains[ch][1], log_n, ffi.IFFT))
                lib.check(lib.ntt_dev(ctx.handle, field, chains[ch][1], log_n, ffi.COSET_FFT))
                if rank != finisher:
                    comm.send(chains[ch], 0, n * 96, finisher)
        # ---- C. everybody else posts the receives of what the finisher will send: the slice of h its H
        #         shard multiplies, and h[0..ni] (the head rides in a short MSM on every rank) ----------
        h_parts = plan.shards["h"]
        slice_recvs, head_recvs = [], []
        if rank != finisher:
            for owner, lo, hi in h_parts:
                if owner == rank:
                    slice_recvs.append(comm.recv(d_h, (ni + lo) * 96, (hi - lo) * 96, finisher))
            head_recvs.append(comm.recv(d_h, 0, ni * 96, finisher))
        if timings is not None:
            ctx.sync()
            timings["chains"] = time.perf_counter() - t0
            t0 = time.perf_counter()
        # ---- D. the slices of the assignment my shards multiply, into_repr on the device
        #         (prover.rs:241-267), and the MSM shards that do not depend on h ----------------------
        need = [(lo, hi) for nm, (bs, lo, hi) in params.shards.items() if nm != "h"]
        z_lo = ni + min(lo for lo, _ in need) if need else n_vars
        z_hi = ni + max(hi for _, hi in need) if need else n_vars
        put(at(d_z, 0), z[:ni])                                       # the inputs (short MSMs), every rank
        lib.check(lib.vec_op_dev(ctx.handle, field, ffi.OP_FROM_MONT, at(d_z, 0), None, ni))
        if z_hi > z_lo:
            put(at(d_z, z_lo), z[z_lo:z_hi])
            lib.check(lib.vec_op_dev(ctx.handle, field, ffi.OP_FROM_MONT, at(d_z, z_lo), None, z_hi - z_lo))
        for dv, blind in ((sr, r), (ss, s)):
            put(dv, ONE)
            if ni > 1:
                lib.check(lib.d2d(ctx.handle, ctypes.c_void_p(dv.value + 96), at(d_z, 1), (ni - 1) * 96))
            put(ctypes.c_void_p(dv.value + 96 * ni), _limbs(blind))
            put(ctypes.c_void_p(dv.value + 96 * (ni + 1)), ONE)
        # the partial sums of the MSMs I do not run are the point at infinity
        inf1 = np.zeros((1, 3 * LIMBS), dtype=np.uint64)
        inf1[0, LIMBS:2 * LIMBS] = 1           # any (0 : y : 0) is infinity; the fold only tests Z
        for i in range(4):
            put(ctypes.c_void_p(part[1].value + 288 * i), inf1)
        inf2 = np.zeros((1, 3 * k2 * LIMBS), dtype=np.uint64)
        inf2[0, k2 * LIMBS] = 1
        put(ctypes.c_void_p(part[1].value + 4 * 288), inf2)

        def run_shard(nm):
            if nm not in params.shards:
                return
            bases, lo, hi = params.shards[nm]
            src = at(d_h, ni + lo) if nm == "h" else at(d_z, ni + lo)
            dst = ctypes.c_void_p(part[1].value + (4 * 288 if nm == "b2" else 288 * G1_SLOTS[nm]))
            lib.check(lib.msm_dev(ctx.handle, bases.handle, 0, hi - lo, src, dst))
            if profile is not None:
                ctx.sync()
                profile[nm] = dict(ctx.last_msm_phases(), points=hi - lo, plan=ctx.last_msm_plan(), rank=rank)
        def msm(cx, bases, count, d_scalars, d_out):
            lib.check(lib.msm_dev(cx.handle, bases.handle, 0, min(count, len(bases)), d_scalars, d_out))
        for nm in ("b2", "a", "b1", "l"):
            run_shard(nm)
        # the short MSMs over the inputs and the vk points depend on nothing that is exchanged
        msm(ctx, params.a_small, ni + 2, sr, slot1(0))
        msm(ctx, params.b1_small, ni + 2, ss, slot1(2))
        msm(ctx, params.b2_small, ni + 2, ss, slot2(0))
        if timings is not None:
            ctx.sync()
            timings["z_msms"] = time.perf_counter() - t0
            t0 = time.perf_counter()
        # ---- E. the finisher completes h and forwards what the others wait for ----------------------------
        if rank == finisher:
            comm.wait(chain_recvs)
            lib.check(lib.witness_map_tail_dev(ctx.handle, field, chains["a"][1], chains["b"][1], chains["c"][1], log_n,
                                               ffi.ptr(dm), d_h[1]))
            lib.check(lib.vec_op_dev(ctx.handle, field, ffi.OP_FROM_MONT, d_h[1], None, n + 1))
            for dst in range(world):
                if dst == finisher:
                    continue
                for owner, lo, hi in h_parts:
                    if owner == dst:
                        comm.send(d_h, (ni + lo) * 96, (hi - lo) * 96, dst)
                comm.send(d_h, 0, ni * 96, dst)
        # ---- F. my H shard ------------------------------------------------------------------------------------
        comm.wait(slice_recvs)
        run_shard("h")
        comm.wait(head_recvs)
        if timings is not None:
            ctx.sync()
            timings["h_and_h_msm"] = time.perf_counter() - t0
            t0 = time.perf_counter()
        comm.all_gather(part, part_bytes, gathered)
        # fold slot by slot: gather the `world` partial points of a slot into a contiguous run, sum them
        def fold(slot_off, nbytes, group, dst):
            assert world <= 8
            for rr in range(world):
                lib.check(lib.d2d(ctx.handle, ctypes.c_void_p(fold1.value + rr * nbytes),
                                  ctypes.c_void_p(gathered[1].value + rr * part_bytes + slot_off), nbytes))
            lib.check(lib.points_sum_dev(ctx.handle, group, fold1, world, dst))
        # A (prover.rs:270-283) -> slots 0, 1; B in G1 (:286-299) -> slots 2, 3; g_a, g1_b -> slots 8, 9
        fold(288 * G1_SLOTS["a"], 288, g1, slot1(1))
        fold(288 * G1_SLOTS["b1"], 288, g1, slot1(3))
        fold(288 * G1_SLOTS["l"], 288, g1, slot1(6))
        fold(288 * G1_SLOTS["h"], 288, g1, slot1(5))
        fold(4 * 288, 3 * k2 * 96, g2, slot2(1))

        lib.check(lib.points_sum_dev(ctx.handle, g1, slot1(0), 2, slot1(8)))
        lib.check(lib.points_sum_dev(ctx.handle, g1, slot1(2), 2, slot1(9)))
        lib.check(lib.ctx_wait(ctx2.handle, ctx.handle))
        msm(ctx, params.h_head, min(ni, n + 1), at(d_h, 0), slot1(4))
        lib.check(lib.points_sum_dev(ctx.handle, g2, slot2(0), 2, slot2(2)))
        # second context: s * g_a + r * g1_b - rs * delta_g1 (:322-329) as one MSM over fresh bases
        ga_b1 = np.empty((2, 3 * LIMBS), dtype=np.uint64)
        lib.check(lib.d2h(ctx2.handle, ffi.ptr(ga_b1), slot1(8), 2 * 3 * 96))
        xy = np.zeros((2, 2 * LIMBS), dtype=np.uint64)
        inf = np.zeros(2, dtype=np.uint8)
        lib.check(lib.batch_normalize(ctx2.handle, g1, ffi.ptr(ga_b1), 2, ffi.ptr(xy), ffi.ptr(inf)))
        lib.check(lib.bases_update(ctx2.handle, params.fresh.handle, 0, 2, ffi.ptr(xy), ffi.ptr(inf)))
        lib.check(lib.msm_dev(ctx2.handle, params.fresh.handle, 0, 3, sc3, slot1(7)))
        lib.check(lib.ctx_wait(ctx.handle, ctx2.handle))
        lib.check(lib.points_sum_dev(ctx.handle, g1, slot1(4), 4, slot1(10)))
        gc = np.empty((1, 3 * LIMBS), dtype=np.uint64)
        lib.check(lib.d2h(ctx.handle, ffi.ptr(gc), slot1(10), 288))
        gb2 = np.empty((1, 3 * k2 * LIMBS), dtype=np.uint64)
        lib.check(lib.d2h(ctx.handle, ffi.ptr(gb2), slot2(2), 3 * k2 * 96))
        xyc = np.zeros((1, 2 * LIMBS), dtype=np.uint64)
        infc = np.zeros(1, dtype=np.uint8)
        lib.check(lib.batch_normalize(ctx.handle, g1, ffi.ptr(gc), 1, ffi.ptr(xyc), ffi.ptr(infc)))
        xyb = np.zeros((1, 2 * k2 * LIMBS), dtype=np.uint64)
        infb = np.zeros(1, dtype=np.uint8)
        lib.check(lib.batch_normalize(ctx.handle, g2, ffi.ptr(gb2), 1, ffi.ptr(xyb), ffi.ptr(infb)))
        if timings is not None:
            ctx.sync()
            timings["exchange_and_assembly"] = time.perf_counter() - t0
        comm.done()
        return Proof(xy[0].reshape(2, LIMBS), xyb[0].reshape(2, k2 * LIMBS), xyc[0].reshape(2, LIMBS),
                     (bool(inf[0]), bool(infb[0]), bool(infc[0])))
    finally:
        ctx2.sync()
