"""Multi-GPU sharding of the MSM: one process per GPU, point-range shards, one tiny
exchange.

sum_i s_i P_i is associative, so rank g owns bases/scalars [lo_g, hi_g), runs a complete local MSM
on its own GPU and produces one partial projective point (3*k*12 limbs, 288 B for G1).  The only
collective is an all-gather of those partial points (N x 288 B over NVLink/NVSwitch with NCCL),
followed by a fold of N points on every rank's device (g753_points_sum_dev), so every rank ends up
with the full result - the shape `ncclAllReduce` would have if NCCL had a user-defined group-law
reduction.  There is no data-path collective on the bases or scalars.

The reference has no multi-process code (its parallelism is one rayon task per window,
variable_base.rs:30-32); the contract here is "N-rank result == 1-rank result == oracle".
"""
import ctypes

import numpy as np

from . import ffi

LIMBS = 12


def shard_range(n, rank, world):
    """contiguous point range of `rank`: sizes differ by at most one"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedMSM:
    """`multi_scalar_mul` over a key sharded across the ranks of a torch.distributed group.

    `bases` is this rank's resident shard (algebra.Bases).  With the NCCL backend the partial
    points stay on the device (all_gather_into_tensor on the context's stream); with any other
    backend (gloo in the CPU tests) they are staged through host memory - same fold, same result.
    """

    def __init__(self, ctx, bases, process_group=None):
        import torch.distributed as dist
        self.dist = dist
        self.ctx, self.bases, self.pg = ctx, bases, process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.k = bases.k
        self.point_limbs = 3 * self.k * LIMBS
        self.backend = dist.get_backend(process_group) if dist.is_initialized() else None

    def multi_scalar_mul(self, scalars_shard, first=0):
        """scalars_shard: this rank's (m, 12) canonical scalars, zipped with bases[first..] of the shard.
        Returns the full (3, k*12) projective result on every rank."""
        import torch
        lib, ctx = self.ctx.lib, self.ctx
        scalars = ffi.as_u64(scalars_shard).reshape(-1, LIMBS)
        count = min(len(self.bases) - first, scalars.shape[0])
        part = np.zeros((3, self.k * LIMBS), dtype=np.uint64)
        lib.check(lib.msm(ctx.handle, self.bases.handle, first, count, ffi.ptr(scalars), ffi.ptr(part)))
        if self.world == 1:
            return part
        if self.backend == "nccl":
            dev = torch.device("cuda", ctx.device)
            mine = torch.from_numpy(part.view(np.int64).reshape(-1)).to(dev)
            allp = torch.empty(self.point_limbs * self.world, dtype=torch.int64, device=dev)
            self.dist.all_gather_into_tensor(allp, mine, group=self.pg)
            out = torch.empty(self.point_limbs, dtype=torch.int64, device=dev)
            torch.cuda.current_stream(dev).synchronize()
            lib.check(lib.points_sum_dev(ctx.handle, self.bases.group, ctypes.c_void_p(allp.data_ptr()),
                                         self.world, ctypes.c_void_p(out.data_ptr())))
            lib.check(lib.sync(ctx.handle))
            return out.cpu().numpy().view(np.uint64).reshape(3, self.k * LIMBS)
        # host-staged exchange (gloo): gather the partial points, fold them through the library
        mine = torch.from_numpy(part.view(np.int64).reshape(-1).copy())
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(parts, mine, group=self.pg)
        allp = np.concatenate([p.numpy().view(np.uint64) for p in parts])
        return fold_points(ctx, self.bases.group, allp, self.world)


def fold_points(ctx, group, points_xyz, count):
    """sum of `count` projective points given as host limbs -> (3, k*12) projective"""
    lib = ctx.lib
    k = ffi.GROUP_K[group]
    pts = ffi.as_u64(points_xyz).reshape(-1)
    nbytes = count * 3 * k * 96
    d_in, d_out = ctypes.c_void_p(), ctypes.c_void_p()
    lib.check(lib.dev_alloc(ctx.handle, nbytes, ctypes.byref(d_in)))
    lib.check(lib.dev_alloc(ctx.handle, 3 * k * 96, ctypes.byref(d_out)))
    try:
        lib.check(lib.h2d(ctx.handle, d_in, ffi.ptr(pts), nbytes))
        lib.check(lib.points_sum_dev(ctx.handle, group, d_in, count, d_out))
        out = np.zeros((3, k * LIMBS), dtype=np.uint64)
        lib.check(lib.d2h(ctx.handle, ffi.ptr(out), d_out, 3 * k * 96))
    finally:
        lib.dev_free(ctx.handle, d_in)
        lib.dev_free(ctx.handle, d_out)
    return out


class ShardedEvaluationDomain:
    """`EvaluationDomain` transforms over a vector sharded across the ranks of a process group
    (four-step NTT).  Rank g holds the `cols x n1` input shard
    local[i2l][i1] = x[i1*n2 + g*cols + i2l]; a transform leaves the `rows x n2` output shard
    local[k1l][k2] = X[(g*rows + k1l) + n1*k2].  The only collective is one all-to-all per transform
    (`all_to_all_single` on device tensors with NCCL; host-staged with gloo in the CPU tests)."""

    def __init__(self, ctx, field, log_n, process_group=None):
        import torch.distributed as dist
        self.dist, self.ctx, self.field, self.log_n, self.pg = dist, ctx, field, log_n, process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.backend = dist.get_backend(process_group) if dist.is_initialized() else None
        h = ctypes.c_void_p()
        ctx.lib.check(ctx.lib.ntt_shard_create(ctx.handle, field, log_n, self.world, self.rank, ctypes.byref(h)))
        self.plan = h
        v = [ctypes.c_size_t() for _ in range(4)]
        ctx.lib.check(ctx.lib.ntt_shard_shape(h, *[ctypes.byref(x) for x in v]))
        self.n1, self.n2, self.cols, self.rows = (int(x.value) for x in v)
        self.local = (1 << log_n) // self.world

    # ---- layout helpers (host side, for tests / callers that start from a whole vector) ----------
    def scatter(self, full):
        """whole natural-order vector (n, 12) -> this rank's input shard (local, 12)"""
        x = ffi.as_u64(full).reshape(self.n1, self.n2, LIMBS)            # x[i1][i2]
        lo = self.rank * self.cols
        return np.ascontiguousarray(x[:, lo:lo + self.cols].transpose(1, 0, 2)).reshape(-1, LIMBS)

    def gather(self, shard):
        """output shards of all ranks -> whole natural-order vector (n, 12), on every rank"""
        import torch
        mine = torch.from_numpy(np.ascontiguousarray(shard).view(np.int64).reshape(-1).copy())
        if self.world == 1:
            parts = [mine]
        elif self.backend == "nccl":
            dev = torch.device("cuda", self.ctx.device)
            buf = [torch.empty_like(mine, device=dev) for _ in range(self.world)]
            self.dist.all_gather(buf, mine.to(dev), group=self.pg)
            parts = [b.cpu() for b in buf]
        else:
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            self.dist.all_gather(parts, mine, group=self.pg)
        out = np.empty((self.n2, self.n1, LIMBS), dtype=np.uint64)          # X[k2][k1], k = k1 + n1*k2
        for g, p in enumerate(parts):
            blk = p.numpy().view(np.uint64).reshape(self.rows, self.n2, LIMBS)   # [k1l][k2]
            out[:, g * self.rows:(g + 1) * self.rows] = blk.transpose(1, 0, 2)
        return out.reshape(-1, LIMBS)

    # ---- the transform -----------------------------------------------------------------------------
    def transform_dev(self, d_data, d_send, d_recv, mode, exchange):
        """d_data / d_send / d_recv: device pointers to `local` elements each; `exchange(d_send, d_recv)`
        performs the all-to-all of world blocks of rows*cols elements"""
        lib, ctx = self.ctx.lib, self.ctx
        lib.check(lib.ntt_shard_step1(ctx.handle, self.plan, d_data, d_send, mode))
        exchange()
        lib.check(lib.ntt_shard_step2(ctx.handle, self.plan, d_recv, d_data, mode))

    def transform(self, shard, mode):
        """host shard in, host shard out (any backend)"""
        import torch
        lib, ctx = self.ctx.lib, self.ctx
        shard = ffi.as_u64(shard).reshape(self.local, LIMBS)
        nbytes = self.local * 96
        if self.backend == "nccl":
            dev = torch.device("cuda", ctx.device)
            t_data = torch.from_numpy(shard.view(np.int64).reshape(-1).copy()).to(dev)
            t_send, t_recv = torch.empty_like(t_data), torch.empty_like(t_data)
            torch.cuda.synchronize(dev)

            def exchange():
                lib.check(lib.sync(ctx.handle))
                self.dist.all_to_all_single(t_recv, t_send, group=self.pg)
                torch.cuda.synchronize(dev)
            self.transform_dev(ctypes.c_void_p(t_data.data_ptr()), ctypes.c_void_p(t_send.data_ptr()),
                               ctypes.c_void_p(t_recv.data_ptr()), mode, exchange)
            lib.check(lib.sync(ctx.handle))
            return t_data.cpu().numpy().view(np.uint64).reshape(self.local, LIMBS)
        bufs = [ctypes.c_void_p() for _ in range(3)]
        for b in bufs:
            lib.check(lib.dev_alloc(ctx.handle, nbytes, ctypes.byref(b)))
        try:
            lib.check(lib.h2d(ctx.handle, bufs[0], ffi.ptr(shard), nbytes))

            def exchange():
                send = np.empty((self.local, LIMBS), dtype=np.uint64)
                lib.check(lib.d2h(ctx.handle, ffi.ptr(send), bufs[1], nbytes))
                if self.world > 1:
                    t_send = torch.from_numpy(send.view(np.int64).reshape(-1))
                    t_recv = torch.empty_like(t_send)
                    if self.backend == "gloo":   # gloo has no all_to_all_single on CPU tensors in every build
                        ins = list(t_send.chunk(self.world))
                        outs = [torch.empty_like(c) for c in ins]
                        reqs = []
                        for peer in range(self.world):
                            if peer == self.rank:
                                outs[peer].copy_(ins[peer])
                                continue
                            reqs.append(self.dist.isend(ins[peer].contiguous(), peer, group=self.pg))
                            reqs.append(self.dist.irecv(outs[peer], peer, group=self.pg))
                        for q in reqs:
                            q.wait()
                        t_recv = torch.cat(outs)
                    else:
                        self.dist.all_to_all_single(t_recv, t_send, group=self.pg)
                    recv = t_recv.numpy().view(np.uint64).reshape(self.local, LIMBS)
                else:
                    recv = send
                lib.check(lib.h2d(ctx.handle, bufs[2], ffi.ptr(np.ascontiguousarray(recv)), nbytes))
            self.transform_dev(bufs[0], bufs[1], bufs[2], mode, exchange)
            out = np.empty((self.local, LIMBS), dtype=np.uint64)
            lib.check(lib.d2h(ctx.handle, ffi.ptr(out), bufs[0], nbytes))
            return out
        finally:
            for b in bufs:
                lib.dev_free(ctx.handle, b)

    def close(self):
        if getattr(self, "plan", None):
            self.ctx.lib.ntt_shard_destroy(self.ctx.handle, self.plan)
            self.plan = None


class FusedShardedNTT:
    """The sharded transform with the exchange fused into the compute kernel (NVLink peer memory).

    Two row buffers per rank live in symmetric memory (`torch.distributed._symmetric_memory`: every
    rank maps every peer's buffer).  A transform reads the current buffer, and the LAST butterfly pass
    of its column transforms stores each output directly into the OTHER buffer of the destination
    GPU, already in row order (`g753_ntt_shard_step1_fused`); one device-side barrier later the row
    transforms run in place there (`g753_ntt_shard_step2_local`).  No pack / unpack kernels and no
    collective call: the only cross-GPU traffic is the kernel's own stores.  Ping-ponging the two
    buffers makes one barrier per transform sufficient when transforms chain (n1 == n2)."""

    def __init__(self, dom, stream):
        import torch
        import torch.distributed._symmetric_memory as symm_mem
        self.torch, self.dom, self.stream = torch, dom, stream
        dev = torch.device("cuda", dom.ctx.device)
        group = dom.pg if dom.pg is not None else dom.dist.group.WORLD
        self.bufs = [symm_mem.empty(dom.local * LIMBS, dtype=torch.int64, device=dev) for _ in range(2)]
        self.hdls = [symm_mem.rendezvous(b, group) for b in self.bufs]
        self.peers = [(ctypes.c_void_p * dom.world)(*[int(h.buffer_ptrs[r]) for r in range(dom.world)]) for h in self.hdls]
        self.cur = 0

    def load(self, shard):
        """host input shard (local, 12) -> the current buffer"""
        t = self.torch.from_numpy(ffi.as_u64(shard).reshape(-1).view(np.int64).copy())
        with self.torch.cuda.stream(self.stream):
            self.bufs[self.cur].copy_(t, non_blocking=False)
        self.stream.synchronize()

    def store(self):
        self.stream.synchronize()
        return self.bufs[self.cur].cpu().numpy().view(np.uint64).reshape(self.dom.local, LIMBS)

    def transform(self, mode):
        lib, ctx, dom = self.dom.ctx.lib, self.dom.ctx, self.dom
        src, dst = self.cur, 1 - self.cur
        lib.check(lib.ntt_shard_step1_fused(ctx.handle, dom.plan, ctypes.c_void_p(self.bufs[src].data_ptr()),
                                            self.peers[dst], mode))
        with self.torch.cuda.stream(self.stream):
            self.hdls[dst].barrier(channel=0)
        lib.check(lib.ntt_shard_step2_local(ctx.handle, dom.plan, ctypes.c_void_p(self.bufs[dst].data_ptr()), mode))
        self.cur = dst
