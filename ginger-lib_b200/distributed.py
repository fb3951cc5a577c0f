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
