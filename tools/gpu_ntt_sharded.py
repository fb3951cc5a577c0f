#!/usr/bin/env python3
"""Sharded four-step NTT on N GPUs (torchrun): parity against the single-GPU transform of the gathered
vector for all four modes, then device-timed throughput of the chained transform (step1 -> NCCL
all-to-all -> step2), max over ranks."""
import ctypes, importlib, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
G = importlib.import_module("ginger-lib_b200")
D = importlib.import_module("ginger-lib_b200.distributed")
import bench
ffi = G.ffi
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
ctx = G.Context(local, stream=stream.cuda_stream)
lib = ctx.lib
field = ffi.FIELD_MNT4_FR
out = {"world": world}
# parity at 2^16 (n1 = n2) and 2^15 (n1 = 2 n2)
for log_n in (16, 15):
    n = 1 << log_n
    raw = bench.random_scalars(n, 0x55 + log_n)
    raw[:, 11] &= np.uint64(0xFFFF)
    dom = D.ShardedEvaluationDomain(ctx, field, log_n)
    full = G.EvaluationDomain.new(field, n, ctx=ctx)
    for mode, name in ((ffi.FFT, "fft"), (ffi.IFFT, "ifft"), (ffi.COSET_FFT, "coset_fft"), (ffi.COSET_IFFT, "coset_ifft")):
        got = dom.gather(dom.transform(dom.scatter(raw), mode))
        want = full._run(raw, mode)
        assert (got == want).all(), (rank, log_n, name)
    dom.close()
out["parity"] = "2^16 and 2^15, four modes, sharded == single-GPU, every limb"
# timing at 2^log_n
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << log_n
dom = D.ShardedEvaluationDomain(ctx, field, log_n)
dev = torch.device("cuda", local)
with torch.cuda.stream(stream):
    t_data = torch.from_numpy(dom.scatter(np.zeros((n, 12), dtype=np.uint64) + 1).view(np.int64).reshape(-1).copy()).to(dev)
    t_send, t_recv = torch.empty_like(t_data), torch.empty_like(t_data)
stream.synchronize()
p = lambda t: ctypes.c_void_p(t.data_ptr())


def exchange():
    with torch.cuda.stream(stream):
        dist.all_to_all_single(t_recv, t_send)


def one(mode):
    dom.transform_dev(p(t_data), p(t_send), p(t_recv), mode, exchange)


for _ in range(3):
    one(ffi.FFT)
stream.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record(stream)
for _ in range(reps):
    one(ffi.FFT)
e1.record(stream)
stream.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
out.update({"log_n": log_n, "ms": float(ms.item()), "elements_per_s": n / (float(ms.item()) * 1e-3),
            "exchange_bytes_per_rank": (n // world) * 96 * (world - 1) // world})
# ---- fused exchange over peer memory (symmetric memory): parity, then the same timing ------------
try:
    fused = D.FusedShardedNTT(dom, stream)
    raw = bench.random_scalars(n, 0x99)
    raw[:, 11] &= np.uint64(0xFFFF)
    for mode, name in ((ffi.FFT, "fft"), (ffi.COSET_IFFT, "coset_ifft")):
        fused.load(dom.scatter(raw))
        fused.transform(mode)
        got = fused.store()
        want = dom.transform(dom.scatter(raw), mode)
        assert (got == want).all(), (rank, "fused", name)
    if log_n % 2 == 0:   # chaining without re-layout
        fused.load(dom.scatter(raw))
        fused.transform(ffi.IFFT)
        fused.transform(ffi.COSET_FFT)
        want = dom.transform(dom.transform(dom.scatter(raw), ffi.IFFT), ffi.COSET_FFT)
        assert (fused.store() == want).all(), (rank, "fused chain")
    for _ in range(3):
        fused.transform(ffi.FFT)
    stream.synchronize()
    dist.barrier()
    e0.record(stream)
    for _ in range(reps):
        fused.transform(ffi.FFT)
    e1.record(stream)
    stream.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out["fused_peer_memory"] = {"ms": float(ms.item()), "elements_per_s": n / (float(ms.item()) * 1e-3),
                                "parity": "fft, coset_ifft and a chained ifft -> coset_fft equal the NCCL path, every limb"}
except Exception as ex:   # symmetric memory unavailable on this box / build
    out["fused_peer_memory"] = {"error": repr(ex)[:300]}
if rank == 0:
    print(json.dumps(out), flush=True)
dom.close()
ctx.close()
dist.destroy_process_group()
