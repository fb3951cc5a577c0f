#!/usr/bin/env python3
"""Headline benchmark: MNT4-753 G1 VariableBaseMSM::multi_scalar_mul at 2^22 points (BASELINE.json
metric / config 3), plus the 2^22 radix-2 FFT as a secondary figure.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference algorithm

One process per GPU (torchrun sets RANK / LOCAL_RANK / WORLD_SIZE); rank 0 prints ONE JSON line.
A "step" is one full MSM over the synthetic key.  With N > 1 the 2^22 points are sharded by point
range (strong scaling, as config 3 states), every rank runs a complete local MSM and the N partial
points are all-gathered over NCCL and folded on the device.

Timed regions
  value  : K steps with bases AND scalars resident in HBM (g753_msm_dev), CUDA events on the
           context's stream, barrier + synchronize on both sides, max over ranks.
  e2e    : K steps through the host-buffer C-ABI call the reference-facing shim makes
           (g753_msm: scalars in pinned host memory -> H2D -> MSM -> D2H of the 288-byte result),
           bases resident as a proving key is (uploaded once per key).
Inputs (1.1 GiB at N=1) are far larger than the 126 MB L2, so no explicit flush is needed between steps.
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GROUP = 0                     # G753_MNT4_G1
LIMB_MACS_PER_MUL = 1176      # 2 * 24^2 + 24
SEED = 0x5EED0001


def ref_window_params(n):
    """c and number of windows of the reference's Pippenger (variable_base.rs:14-25)"""
    import math
    c = 3 if n < 32 else int(math.ceil(math.log2(n) * 2.0 / 3.0 + 2.0))
    return c, (753 + c - 1) // c


def canonical_field_muls(n):
    """the REFERENCE algorithm's field-multiplication count for an n-point G1 MSM
    (mixed add 11, full add 14): n*W*11 + W*(2^c - 1)*(11 + 14)"""
    c, W = ref_window_params(n)
    return n * W * 11 + W * ((1 << c) - 1) * 25, n * W * 11


def random_scalars(n, seed):
    """uniform 752-bit canonical scalars (< r, since r > 2^752), (n, 12) uint64"""
    rng = np.random.default_rng(seed)
    sc = rng.integers(0, np.iinfo(np.uint64).max, size=(n, 12), dtype=np.uint64, endpoint=True)
    sc[:, 11] &= np.uint64((1 << 48) - 1)
    return sc


def dot_mod(scalars, logs, r):
    """sum_i scalars[i] * logs[i] mod r with 16-bit-chunk integer matmuls (exact: every partial
    sum stays below 2^32 * 2^22 * ... < 2^63 for n <= 2^30 / 4 chunks)"""
    total = 0
    step = 1 << 20
    for lo in range(0, scalars.shape[0], step):
        s16 = scalars[lo:lo + step].view(np.uint16).astype(np.uint64)          # (m, 48)
        a16 = logs[lo:lo + step].reshape(-1, 1).view(np.uint16).astype(np.uint64)  # (m, 4)
        m = a16.T @ s16                                                          # (4, 48), < 2^52
        for p in range(4):
            for q in range(48):
                total += int(m[p, q]) << (16 * (p + q))
    return total % r


def limbs_to_int(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1)
    return sum(int(v) << (64 * i) for i, v in enumerate(a))


def int_to_limbs(v, n=12):
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)], dtype=np.uint64)


def affine_of(xyz, p):
    """(X:Y:Z) homogeneous, Montgomery limbs -> (x, y) as ratios (independent of the Montgomery
    factor), None for infinity"""
    X, Y, Z = (limbs_to_int(xyz[i]) for i in range(3))
    if Z == 0:
        return None
    zi = pow(Z, -1, p)
    return (X * zi % p, Y * zi % p)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm: the CPU restatement of the reference's algorithm (oracle/ref753.cpp), all host threads
# ---------------------------------------------------------------------------------------------
def cpu_msm_rate(log_n, repeats, warmup, coords=None, scalars=None):
    """times oracle.ref753.msm (variable_base.rs:10-83 restated, one task per window) on 2^log_n
    points; returns (Mpts/s best, per-step seconds list, threads, result)"""
    from oracle import ref753                      # the one place bench.py executes oracle/ code
    params = importlib.import_module("ginger-lib_b200.params")
    n = 1 << log_n
    if coords is None:
        g = np.stack([int_to_limbs(v) for v in params.GENERATOR_MONT[GROUP]])
        coords = ref753.walk(GROUP, g, g, n)       # P_i = (i + 1) * G, affine
        scalars = random_scalars(n, SEED + 7)
    threads = ref753.hardware_threads()
    times, out = [], None
    for it in range(warmup + repeats):
        t0 = time.perf_counter()
        out = ref753.msm(GROUP, coords, None, scalars, threads)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return n / min(times) / 1e6, times, threads, out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    log_n = args.cpu_log_n
    rate, times, threads, _ = cpu_msm_rate(log_n, args.steps, args.warmup)
    n = 1 << log_n
    mean = float(np.mean(times))
    value = n / mean / 1e6
    line = {
        "impl": "reference", "metric": "mnt4753_g1_msm_throughput", "value": value, "unit": "Mpts/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64x12 (753-bit Montgomery)",
        "data": "synthetic", "gpu_launches": 0,
        "config": {"workload": "MNT4-753 G1 VariableBaseMSM::multi_scalar_mul, 2^22 points (BASELINE config 3)",
                   "sample": "2^%d points per step (bounded CPU sample of the same workload)" % log_n},
        "cpu_baseline": {"value": value, "unit": "Mpts/s", "cores": threads, "kind": "port",
                         "sample": "2^%d-point MNT4-753 G1 MSM, C++ restatement of variable_base.rs:10-83 "
                                   "(one task per window), %d threads" % (log_n, threads)},
        "e2e": {"value": value, "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    G = importlib.import_module("ginger-lib_b200")
    ffi = G.ffi
    params = importlib.import_module("ginger-lib_b200.params")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    ctx = G.Context(local_rank, stream=stream.cuda_stream)
    lib = ctx.lib

    n_total = 1 << args.log_n
    n_local = n_total // world if args.scaling == "strong" else n_total
    n_global = n_local * world
    r = params.GROUP_ORDER[GROUP]
    p = params.GROUP_BASE_MODULUS[GROUP]
    seed = SEED + 0x1000 * rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic key and scalars ---------------------------------------------------------
    t_setup = time.perf_counter()
    bases = ctx.generate_bases(GROUP, n_local, seed)           # bases[i] = a_i * G, resident
    t_pre = time.perf_counter()
    if args.copies != 1:
        bases.precompute(args.copies)                          # once per key: shifted copies 2^(j*rows*c) * P_i
        ctx.sync()
    precompute_s = time.perf_counter() - t_pre
    sc_np = random_scalars(n_local, seed + 1)
    sc_pinned = torch.from_numpy(sc_np.view(np.int64)).pin_memory()
    sc_host = sc_pinned.numpy().view(np.uint64)
    with torch.cuda.stream(stream):
        d_scalars = sc_pinned.to("cuda", non_blocking=True)
        d_out = torch.zeros(36, dtype=torch.int64, device="cuda")          # X, Y, Z
        d_gather = torch.zeros(36 * world, dtype=torch.int64, device="cuda")
        d_final = torch.zeros(36, dtype=torch.int64, device="cuda")
    stream.synchronize()
    out_host = np.zeros((3, 12), dtype=np.uint64)

    def step_resident():
        """one MSM, inputs resident; N > 1: all-gather of the partial points + device fold"""
        lib.check(lib.msm_dev(ctx.handle, bases.handle, 0, n_local, ctypes.c_void_p(d_scalars.data_ptr()),
                              ctypes.c_void_p(d_out.data_ptr())))
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(d_gather, d_out)
            lib.check(lib.points_sum_dev(ctx.handle, GROUP, ctypes.c_void_p(d_gather.data_ptr()), world,
                                         ctypes.c_void_p(d_final.data_ptr())))

    def step_e2e():
        """the reference-facing call: host scalars in, host point out"""
        if world == 1:
            lib.check(lib.msm(ctx.handle, bases.handle, 0, n_local, ffi.ptr(sc_host), ffi.ptr(out_host)))
            return
        lib.check(lib.h2d(ctx.handle, ctypes.c_void_p(d_scalars.data_ptr()), ffi.ptr(sc_host), n_local * 96))
        step_resident()
        lib.check(lib.d2h(ctx.handle, ffi.ptr(out_host), ctypes.c_void_p(d_final.data_ptr()), 288))

    # ---- correctness first: (sum s_i a_i mod r) * G, checked at full size ------------------
    launches0 = ctx.launches
    step_resident()
    stream.synchronize()
    launches_per_step = ctx.launches - launches0
    res = (d_final if world > 1 else d_out).cpu().numpy().view(np.uint64).reshape(3, 12)
    k_local = dot_mod(sc_np, G.Bases.generated_logs(n_local, seed), r)
    if world > 1:
        ks = [None] * world
        dist.all_gather_object(ks, k_local)
        k_total = sum(ks) % r
    else:
        k_total = k_local
    gen = np.stack([int_to_limbs(v) for v in params.GENERATOR_MONT[GROUP]])
    expect = np.zeros((3, 12), dtype=np.uint64)
    lib.check(lib.point_op(ctx.handle, GROUP, 2, ffi.ptr(gen), ffi.ptr(int_to_limbs(k_total)), ffi.ptr(expect)))
    ok = affine_of(res, p) == affine_of(expect, p)
    if not ok:
        raise SystemExit("rank %d: MSM result differs from (sum s_i a_i) * G - refusing to report a number" % rank)
    setup_s = time.perf_counter() - t_setup

    # ---- integer-MAC roofline probe (live) -------------------------------------------------
    ms = ctypes.c_float(0)
    lib.check(lib.mac_probe(ctx.handle, 0, 592, 256, 200, ctypes.byref(ms)))
    lib.check(lib.mac_probe(ctx.handle, 0, 592, 256, 2000, ctypes.byref(ms)))
    peak_mul_per_s = 592 * 256 * 2000 / (ms.value * 1e-3)
    peak_mac_per_s = peak_mul_per_s * LIMB_MACS_PER_MUL

    # ---- timed: resident inputs -------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launches
    ev0.record(stream)
    for _ in range(args.steps):
        step_resident()
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    gpu_launches = ctx.launches - launches0
    ms_total = ev0.elapsed_time(ev1)
    phases = ctx.last_msm_phases()            # CUDA events of the last timed step, on the same stream
    try:
        plan = ctx.last_msm_plan()            # window bits / windows / bucket rows / key copies of that step
    except Exception:                         # introspection only: never fail the measurement over it
        plan = None
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = n_global / (ms_per_step * 1e-3) / 1e6

    # ---- timed: end to end through the host-buffer call -----------------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_global / (float(t.item()) / args.steps) / 1e6
    res2 = out_host.copy()
    if affine_of(res2, p) != affine_of(expect, p):
        raise SystemExit("rank %d: e2e result differs from the resident-path result" % rank)

    # ---- secondary: 2^log_n radix-2 FFT on mnt4753::Fr (rank 0 figure, per GPU) -------------
    fft = None
    if not args.no_fft:
        nf = 1 << args.fft_log_n
        raw = random_scalars(nf, 99)
        raw[:, 11] &= np.uint64(0xFFFF)        # < p: a valid Montgomery representation
        vec = G.DeviceVector(ctx, ffi.FIELD_MNT4_FR, nf, raw)
        for _ in range(3):
            vec.ntt(ffi.FFT)
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(args.steps, 5)
        e0.record(stream)
        for _ in range(reps):
            vec.ntt(ffi.FFT)
        e1.record(stream)
        stream.synchronize()
        fft_ms = e0.elapsed_time(e1) / reps
        hbm_peak = 6543.4
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            peak_src = "measured (MEASURED_PEAKS.json)"
        except (OSError, KeyError, ValueError):
            hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        muls = (nf // 2) * args.fft_log_n
        fft = {"metric": "mnt4753_fr_fft_throughput", "log_n": args.fft_log_n, "value": nf / (fft_ms * 1e-3),
               "unit": "elements/s", "ms": fft_ms,
               "roofline_hbm": {"bound": "hbm", "achieved": 192.0 * nf / (fft_ms * 1e-3) / 1e9, "peak": hbm_peak,
                                "unit": "GB/s", "frac": 192.0 * nf / (fft_ms * 1e-3) / 1e9 / hbm_peak,
                                "peak_source": peak_src, "algorithmic_bytes": 192 * nf},
               "roofline_int": {"bound": "int32-mac", "achieved": muls * LIMB_MACS_PER_MUL / (fft_ms * 1e-3) / 1e12,
                                "peak": peak_mac_per_s / 1e12, "unit": "Tlimb-MAC/s",
                                "frac": muls * LIMB_MACS_PER_MUL / (fft_ms * 1e-3) / peak_mac_per_s}}
        vec.free()
        if world > 1:
            fft["sharded"] = sharded_fft(ctx, stream, ffi, args.fft_log_n, world, max(args.steps, 5))

    # ---- CPU baseline beside it (rank 0, N = 1): same bases / scalars, bounded sample ------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        m = 1 << args.cpu_log_n
        m = min(m, n_local)
        coords = bases.download(0, m)
        rate, times, threads, cpu_out = cpu_msm_rate(args.cpu_log_n, 1, 0, coords, sc_np[:m])
        got = G.VariableBaseMSM.multi_scalar_mul(bases, sc_np[:m])
        if affine_of(got, p) != affine_of(cpu_out, p):
            raise SystemExit("CUDA MSM and the CPU restatement disagree on the %d-point sample" % m)
        cpu = {"value": rate, "unit": "Mpts/s", "cores": threads, "kind": "port",
               "sample": "first 2^%d of the benchmark's own bases/scalars, C++ restatement of "
                         "variable_base.rs:10-83, 1 run of %.1f s; result equal to the CUDA path's" %
                         (args.cpu_log_n, times[0])}

    # ---- config 4: MNT6-753 G2 (Fq3) MSM at 2^20 + a mixed-radix transform (rank 0, N = 1) ----
    cfg4 = None
    if rank == 0 and world == 1 and not args.no_config4:
        bases.free()
        cfg4 = run_config4(ctx, G, ffi, params, args)

    # ---- config 5: end-to-end Groth16 proof; N > 1: the long MSMs sharded by point range -------
    g16 = None
    if not args.no_groth16:
        import bench_groth16
        bases.free()
        g16 = bench_groth16.run(ctx, args.groth16_log_n, steps=max(2, min(args.steps, 3)), warmup=1, copies=args.copies,
                                rank=rank, world=world, barrier=barrier if world > 1 else None)

    if rank == 0:
        canon_all, canon_acc = canonical_field_muls(n_local)
        acc_ms = phases.get("accumulate", 0.0)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_bucket_acc_bytes_2p22")
        except (OSError, ValueError):
            pass
        roofline = {
            "bound": "int32-mac",
            "kernel": "k_bucket_acc",
            "achieved": canon_acc * LIMB_MACS_PER_MUL / (acc_ms * 1e-3) / 1e12 if acc_ms else None,
            "peak": peak_mac_per_s / 1e12,
            "unit": "Tlimb-MAC/s",
            "frac": canon_acc * LIMB_MACS_PER_MUL / (acc_ms * 1e-3) / peak_mac_per_s if acc_ms else None,
            "traffic": traffic,
            "peak_source": "measured live: dependent fq_mul stream on 592x256 threads (g753_mac_probe), 1176 limb-MACs per product",
            "algorithmic": "reference op count of the bucket accumulation: n*W*11 field muls * 1176 limb-MACs "
                           "(SURVEY.md 8d), n=%d per GPU" % n_local,
            "kernel_ms": acc_ms,
            # what the kernel actually executes: n * W mixed additions (XYZZ, 8 products + 2 squarings
            # of 1176 / 876 limb-MACs) - the fraction of the integer pipe's measured ceiling it sustains
            "executed": None if not plan else {
                "plan": plan, "mixed_additions": n_local * plan["windows"],
                "limb_macs": n_local * plan["windows"] * (8 * 1176 + 2 * 876),
                "frac": n_local * plan["windows"] * (8 * 1176 + 2 * 876) / (acc_ms * 1e-3) / peak_mac_per_s
                if acc_ms else None},
            "whole_msm_frac": canon_all * LIMB_MACS_PER_MUL / (ms_per_step * 1e-3) / peak_mac_per_s,
            "phases_ms": phases,
        }
        line = {
            "metric": "mnt4753_g1_msm_throughput", "value": value, "unit": "Mpts/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u32x24 (753-bit Montgomery, integer)", "data": "synthetic",
            "config": {"workload": "MNT4-753 G1 VariableBaseMSM::multi_scalar_mul, 2^%d points (BASELINE config 3)"
                                   % args.log_n,
                       "points_total": n_global, "points_per_gpu": n_local,
                       "bases": "a_i*G, a_i = splitmix64 (g753_bases_generate), resident in HBM" +
                                (" with precomputed shifted copies (g753_bases_precompute, %s, %.1f s once per key)"
                                 % ("%d copies" % args.copies if args.copies else "as many as fit 6 GiB", precompute_s)
                                 if args.copies != 1 else ""),
                       "scalars": "uniform 752-bit canonical", "l2": "inputs (%.0f MiB/GPU) larger than L2" %
                       ((n_local * 288) / 2**20),
                       "parallelism": "point-range shards x%d, NCCL all-gather of partial points + device fold" % world
                       if world > 1 else "single GPU",
                       "verified": "result == (sum s_i a_i mod r)*G at full size"},
            "e2e": {"value": e2e_value, "unit": "Mpts/s", "h2d_bytes_per_step": n_local * 96,
                    "d2h_bytes_per_step": 288,
                    "path": "g753_msm: pinned host scalars -> H2D -> MSM -> D2H result; bases resident (proving key)"},
            "gpu_launches": gpu_launches, "launches_per_step": launches_per_step,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "fft": fft, "config4": cfg4, "groth16": g16,
            "setup_s": setup_s, "key_precompute_s": precompute_s,
        }
        print(json.dumps(line), flush=True)
    bases.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def sharded_fft(ctx, stream, ffi, log_n, world, reps):
    """the 2^log_n transform sharded over all ranks (four-step): exchange fused into the
    last butterfly pass over peer memory when symmetric memory is available, else NCCL all-to-all.
    Device-timed, max over ranks."""
    import torch
    import torch.distributed as dist
    D = importlib.import_module("ginger-lib_b200.distributed")
    field = ffi.FIELD_MNT4_FR
    n = 1 << log_n
    dom = D.ShardedEvaluationDomain(ctx, field, log_n)
    raw = random_scalars(dom.local, 123)
    raw[:, 11] &= np.uint64(0xFFFF)
    dev = torch.device("cuda", ctx.device)
    path = "fused: last butterfly pass stores into peer memory (symmetric memory over NVLink), one barrier"
    try:
        # measured (profiles/r01_ntt_sharded_fused_n{2,4,8}.json): the fused exchange wins on 2 GPUs (3.36 vs
        # 3.79 ms), ties on 4 and loses 5 % on 8, where its 96-byte remote stores cost more than NCCL's bulk
        # copies save
        if world > 4:
            raise RuntimeError("NCCL preferred above 4 ranks")
        fused = D.FusedShardedNTT(dom, stream)
        fused.load(raw)
        one = lambda: fused.transform(ffi.FFT)
    except Exception as ex:                                   # no symmetric memory on this box / build
        path = "NCCL all_to_all_single between the column and row passes (%s)" % str(ex)[:60]
        with torch.cuda.stream(stream):
            t_data = torch.from_numpy(raw.view(np.int64).reshape(-1).copy()).to(dev)
            t_send, t_recv = torch.empty_like(t_data), torch.empty_like(t_data)
        stream.synchronize()
        p = lambda t: ctypes.c_void_p(t.data_ptr())

        def exchange():
            with torch.cuda.stream(stream):
                dist.all_to_all_single(t_recv, t_send)
        one = lambda: dom.transform_dev(p(t_data), p(t_send), p(t_recv), ffi.FFT, exchange)
    for _ in range(3):
        one()
    stream.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        one()
    e1.record(stream)
    stream.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    return {"n_gpus": world, "log_n": log_n, "ms": ms, "value": n / (ms * 1e-3), "unit": "elements/s", "exchange": path,
            "exchange_bytes_per_rank": (n // world) * 96 * (world - 1) // world}


def run_config4(ctx, G, ffi, params, args):
    """BASELINE config 4: MNT6-753 G2 (over Fq3) MSM at 2^20 points on a resident synthetic key, verified
    by its discrete logs, and the mixed-radix transform at 2^15 * 25 = 819 200 points on mnt6753::Fr
    (round trip checked; parity unpinned - the reference has no mixed-radix domain)"""
    import torch
    group, log_n = ffi.MNT6_G2, 20
    n = 1 << log_n
    t0 = time.perf_counter()
    bases = ctx.generate_bases(group, n, 0xC4)
    bases.precompute(args.copies)
    ctx.sync()
    key_s = time.perf_counter() - t0
    sc = random_scalars(n, 0xC5)
    r = params.GROUP_ORDER[group]
    p = params.GROUP_BASE_MODULUS[group]
    out = None
    times = []
    for it in range(3):
        t1 = time.perf_counter()
        out = G.VariableBaseMSM.multi_scalar_mul(bases, sc)        # host scalars in, host point out
        times.append(time.perf_counter() - t1)
    phases = ctx.last_msm_phases()
    k = dot_mod(sc, G.Bases.generated_logs(n, 0xC4), r)
    gen = np.stack([int_to_limbs(v) for v in params.GENERATOR_MONT[group]]).reshape(-1)
    expect = np.zeros((3, 36), dtype=np.uint64)
    ctx.lib.check(ctx.lib.point_op(ctx.handle, group, 2, ffi.ptr(gen), ffi.ptr(int_to_limbs(k)), ffi.ptr(expect)))
    xy = np.zeros((2, 72), dtype=np.uint64)
    inf = np.zeros(2, dtype=np.uint8)
    both = np.ascontiguousarray(np.stack([out.reshape(-1), expect.reshape(-1)]))
    ctx.lib.check(ctx.lib.batch_normalize(ctx.handle, group, ffi.ptr(both), 2, ffi.ptr(xy), ffi.ptr(inf)))
    if not (xy[0] == xy[1]).all():
        raise SystemExit("config 4: MNT6 G2 MSM differs from (sum s_i a_i) * G")
    bases.free()
    # mixed-radix transform
    N = (1 << 15) * 25
    field = ffi.FIELD_MNT6_FR
    raw = random_scalars(N, 0xC6)
    raw[:, 11] &= np.uint64(0xFFFF)
    vec = G.DeviceVector(ctx, field, N, raw)
    lib = ctx.lib
    for mode in (ffi.FFT, ffi.IFFT):
        lib.check(lib.ntt_mixed_dev(ctx.handle, field, vec.ptr, N, mode))
    if not (vec.download() == raw).all():
        raise SystemExit("config 4: mixed-radix ifft(fft(x)) != x")
    ctx.sync()
    t1 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        lib.check(lib.ntt_mixed_dev(ctx.handle, field, vec.ptr, N, ffi.FFT))
    ctx.sync()
    mixed_ms = (time.perf_counter() - t1) / reps * 1e3
    vec.free()
    return {"msm": {"workload": "MNT6-753 G2 (Fq3) MSM, 2^20 points, resident key with precomputed copies (built in %.1f s)"
                                % key_s,
                    "ms": min(times) * 1e3, "mpts_per_s": n / min(times) / 1e6, "phases_ms": phases,
                    "verified": "result == (sum s_i a_i mod r) * G2 generator"},
            "mixed_radix_fft": {"field": "mnt6753::Fr", "n": N, "factorisation": "2^15 * 5^2", "ms": mixed_ms,
                                "elements_per_s": N / (mixed_ms * 1e-3), "verified": "ifft(fft(x)) == x",
                                "parity": "unpinned: no mixed-radix domain in the reference"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=22, help="log2 of the total number of points")
    ap.add_argument("--fft-log-n", type=int, default=22)
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--cpu-log-n", type=int, default=20, help="log2 of the CPU baseline's bounded sample (~10 s on 16 threads)")
    ap.add_argument("--copies", type=int, default=0,
                    help="precomputed shifted copies of the resident key (0 = auto by memory budget, 1 = plain key)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-fft", action="store_true")
    ap.add_argument("--no-groth16", action="store_true")
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--groth16-log-n", type=int, default=20)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
