#!/usr/bin/env python3
"""Headline benchmark: MNT4-753 G1 VariableBaseMSM::multi_scalar_mul at 2^22 points (BASELINE.json
metric / config 3), with the 2^22 radix-2 FFT, configs 1, 2, 4 and the Groth16 proof (config 5) beside it.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference algorithm

One process per GPU (torchrun sets RANK / LOCAL_RANK / WORLD_SIZE); rank 0 prints ONE JSON line.
A "step" is one full MSM over the synthetic key.  With N > 1 the 2^22 points are sharded by point
range (strong scaling, as config 3 states), every rank runs a complete local MSM and the N partial
points are all-gathered over NCCL and folded on the device.

Timed regions
  value    : K steps with bases AND scalars resident in HBM (g753_msm_dev), CUDA events on the
             context's stream, barrier + synchronize on both sides, max over ranks.
  e2e      : K steps through the host-buffer C-ABI call the reference-facing shim makes
             (g753_msm: scalars in pinned host memory -> H2D -> MSM -> D2H of the 288-byte result),
             bases resident as a proving key is (uploaded once per key).
  e2e_cold : the literal one-shot drop-in multi_scalar_mul(&bases, &scalars) = g753_msm_host: bases AND
             scalars in (pageable) host memory, uploaded per call, plain key (no precomputed copies).
Inputs (1.1 GiB at N=1) are far larger than the 126 MB L2, so no explicit flush is needed between steps.

Every figure is checked before it is reported: the MSM against (sum s_i a_i mod r) * G at full size and,
at N = 1, against the C++ restatement of the reference on the same 2^22 bases and scalars; the FFT every
limb against the restated transform; the sharded FFT against the single-GPU one; the proof against its
discrete-log prediction.  The CPU restatement (oracle/ref753.cpp, kind "port") is the checker and the
timed CPU baseline - it is never part of the measured GPU path.
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
LIMB_MACS_PER_SQR = 876       # dedicated squaring (fq.cuh fq_sqr)
LIMB_MACS_MUL2 = 1752         # two products under one reduction (fq.cuh fq_mul2)
SEED = 0x5EED0001
METRIC = "mnt4753_g1_msm_throughput"

# limb-MACs one mixed addition of the accumulation kernel EXECUTES (ec_slots.cuh madd_g: 8 products + 2
# squarings of the coordinate field):
#   G1:        8 x 1176 + 2 x 876
#   G2 / Fq2:  10 tower products, each two lanes x one two-product body (slots.cuh Tw2C, G753_FQ2_LAZY = 2)
#   G2 / Fq3:  10 tower products, each three lanes x one three-product body (slots.cuh Tw3L: 3 x 576 + 600)
LIMB_MACS_MUL3 = 3 * 576 + 600
EXECUTED_MACS_PER_MADD = {0: 8 * 1176 + 2 * 876, 2: 8 * 1176 + 2 * 876, 1: 10 * 2 * 1752, 3: 10 * 3 * LIMB_MACS_MUL3}
# ... and one AFFINE addition of the pairwise tree (msm.cuh k_tree_round): 3 products for its share of the
# shared inversion (prefix product, 1 / den, peel), lambda, lambda^2, y3 = 5 products + 1 squaring.  The one
# safegcd inversion per thread and round (shifts and adds, no limb products) is not counted.
EXECUTED_MACS_PER_TREE_ADD = {0: 5 * 1176 + 876, 2: 5 * 1176 + 876, 1: 6 * 2 * 1752, 3: 6 * 3 * LIMB_MACS_MUL3}


def executed_accumulation(group, n, plan):
    """(additions, limb-MACs, description) the accumulation phase of an n-point MSM EXECUTES under `plan`"""
    entries = n * plan["windows"]
    if plan.get("accumulation") == "affine_tree":
        # every bucket's run of k entries takes k - 1 additions (all buckets are occupied at these sizes)
        adds = max(entries - plan["rows"] * (1 << (plan["c"] - 1)), 0)
        return adds, adds * EXECUTED_MACS_PER_TREE_ADD[group], "affine additions, 5 products + 1 squaring each"
    return entries, entries * EXECUTED_MACS_PER_MADD[group], "XYZZ mixed additions, 8 products + 2 squarings each"


def workload_config(log_n, world, scaling):
    """the workload, as both arms of the comparison name it (nothing implementation-specific)"""
    n = 1 << log_n
    n_local = n // world if scaling == "strong" else n
    return {"workload": "MNT4-753 G1 VariableBaseMSM::multi_scalar_mul, 2^%d points (BASELINE config 3)" % log_n,
            "points_total": n_local * world, "points_per_step": n_local * world,
            "scalars": "uniform 752-bit canonical",
            "bases": "synthetic affine points of the prime-order subgroup",
            "l2": "inputs (%.0f MiB per step and GPU) larger than the 126 MB L2: no flush between steps" %
                  ((n_local * 288) / 2**20)}


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


def random_field_elements(n, seed):
    """valid Montgomery representations (< 2^752 < p), (n, 12) uint64"""
    raw = random_scalars(n, seed)
    raw[:, 11] &= np.uint64(0xFFFF)
    return raw


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


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


def library_provenance(lib):
    """which sources the loaded libg753.so was built from, against this tree's"""
    build = importlib.import_module("ginger-lib_b200.build")
    loaded, tree = lib.source_hash().decode(), build.source_hash()
    return {"version": lib.version().decode(), "source_hash": loaded, "tree_hash": tree, "matches_tree": loaded == tree}


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
# CPU legs: the C++ restatement of the reference's algorithms (oracle/ref753.cpp) on all host threads.
# These functions are the only places bench.py executes oracle/ code.
# ---------------------------------------------------------------------------------------------
def cpu_walk_bases(group, n):
    """P_i = (i + 1) * G, affine Montgomery limbs (oracle-side generator for the CPU arm)"""
    from oracle import ref753
    params = importlib.import_module("ginger-lib_b200.params")
    k = ref753.GROUP_K[group]
    g = np.stack([int_to_limbs(v) for v in params.GENERATOR_MONT[group]]).reshape(2, k * 12)
    return ref753.walk(group, g, g, n)


def cpu_msm_full(group, coords, scalars):
    """one whole multi_scalar_mul (variable_base.rs:10-83 restated, one task per window); (seconds, result, threads)"""
    from oracle import ref753
    threads = ref753.hardware_threads()
    t0 = time.perf_counter()
    out = ref753.msm(group, coords, None, scalars, threads)
    return time.perf_counter() - t0, out, threads


def cpu_fft(field, raw, mode):
    from oracle import ref753
    threads = ref753.hardware_threads()
    t0 = time.perf_counter()
    out = ref753.fft(field, raw, mode, threads)
    return time.perf_counter() - t0, out, threads


def cpu_groth16(log_n):
    """create_proof's hot path (prover.rs:241-337 + r1cs_to_qap.rs:121-166) restated on the CPU port at a
    2^log_n domain: seven transforms with the element-wise steps, into_repr, five long MSMs.  Wall seconds
    per phase."""
    from oracle import ref753
    ffi = importlib.import_module("ginger-lib_b200").ffi
    field, g1, g2 = ffi.FIELD_MNT4_FR, ffi.MNT4_G1, ffi.MNT4_G2
    n = 1 << log_n
    threads = ref753.hardware_threads()
    a, b, c = (random_field_elements(n, 0x91 + i) for i in range(3))
    z = random_field_elements(n, 0x94)
    q_g1 = cpu_walk_bases(g1, n)
    q_g2 = cpu_walk_bases(g2, n)
    zinv = random_field_elements(1, 0x95)
    t = {}
    t0 = time.perf_counter()
    fa = ref753.fft(field, ref753.fft(field, a, ffi.IFFT, threads), ffi.COSET_FFT, threads)
    fb = ref753.fft(field, ref753.fft(field, b, ffi.IFFT, threads), ffi.COSET_FFT, threads)
    fc = ref753.fft(field, ref753.fft(field, c, ffi.IFFT, threads), ffi.COSET_FFT, threads)
    ab = ref753.field_op(field, 2, ref753.field_op(field, 0, fa, fb), fc)
    ab = ref753.field_op(field, 0, ab, np.tile(zinv, (n, 1)))
    h = ref753.fft(field, ab, ffi.COSET_IFFT, threads)
    t["witness_map"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    hr = ref753.field_op(field, 7, h)            # into_repr (prover.rs:241-267)
    zr = ref753.field_op(field, 7, z)
    t["into_repr"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(3):                            # A, B1, L over the assignment
        ref753.msm(g1, q_g1, None, zr, threads)
    ref753.msm(g1, q_g1, None, hr, threads)       # H
    t["msm_g1_x4"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref753.msm(g2, q_g2, None, zr, threads)       # B2
    t["msm_g2"] = time.perf_counter() - t0
    return t, threads


def run_reference(args):
    """the reference arm: the restated CPU algorithm on the SAME workload (2^log_n points, same window
    size c and window count as the reference picks for that input).  A step is a bounded sample of one
    MSM: `threads` of its ceil(753 / c) independent per-window tasks (variable_base.rs:30-70: one rayon
    task per window, every window the same work: n mixed additions into 2^c - 1 buckets + the
    running-sum reduction), i.e. one full wave of the thread pool; value = points x (windows timed /
    windows) / step time.  The Horner fold of the window sums (:72-82, ~750 doublings, < 1 ms) is not in
    the sample.  Warm-up steps run the same tasks on a 2^16-point prefix."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref753
    log_n = args.log_n
    n = 1 << log_n
    threads = ref753.hardware_threads()
    coords = cpu_walk_bases(GROUP, n)
    scalars = random_scalars(n, SEED + 7)
    c, W = ref_window_params(n)
    assert c == ref753.msm_window_bits(n)
    per_step = min(W, threads)
    small = min(n, 1 << 16)
    for _ in range(args.warmup):
        ref753.msm_windows(GROUP, coords[:small], None, scalars[:small], 0, min(per_step, ref_window_params(small)[1]), threads)
    times = []
    for it in range(args.steps):
        first = (it * per_step) % max(1, W - per_step)      # the last (partial) window is never the only one
        t0 = time.perf_counter()
        ref753.msm_windows(GROUP, coords, None, scalars, first, per_step, threads)
        times.append(time.perf_counter() - t0)
    mean = float(np.mean(times))
    value = n * (per_step / W) / mean / 1e6
    sample = ("%d of the %d per-window tasks (c = %d) of the 2^%d-point MSM per step = one wave of the %d-thread pool, "
              "C++ restatement of variable_base.rs:10-83; full-MSM equivalent %.1f s"
              % (per_step, W, c, log_n, threads, mean * W / per_step))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpts/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u64x12 (753-bit Montgomery)",
        "data": "synthetic", "gpu_launches": 0,
        "config": workload_config(log_n, args.gpus, args.scaling),
        "sample": sample,
        "cpu_baseline": {"value": value, "unit": "Mpts/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------
def device_time(stream, fn, reps, warm=1):
    """mean ms of fn() over reps, CUDA events on `stream`"""
    import torch
    for _ in range(warm):
        fn()
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    stream.synchronize()
    return e0.elapsed_time(e1) / reps


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
    provenance = library_provenance(lib)
    if not provenance["matches_tree"]:
        raise SystemExit("libg753.so was built from other sources (%s) than this tree (%s): rebuild with "
                         "__graft_entry__.build()" % (provenance["source_hash"], provenance["tree_hash"]))

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
    gen = np.stack([int_to_limbs(v) for v in params.GENERATOR_MONT[GROUP]])

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

    # the discrete-log prediction of the result: (sum s_i a_i mod r) * G
    k_local = dot_mod(sc_np, G.Bases.generated_logs(n_local, seed), r)
    if world > 1:
        ks = [None] * world
        dist.all_gather_object(ks, k_local)
        k_total = sum(ks) % r
    else:
        k_total = k_local
    expect = np.zeros((3, 12), dtype=np.uint64)
    lib.check(lib.point_op(ctx.handle, GROUP, 2, ffi.ptr(gen), ffi.ptr(int_to_limbs(k_total)), ffi.ptr(expect)))
    expect_affine = affine_of(expect, p)

    def check_result(what):
        res = (d_final if world > 1 else d_out).cpu().numpy().view(np.uint64).reshape(3, 12)
        if affine_of(res, p) != expect_affine:
            raise SystemExit("rank %d: %s differs from (sum s_i a_i) * G - refusing to report a number" % (rank, what))

    # ---- the plain key first: correctness + its timing (the cold / one-shot path uses it) ---
    step_resident()
    stream.synchronize()
    check_result("MSM on the plain key")
    plain_ms = device_time(stream, step_resident, 2, warm=1)
    plain_phases = ctx.last_msm_phases()
    plain_plan = ctx.last_msm_plan()

    # ---- CPU baseline and full-size parity (rank 0, N = 1): the restated reference on the SAME 2^22 bases
    # and scalars, one whole multi_scalar_mul; its result must be the CUDA path's -------------------
    cpu = None
    cold = None
    if rank == 0 and world == 1 and not args.no_cpu:
        coords = bases.download()
        cpu_s, cpu_out, threads = cpu_msm_full(GROUP, coords, sc_np)
        if affine_of(cpu_out, p) != expect_affine:
            raise SystemExit("the CPU restatement and the CUDA MSM disagree at full size")
        c_ref, w_ref = ref_window_params(n_local)
        cpu = {"value": n_local / cpu_s / 1e6, "unit": "Mpts/s", "cores": threads, "kind": "port",
               "sample": "one whole 2^%d-point MSM on the benchmark's own bases and scalars (c = %d, %d window tasks), "
                         "C++ restatement of variable_base.rs:10-83, %d threads, %.1f s; its result equals the CUDA "
                         "path's (affine, every limb)" % (args.log_n, c_ref, w_ref, threads, cpu_s)}
        # the literal one-shot drop-in: host bases + host scalars per call (pageable memory, plain key)
        cold_out = np.zeros((3, 12), dtype=np.uint64)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            lib.check(lib.msm_host(ctx.handle, GROUP, ffi.ptr(coords), None, n_local, ffi.ptr(sc_np), n_local, ffi.ptr(cold_out)))
            ts.append(time.perf_counter() - t0)
        if affine_of(cold_out, p) != expect_affine:
            raise SystemExit("g753_msm_host result differs")
        cold = {"value": n_local / min(ts) / 1e6, "unit": "Mpts/s", "ms": min(ts) * 1e3,
                "h2d_bytes_per_step": n_local * (192 + 96), "d2h_bytes_per_step": 288,
                "path": "g753_msm_host: multi_scalar_mul(&bases, &scalars) with bases AND scalars in pageable host memory, "
                        "uploaded per call; plain key, no precomputed copies",
                "device_ms_plain_key": plain_ms, "phases_ms_plain_key": plain_phases, "plan_plain_key": plain_plan}
        del coords

    # ---- resident key with precomputed copies (once per key) ---------------------------------------
    t_pre = time.perf_counter()
    if args.copies != 1:
        bases.precompute(args.copies)                          # shifted copies 2^(j*rows*c) * P_i
        ctx.sync()
    precompute_s = time.perf_counter() - t_pre
    launches0 = ctx.launches
    step_resident()
    stream.synchronize()
    launches_per_step = ctx.launches - launches0
    check_result("MSM on the key with precomputed copies")
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
    check_result("MSM after the timed steps")

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
    if affine_of(out_host.copy(), p) != expect_affine:
        raise SystemExit("rank %d: e2e result differs from the resident-path result" % rank)
    bases.free()
    del d_scalars

    # ---- the 2^log_n FFT, configs 1 / 2 / 4, Groth16 --------------------------------------------------
    fft = None
    if not args.no_fft:
        fft = bench_fft(ctx, stream, G, args, rank, world, peak_mac_per_s)
    cfg1 = cfg2 = cfg4 = None
    if rank == 0 and world == 1 and not args.no_configs:
        cfg1 = run_config1(ctx, stream, G, params, args)
        cfg2 = run_config2(ctx, stream, G, args)
        cfg4 = run_config4(ctx, stream, G, ffi, params, args, peak_mac_per_s)
    g16 = None
    if not args.no_groth16:
        import bench_groth16
        g16 = bench_groth16.run(ctx, args.groth16_log_n, steps=max(2, min(args.steps, 3)), warmup=1, copies=args.copies,
                                rank=rank, world=world, barrier=barrier if world > 1 else None,
                                peak_mac_per_s=peak_mac_per_s)
        if rank == 0 and world == 1 and not args.no_cpu:
            tcpu, threads = cpu_groth16(args.cpu_groth16_log_n)
            total = sum(tcpu.values())
            scale = 1 << (args.groth16_log_n - args.cpu_groth16_log_n)
            g16["cpu_baseline"] = {
                "value": total * 1e3, "unit": "ms", "cores": threads, "kind": "port",
                "sample": "the prover's hot path restated on the CPU port at a 2^%d domain (7 transforms + element-wise "
                          "steps, into_repr, A / B1 / L / H in G1 and B2 in G2 through the restated Pippenger); the 2^%d "
                          "proof is >= %d x this (MSM cost per point falls slowly with n)"
                          % (args.cpu_groth16_log_n, args.groth16_log_n, scale),
                "phases_s": tcpu, "extrapolated_ms_at_bench_size": total * 1e3 * scale}

    if rank == 0:
        canon_all, canon_acc = canonical_field_muls(n_local)
        acc_ms = phases.get("accumulate", 0.0)
        executed = None
        if plan and acc_ms:
            adds, macs, what = executed_accumulation(GROUP, n_local, plan)
            executed = {"plan": plan, "additions": adds, "what": what, "limb_macs": macs,
                        "frac": macs / (acc_ms * 1e-3) / peak_mac_per_s}
        traffic = None
        if world == 1 and args.log_n == 22:
            try:
                tree_form = bool(plan) and plan.get("accumulation") == "affine_tree"
                traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(
                    "k_tree_round_level1_bytes_2p22" if tree_form else "k_bucket_acc_bytes_2p22")
            except (OSError, ValueError):
                pass
        roofline = {
            "bound": "int32-mac",
            "kernel": "k_tree_round" if plan and plan.get("accumulation") == "affine_tree" else "k_bucket_acc",
            # the fraction of the integer pipe's measured ceiling the accumulation sustains on the limb products it
            # EXECUTES (n x W signed-digit window entries; see executed_accumulation)
            "achieved": executed["limb_macs"] / (acc_ms * 1e-3) / 1e12 if executed else None,
            "peak": peak_mac_per_s / 1e12,
            "unit": "Tlimb-MAC/s",
            "frac": executed["frac"] if executed else None,
            "traffic": traffic,
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel at "
                              "2^22 on one GPU (profiles/; for k_tree_round: its level-1 launch, half of the additions; "
                              "kernel_ms covers all levels); not measured per run, null for other shapes",
            "peak_source": "measured live: dependent fq_mul stream on 592x256 threads (g753_mac_probe), 1176 limb-MACs per product",
            "kernel_ms": acc_ms,
            "executed": executed,
            # the same kernel time scored on the REFERENCE algorithm's operation count (n x W_ref
            # unsigned windows x 11 products): above 1 because the executed algorithm does less work
            "reference_count": {"limb_macs": canon_acc * LIMB_MACS_PER_MUL,
                                "frac": canon_acc * LIMB_MACS_PER_MUL / (acc_ms * 1e-3) / peak_mac_per_s if acc_ms else None,
                                "whole_msm_frac": canon_all * LIMB_MACS_PER_MUL / (ms_per_step * 1e-3) / peak_mac_per_s},
            "phases_ms": phases,
        }
        config = workload_config(args.log_n, world, args.scaling)
        line = {
            "metric": METRIC, "value": value, "unit": "Mpts/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u32x24 (753-bit Montgomery, integer)", "data": "synthetic",
            "config": config,
            "run": {"points_per_gpu": n_local,
                    "bases": "a_i*G, a_i = splitmix64 (g753_bases_generate), resident in HBM" +
                             (" with precomputed shifted copies (g753_bases_precompute, %s, %.1f s once per key)"
                              % ("%d copies" % args.copies if args.copies else "as many as fit 6 GiB", precompute_s)
                              if args.copies != 1 else ""),
                    "parallelism": "point-range shards x%d, NCCL all-gather of partial points + device fold" % world
                    if world > 1 else "single GPU",
                    "verified": "result == (sum s_i a_i mod r)*G at full size, plain key and precomputed key, before and "
                                "after the timed steps" + ("; == the CPU restatement's result at full size" if cpu else "")},
            "e2e": {"value": e2e_value, "unit": "Mpts/s", "h2d_bytes_per_step": n_local * 96,
                    "d2h_bytes_per_step": 288,
                    "path": "g753_msm: pinned host scalars -> H2D -> MSM -> D2H result; bases resident (proving key)"},
            "e2e_cold": cold,
            "gpu_launches": gpu_launches, "launches_per_step": launches_per_step,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "fft": fft,
            "config1": cfg1, "config2": cfg2, "config4": cfg4, "groth16": g16,
            "setup_s": setup_s, "key_precompute_s": precompute_s, "library": provenance,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def bench_fft(ctx, stream, G, args, rank, world, peak_mac_per_s):
    """the 2^fft_log_n radix-2 transform on mnt4753::Fr: verified, then device-timed (value), timed through
    the host-buffer call (e2e), timed on the CPU port (cpu_baseline); N > 1: the sharded transform as well"""
    import torch
    ffi = G.ffi
    field = ffi.FIELD_MNT4_FR
    log_n = args.fft_log_n
    nf = 1 << log_n
    raw = random_field_elements(nf, 99)
    vec = G.DeviceVector(ctx, field, nf, raw)
    vec.ntt(ffi.FFT)
    got = vec.download()
    cpu = None
    verified = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_s, want, threads = cpu_fft(field, raw, ffi.FFT)
        if not np.array_equal(got, want):
            raise SystemExit("2^%d FFT differs from the CPU restatement - refusing to report a number" % log_n)
        verified = "every limb of the 2^%d outputs equals the C++ restatement of domain.rs:120-123 / 305-416" % log_n
        cpu = {"value": nf / cpu_s, "unit": "elements/s", "cores": threads, "kind": "port",
               "sample": "one whole 2^%d fft_in_place (best_fft) on the same input, %d threads, %.2f s" % (log_n, threads, cpu_s)}
    else:
        vec.ntt(ffi.IFFT)
        if not np.array_equal(vec.download(), raw):
            raise SystemExit("2^%d FFT: ifft(fft(x)) != x" % log_n)
        verified = "ifft(fft(x)) == x at full size (the every-limb comparison with the CPU restatement runs at N = 1)"
    vec.upload(raw)
    reps = max(args.steps, 5)
    fft_ms = device_time(stream, lambda: vec.ntt(ffi.FFT), reps, warm=3)
    # end to end through g753_ntt: pinned host buffer in and out
    pinned = torch.from_numpy(raw.view(np.int64).copy()).pin_memory()
    host = pinned.numpy().view(np.uint64)
    lib = ctx.lib
    for _ in range(2):
        lib.check(lib.ntt(ctx.handle, field, ffi.ptr(host), log_n, ffi.FFT))
    t0 = time.perf_counter()
    for _ in range(reps):
        lib.check(lib.ntt(ctx.handle, field, ffi.ptr(host), log_n, ffi.FFT))
    e2e_ms = (time.perf_counter() - t0) / reps * 1e3
    peak, peak_src = hbm_peak()
    muls = (nf // 2) * log_n
    res = {"metric": "mnt4753_fr_fft_throughput", "log_n": log_n, "value": nf / (fft_ms * 1e-3),
           "unit": "elements/s", "ms": fft_ms, "verified": verified,
           "e2e": {"value": nf / (e2e_ms * 1e-3), "unit": "elements/s", "ms": e2e_ms, "h2d_bytes_per_step": nf * 96,
                   "d2h_bytes_per_step": nf * 96, "path": "g753_ntt: pinned host vector -> H2D -> transform -> D2H"},
           "cpu_baseline": cpu,
           "roofline_hbm": {"bound": "hbm", "achieved": 192.0 * nf / (fft_ms * 1e-3) / 1e9, "peak": peak,
                            "unit": "GB/s", "frac": 192.0 * nf / (fft_ms * 1e-3) / 1e9 / peak,
                            "peak_source": peak_src, "algorithmic_bytes": 192 * nf},
           "roofline_int": {"bound": "int32-mac", "achieved": muls * LIMB_MACS_PER_MUL / (fft_ms * 1e-3) / 1e12,
                            "peak": peak_mac_per_s / 1e12, "unit": "Tlimb-MAC/s",
                            "frac": muls * LIMB_MACS_PER_MUL / (fft_ms * 1e-3) / peak_mac_per_s,
                            "algorithmic": "(n / 2) log2 n butterflies x one 1176-limb-MAC product"}}
    single = got if world > 1 else None
    vec.free()
    if world > 1:
        res["sharded"] = sharded_fft(ctx, stream, G.ffi, log_n, world, reps, raw, single, rank)
    return res


def sharded_fft(ctx, stream, ffi, log_n, world, reps, raw_full, single_gpu_result, rank):
    """the 2^log_n transform sharded over all ranks (four-step): exchange fused into the
    last butterfly pass over peer memory when symmetric memory is available, else NCCL all-to-all.
    The gathered output is compared, every limb, with the single-GPU transform of the same vector;
    then device-timed, max over ranks."""
    import torch
    import torch.distributed as dist
    D = importlib.import_module("ginger-lib_b200.distributed")
    field = ffi.FIELD_MNT4_FR
    n = 1 << log_n
    dom = D.ShardedEvaluationDomain(ctx, field, log_n)
    shard = dom.scatter(raw_full)
    dev = torch.device("cuda", ctx.device)
    path = "fused: last butterfly pass stores into peer memory (symmetric memory over NVLink), one barrier"
    fused = None
    try:
        if world > args_fused_max():
            raise RuntimeError("NCCL preferred above %d ranks" % args_fused_max())
        fused = D.FusedShardedNTT(dom, stream)
        fused.load(shard)
        one = lambda: fused.transform(ffi.FFT)
        fetch = fused.store
    except Exception as ex:                                   # no symmetric memory on this box / build
        path = "NCCL all_to_all_single between the column and row passes (%s)" % str(ex)[:60]
        with torch.cuda.stream(stream):
            t_data = torch.from_numpy(shard.view(np.int64).reshape(-1).copy()).to(dev)
            t_send, t_recv = torch.empty_like(t_data), torch.empty_like(t_data)
        stream.synchronize()
        pp = lambda t: ctypes.c_void_p(t.data_ptr())

        def exchange():
            with torch.cuda.stream(stream):
                dist.all_to_all_single(t_recv, t_send)
        one = lambda: dom.transform_dev(pp(t_data), pp(t_send), pp(t_recv), ffi.FFT, exchange)

        def fetch():
            stream.synchronize()
            return t_data.cpu().numpy().view(np.uint64).reshape(dom.local, 12)
    # ---- correctness: one transform of the scattered input, gathered, against the single-GPU result ----
    one()
    full = dom.gather(fetch())
    ok = bool(np.array_equal(full, single_gpu_result))
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if not all(flags):
        raise SystemExit("sharded FFT differs from the single-GPU transform - refusing to report a number")
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
            "exchange_bytes_per_rank": (n // world) * 96 * (world - 1) // world,
            "verified": "gathered output == the single-GPU transform of the same vector, every limb"}


def args_fused_max():
    """largest world size the peer-memory exchange is used at (G753_FUSED_MAX_RANKS overrides)"""
    return int(os.environ.get("G753_FUSED_MAX_RANKS", "4"))


def run_config1(ctx, stream, G, params, args):
    """BASELINE config 1: MNT4-753 G1 multi_scalar_mul at 2^16 points ('CPU reference runs today'), checked
    against the C++ restatement; resident key with copies, resident plain key, one-shot host call"""
    ffi = G.ffi
    group, log_n = ffi.MNT4_G1, 16
    n = 1 << log_n
    p = params.GROUP_BASE_MODULUS[group]
    bases = ctx.generate_bases(group, n, 0xC1)
    coords = bases.download()
    sc = random_scalars(n, 0xC11)
    cpu_s, want, threads = cpu_msm_full(group, coords, sc) if not args.no_cpu else (None, None, None)
    lib = ctx.lib
    d_sc, d_out = G.DeviceVector(ctx, 0, n, sc), G.DeviceVector(ctx, 0, 3)
    run = lambda: lib.check(lib.msm_dev(ctx.handle, bases.handle, 0, n, d_sc.ptr, d_out.ptr))
    res = {"workload": "MNT4-753 G1 MSM, 2^16 points (BASELINE config 1)"}
    out = np.zeros((3, 12), dtype=np.uint64)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        lib.check(lib.msm_host(ctx.handle, group, ffi.ptr(coords), None, n, ffi.ptr(sc), n, ffi.ptr(out)))
        ts.append(time.perf_counter() - t0)
    if want is not None and affine_of(out, p) != affine_of(want, p):
        raise SystemExit("config 1: one-shot MSM differs from the CPU restatement")
    res["one_shot_host_ms"] = min(ts) * 1e3
    res["plain_key_device_ms"] = device_time(stream, run, 5)
    res["plain_key_phases_ms"] = ctx.last_msm_phases()
    res["plain_key_plan"] = ctx.last_msm_plan()
    bases.precompute(args.copies)
    ctx.sync()
    res["precomputed_key_device_ms"] = device_time(stream, run, 5)
    res["precomputed_key_phases_ms"] = ctx.last_msm_phases()
    got = G.VariableBaseMSM.multi_scalar_mul(bases, sc)
    if want is not None:
        if affine_of(got, p) != affine_of(want, p):
            raise SystemExit("config 1: MSM differs from the CPU restatement")
        res["verified"] = "one-shot and resident results == the C++ restatement of variable_base.rs:10-83 (affine, every limb)"
        res["cpu_baseline"] = {"value": cpu_s * 1e3, "unit": "ms", "cores": threads, "kind": "port",
                               "sample": "the same 2^16-point MSM, whole"}
    d_sc.free()
    d_out.free()
    bases.free()
    return res


def run_config2(ctx, stream, G, args):
    """BASELINE config 2: the four radix-2 transforms at 2^20 (mnt4753::Fr; mnt6753::Fr tops out at 2^14 -
    SURVEY.md F2 - and is timed at that size).  Round trips checked here; the every-limb comparison of all
    four modes with the CPU restatement is tests/test_gpu_parity.py::test_ntt_config2_vs_cpp_restatement."""
    ffi = G.ffi
    out = {"workload": "radix-2 fft / ifft / coset_fft / coset_ifft, 2^20 on mnt4753::Fr and 2^14 (the maximum) on "
                       "mnt6753::Fr (BASELINE config 2)"}
    names = {ffi.FFT: "fft", ffi.IFFT: "ifft", ffi.COSET_FFT: "coset_fft", ffi.COSET_IFFT: "coset_ifft"}
    for tag, field, log_n in (("mnt4753_fr_2e20", ffi.FIELD_MNT4_FR, 20), ("mnt6753_fr_2e14", ffi.FIELD_MNT6_FR, 14)):
        n = 1 << log_n
        raw = random_field_elements(n, 0xC2 + log_n)
        vec = G.DeviceVector(ctx, field, n, raw)
        for fwd, inv in ((ffi.FFT, ffi.IFFT), (ffi.COSET_FFT, ffi.COSET_IFFT)):
            vec.ntt(fwd)
            vec.ntt(inv)
            if not np.array_equal(vec.download(), raw):
                raise SystemExit("config 2: round trip failed (%s)" % tag)
        ms = {names[m]: device_time(stream, lambda m=m: vec.ntt(m), 10, warm=2) for m in names}
        cpu = None
        if not args.no_cpu:
            cpu_s, want, threads = cpu_fft(field, raw, ffi.COSET_FFT)
            vec.upload(raw)
            vec.ntt(ffi.COSET_FFT)
            if not np.array_equal(vec.download(), want):
                raise SystemExit("config 2: coset_fft differs from the CPU restatement (%s)" % tag)
            cpu = {"value": cpu_s * 1e3, "unit": "ms", "cores": threads, "kind": "port", "sample": "one coset_fft_in_place, same input"}
        out[tag] = {"device_ms": ms, "cpu_baseline": cpu,
                    "verified": "ifft(fft(x)) == x, coset_ifft(coset_fft(x)) == x" + ("; coset_fft == the C++ restatement, every limb" if cpu else "")}
        vec.free()
    return out


def run_config4(ctx, stream, G, ffi, params, args, peak_mac_per_s):
    """BASELINE config 4: MNT6-753 G2 (over Fq3) MSM at 2^20 points on a resident synthetic key, verified
    by its discrete logs, and the mixed-radix transform at 2^15 * 25 = 819 200 points on mnt6753::Fr
    (round trip checked; parity unpinned - the reference has no mixed-radix domain)"""
    group, log_n = ffi.MNT6_G2, 20
    n = 1 << log_n
    t0 = time.perf_counter()
    bases = ctx.generate_bases(group, n, 0xC4)
    bases.precompute(args.copies)
    ctx.sync()
    key_s = time.perf_counter() - t0
    sc = random_scalars(n, 0xC5)
    r = params.GROUP_ORDER[group]
    out = None
    times = []
    for it in range(3):
        t1 = time.perf_counter()
        out = G.VariableBaseMSM.multi_scalar_mul(bases, sc)        # host scalars in, host point out
        times.append(time.perf_counter() - t1)
    phases = ctx.last_msm_phases()
    plan = ctx.last_msm_plan()
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
    _, macs, what = executed_accumulation(group, n, plan)
    acc_ms = phases.get("accumulate", 0.0)
    # mixed-radix transform
    N = (1 << 15) * 25
    field = ffi.FIELD_MNT6_FR
    raw = random_field_elements(N, 0xC6)
    vec = G.DeviceVector(ctx, field, N, raw)
    lib = ctx.lib
    for mode in (ffi.FFT, ffi.IFFT):
        lib.check(lib.ntt_mixed_dev(ctx.handle, field, vec.ptr, N, mode))
    if not (vec.download() == raw).all():
        raise SystemExit("config 4: mixed-radix ifft(fft(x)) != x")
    mixed_ms = device_time(stream, lambda: lib.check(lib.ntt_mixed_dev(ctx.handle, field, vec.ptr, N, ffi.FFT)), 5)
    vec.free()
    return {"msm": {"workload": "MNT6-753 G2 (Fq3) MSM, 2^20 points, resident key with precomputed copies (built in %.1f s)"
                                % key_s,
                    "ms": min(times) * 1e3, "mpts_per_s": n / min(times) / 1e6, "phases_ms": phases, "plan": plan,
                    "roofline": {"bound": "int32-mac", "kernel": "accumulation (Fq3 tower on 3 lanes, one lazily reduced coefficient per lane; %s)" % what, "kernel_ms": acc_ms,
                                 "executed_limb_macs": macs, "peak": peak_mac_per_s / 1e12, "unit": "Tlimb-MAC/s",
                                 "frac": macs / (acc_ms * 1e-3) / peak_mac_per_s if acc_ms else None},
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
    ap.add_argument("--cpu-groth16-log-n", type=int, default=16, help="domain of the CPU port's bounded create_proof sample")
    ap.add_argument("--copies", type=int, default=0,
                    help="precomputed shifted copies of the resident key (0 = auto by memory budget, 1 = plain key)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-fft", action="store_true")
    ap.add_argument("--no-groth16", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip configs 1, 2 and 4")
    ap.add_argument("--groth16-log-n", type=int, default=20)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
