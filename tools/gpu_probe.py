#!/usr/bin/env python3
"""First-contact GPU probe: integer-MAC roofline, MSM phase times, NTT times. Writes JSON lines
to gpurun_out/probe.jsonl.  Timing via CUDA events inside the library (mac probe, MSM phases)
or wall clock around synchronous C-ABI calls (reported as such)."""
import ctypes
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
G = importlib.import_module("ginger-lib_b200")
ffi = G.ffi
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
log = open(os.path.join(OUT, "probe.jsonl"), "a")


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    log.write(line + "\n")
    log.flush()


def main():
    what = sys.argv[1:] or ["mac", "msm", "ntt"]
    ctx = G.Context(0)
    lib = ctx.lib
    if "mac" in what:
        for variant, name, macs in ((0, "fq_mul", 1176), (1, "fq_sqr", 1176), (2, "imad_wide_stream", 576)):
            for blocks, threads in ((148, 128), (148, 256), (296, 256), (592, 256), (1184, 128)):
                iters = 2000
                ms = ctypes.c_float(0)
                lib.check(lib.mac_probe(ctx.handle, variant, blocks, threads, 50, ctypes.byref(ms)))
                lib.check(lib.mac_probe(ctx.handle, variant, blocks, threads, iters, ctypes.byref(ms)))
                total = blocks * threads * iters
                emit(probe="mac", variant=name, blocks=blocks, threads=threads, iters=iters, ms=ms.value,
                     ops_per_s=total / (ms.value * 1e-3), limb_macs_per_s=total * macs / (ms.value * 1e-3))
    if "msm" in what:
        rng = np.random.default_rng(1)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import util753 as U
        from oracle import g753 as O
        C = O.MNT4_G1
        base_pts = U.sample_points(C, 64, 0x51)
        coords64, _ = U.points_to_arrays(C, base_pts)
        for log_n in [int(x) for x in os.environ.get("PROBE_MSM_LOGS", "12,14,16,18,20").split(",")]:
            n = 1 << log_n
            coords = np.tile(coords64, (n // 64, 1))
            sc = rng.integers(0, 1 << 63, size=(n, 12), dtype=np.uint64)
            sc[:, 11] &= np.uint64(0xFF)
            bases = ctx.upload_bases(ffi.MNT4_G1, coords)
            for rep in range(2):
                t0 = time.time()
                out = G.VariableBaseMSM.multi_scalar_mul(bases, sc)
                dt = time.time() - t0
            emit(probe="msm", group="mnt4_g1", log_n=log_n, wall_s=dt, mpts_per_s=n / dt / 1e6,
                 phases_ms=ctx.last_msm_phases())
            bases.free()
    if "ntt" in what:
        rng = np.random.default_rng(2)
        for log_n in [int(x) for x in os.environ.get("PROBE_NTT_LOGS", "14,16,18,20,22").split(",")]:
            n = 1 << log_n
            raw = rng.integers(0, 1 << 63, size=(n, 12), dtype=np.uint64)
            raw[:, 11] &= np.uint64(0xFFFF)
            v = G.DeviceVector(ctx, ffi.FIELD_MNT4_FR, n, raw)
            v.ntt(ffi.FFT)
            ctx.sync()
            for mode, name in ((ffi.FFT, "fft"), (ffi.COSET_IFFT, "coset_ifft")):
                v.ntt(mode)
                ctx.sync()
                t0 = time.time()
                reps = 3
                for _ in range(reps):
                    v.ntt(mode)
                ctx.sync()
                dt = (time.time() - t0) / reps
                emit(probe="ntt", mode=name, log_n=log_n, s=dt, elems_per_s=n / dt,
                     hbm_frac_192n=192.0 * n / dt / 6543.4e9)
            v.free()
    ctx.close()


if __name__ == "__main__":
    main()
