#!/usr/bin/env python3
"""per-group MSM phase times at 2^log_n on a generated key (plain and with precomputed copies)"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
G = importlib.import_module("ginger-lib_b200")
import bench
ffi = G.ffi
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
groups = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,1,3").split(",")]
copies_list = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,8").split(",")]
ctx = G.Context(0)
n = 1 << log_n
for g in groups:
    t0 = time.time()
    bases = ctx.generate_bases(g, n, 0x900 + g)
    gen_s = time.time() - t0
    sc = bench.random_scalars(n, 0x77 + g)
    for copies in copies_list:
        t0 = time.time()
        if copies != 1:
            bases.precompute(copies)
            ctx.sync()
        pre_s = time.time() - t0
        for rep in range(3):
            t0 = time.time()
            out = G.VariableBaseMSM.multi_scalar_mul(bases, sc)
            dt = time.time() - t0
        print(json.dumps({"group": g, "log_n": log_n, "copies": copies, "wall_ms": dt * 1e3, "gen_s": gen_s,
                          "precompute_s": pre_s, "phases_ms": ctx.last_msm_phases()}), flush=True)
    bases.free()
