#!/usr/bin/env python3
"""latency of small / plain-key MSMs (the literal multi_scalar_mul(&bases, &scalars) drop-in): wall time of
g753_msm on a resident plain key and of g753_msm_host (bases uploaded per call), with the device phases"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
G = importlib.import_module("ginger-lib_b200")
import bench
ffi = G.ffi
sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "6,10,16").split(",")]
groups = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,1,3").split(",")]
ctx = G.Context(0)
for g in groups:
    for log_n in sizes:
        n = 1 << log_n
        bases = ctx.generate_bases(g, n, 0x900 + g)
        coords = bases.download()
        sc = bench.random_scalars(n, 0x77 + g)
        res = {"group": g, "log_n": log_n}
        for name, fn in (("resident_plain", lambda: G.VariableBaseMSM.multi_scalar_mul(bases, sc)),
                         ("msm_host", lambda: G.VariableBaseMSM.multi_scalar_mul(coords, sc, group=g, ctx=ctx))):
            ts = []
            for rep in range(5):
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            res[name + "_ms"] = min(ts) * 1e3
            res[name + "_phases_ms"] = ctx.last_msm_phases()
            res[name + "_plan"] = ctx.last_msm_plan()
        print(json.dumps(res), flush=True)
        bases.free()
