#!/usr/bin/env python3
"""the batched column / row passes of the SHARDED four-step transform on one GPU (world = 1, no collective):
what `ncu` can capture of the multi-GPU NTT's kernels (ncu is never run on a multi-rank command)"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
G = importlib.import_module("ginger-lib_b200")
D = importlib.import_module("ginger-lib_b200.distributed")
import bench
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
ctx = G.Context(0)
dom = D.ShardedEvaluationDomain(ctx, G.ffi.FIELD_MNT4_FR, log_n)
raw = bench.random_field_elements(1 << log_n, 5)
single = G.EvaluationDomain.new(G.ffi.FIELD_MNT4_FR, 1 << log_n, ctx=ctx).fft(raw)
for _ in range(3):
    t0 = time.perf_counter()
    out = dom.gather(dom.transform(dom.scatter(raw), G.ffi.FFT))
    dt = time.perf_counter() - t0
assert np.array_equal(out, single), "four-step result differs from the single transform"
print("four-step 2^%d on one GPU ok (host-staged, %.1f ms per call incl. copies)" % (log_n, dt * 1e3))
