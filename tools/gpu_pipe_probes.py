#!/usr/bin/env python3
"""multiplier-pipe and latency probes (g753_mac_probe): the IMAD Montgomery stream (the roofline's peak), the
FP64 instruction mix of the same product (bounded experiment), and the one-warp latencies of the cooperative
arithmetic (product, linear operation, G1 doubling / addition micro-programs)"""
import ctypes, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
G = importlib.import_module("ginger-lib_b200")
ctx = G.Context(0)
lib = ctx.lib
ms = ctypes.c_float(0)
out = {}
for name, variant, blocks, threads, iters in (("imad_fq_mul", 0, 592, 256, 2000), ("imad_fq_sqr", 1, 592, 256, 2000),
                                              ("dfma_mix_753", 4, 592, 256, 400), ("dfma_mix_753_128thr", 4, 1184, 128, 400)):
    lib.check(lib.mac_probe(ctx.handle, variant, blocks, threads, max(iters // 10, 1), ctypes.byref(ms)))
    lib.check(lib.mac_probe(ctx.handle, variant, blocks, threads, iters, ctypes.byref(ms)))
    out[name] = {"ms": ms.value, "products_per_s": blocks * threads * iters / (ms.value * 1e-3)}
for name, variant, iters in (("coop_mul_us", 5, 4000), ("coop_lin_us", 6, 4000), ("coop_g1_dbl_us", 7, 2000), ("coop_g1_add_us", 8, 1000)):
    lib.check(lib.mac_probe(ctx.handle, variant, 1, 32, 10, ctypes.byref(ms)))
    lib.check(lib.mac_probe(ctx.handle, variant, 1, 32, iters, ctypes.byref(ms)))
    out[name] = ms.value * 1e3 / iters
out["dfma_vs_imad"] = out["dfma_mix_753"]["products_per_s"] / out["imad_fq_mul"]["products_per_s"]
print(json.dumps(out))
