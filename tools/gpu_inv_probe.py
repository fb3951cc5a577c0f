import ctypes, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
G = importlib.import_module("ginger-lib_b200")
ctx = G.Context(0)
lib = ctx.lib
ms = ctypes.c_float(0)
for variant, iters in ((0, 2000), (3, 50)):
    for blocks, threads in ((592, 256), (296, 128)):
        lib.check(lib.mac_probe(ctx.handle, variant, blocks, threads, max(iters // 10, 2), ctypes.byref(ms)))
        lib.check(lib.mac_probe(ctx.handle, variant, blocks, threads, iters, ctypes.byref(ms)))
        print(json.dumps({"variant": variant, "blocks": blocks, "threads": threads, "iters": iters, "ms": ms.value,
                          "ops_per_s": blocks * threads * iters / (ms.value * 1e-3)}), flush=True)
