#!/usr/bin/env python3
"""Scratch differential debugging on the GPU (not part of the product or the tests)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import g753 as O
from util753 import *

def run(group, n, c=None, seed=0):
    if c: os.environ["G753_MSM_C"] = str(c)
    else: os.environ.pop("G753_MSM_C", None)
    ctx = G.Context(0)
    C = GROUPS[group]
    pts = sample_points(C, n, 0xA0 + group + seed)
    sc = sample_scalars(C, n, 0xB0 + group + seed)
    coords, inf = points_to_arrays(C, pts)
    bases = ctx.upload_bases(group, coords, inf)
    got = projective_to_point(C, G.VariableBaseMSM.multi_scalar_mul(bases, ints_to_array(sc)))
    ok = got == O.msm_naive(C, pts, sc)
    print("group", group, "n", n, "c", c, "ok", ok, flush=True)
    bases.free(); ctx.close()
    return ok

if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "sweep"
    if mode == "sweep":
        for group in (1, 3):
            for n in (1, 2, 5, 40):
                for c in (3, 5, 8):
                    run(group, n, c)
    elif mode == "one":
        run(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))


def smul_sweep():
    import numpy as np
    ctx = G.Context(0)
    for group in (1, 3, 0):
        C = GROUPS[group]
        P = sample_points(C, 1, 0x33 + group)[0]
        ca, _ = points_to_arrays(C, [P])
        for s in (1, 2, 3, 4, 5, 6, 7, 8, 16, 17, 31, 32, 1 << 20, (1 << 40) + 12345, C.r - 1, C.r):
            out = np.zeros((3, C.F.k * 12), dtype=np.uint64)
            cb = ints_to_array([s])
            ctx.lib.check(ctx.lib.point_op(ctx.handle, group, 2, ffi.ptr(ca), ffi.ptr(cb), ffi.ptr(out)))
            try:
                got = projective_to_point(C, out)
                ok = got == C.mul(P, s)
            except AssertionError as e:
                ok = "noncanonical"
            print("smul group", group, "s", s if s < 1 << 41 else "big", "ok", ok, flush=True)
    ctx.close()


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "smul":
    smul_sweep()


def dump_scratch(libpath, group, n, c, out_path):
    import ctypes
    import numpy as np
    os.environ["G753_MSM_C"] = str(c)
    lib = ffi.Library(libpath) if libpath else None
    ctx = G.Context(0, library=lib)
    C = GROUPS[group]
    pts = sample_points(C, n, 0xA0 + group)
    sc = sample_scalars(C, n, 0xB0 + group)
    coords, inf = points_to_arrays(C, pts)
    bases = ctx.upload_bases(group, coords, None)
    out = G.VariableBaseMSM.multi_scalar_mul(bases, ints_to_array(sc))
    cap = ctypes.c_size_t(0)
    ctx.lib.check(ctx.lib.debug_scratch(ctx.handle, None, 0, ctypes.byref(cap)))
    buf = np.zeros(cap.value, dtype=np.uint8)
    ctx.lib.check(ctx.lib.debug_scratch(ctx.handle, ffi.ptr(buf), cap.value, ctypes.byref(cap)))
    np.savez_compressed(out_path, scratch=buf, out=out)
    print("dumped", cap.value, "bytes to", out_path, "ok",
          projective_to_point(C, out) == O.msm_naive(C, pts, sc), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "dump":
    # dump <libpath or ''> group n c out
    dump_scratch(sys.argv[2] or None, int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6])
