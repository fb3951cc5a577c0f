"""CPU tier: the warp-cooperative group law (tools/gen_coop.py -> csrc/coop_programs.inc, interpreted by
csrc/coop.cuh) checked through its bit-exact Python model: the lane-split Montgomery product / addition /
subtraction against plain integers, and every generated micro-program (doubling, full and mixed addition,
conversions; all four groups) against the oracle's group law.  The same programs run on the GPU in
tests/test_gpu_parity.py (window fold, point folds) through the C ABI."""
import importlib.util
import os
import random

import pytest

from oracle import g753 as O

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("gen_coop", os.path.join(HERE, "..", "tools", "gen_coop.py"))
GC = importlib.util.module_from_spec(spec)
spec.loader.exec_module(GC)

CURVES = {0: O.MNT4_G1, 1: O.MNT4_G2, 2: O.MNT6_G1, 3: O.MNT6_G2}
R = 1 << 768


def test_generated_tables_are_current():
    """csrc/coop_programs.inc is what the generator emits from this tree"""
    path = os.path.join(HERE, "..", "ginger-lib_b200", "csrc", "coop_programs.inc")
    have = open(path).read()
    tmp = path + ".check"
    try:
        GC.emit(tmp)
        assert open(tmp).read() == have, "run python tools/gen_coop.py"
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)


@pytest.mark.parametrize("fid", [0, 1])
def test_lane_arithmetic(fid):
    p = GC.field_moduli()[fid]
    f = GC.Field(p)
    rng = random.Random(0xC00 + fid)
    rinv = pow(R, -1, p)
    edge = [0, 1, p - 1, p, 2 * p - 1, (1 << 96) - 1, ((1 << 96) - 1) << 96, (1 << 753) - 1, 100 * p,
            (1 << 192) - 1, sum(((1 << 96) - 1) << (192 * i) for i in range(4))]
    vals = edge + [rng.randrange(0, 181 * p) for _ in range(60)]
    for i, a in enumerate(vals):
        for b in vals[i % 7::7]:
            if (a // p + 1) * (b // p + 1) > (1 << 15):
                continue
            got = GC.join(f.mul(GC.split(a), GC.split(b)))
            assert got % p == a * b * rinv % p
            assert got < p + (a * b >> 768) + 1       # (a b + M p) / R with M < R
    for a in vals:
        for b in vals[::5]:
            if a + b < R:
                assert GC.join(f.add(GC.split(a), GC.split(b))) == a + b
            for k in range(GC.MAX_K_LOG + 1):
                if b <= (p << k) and a + (p << k) - b < R:
                    assert GC.join(f.sub(GC.split(a), GC.split(b), k)) == a + (p << k) - b
    # the fused linear operation ca a +- cb b +- cc c (+ 2^k p)
    for _ in range(300):
        terms, want, neg = [], 0, 0
        for _ in range(rng.randrange(1, 4)):
            x, c = rng.choice(vals), rng.randrange(1, GC.CMAX + 1) * rng.choice((1, -1))
            terms.append((x, c))
            want += c * x
            neg += -c * x if c < 0 else 0
        k = 0
        while (p << k) < neg:
            k += 1
        if k > GC.MAX_K_LOG or not any(c < 0 for _, c in terms):
            k = 0
        want += (p << k) if any(c < 0 for _, c in terms) else 0
        if 0 <= want < R:
            assert GC.join(f.lin([(GC.split(x), c) for x, c in terms], k)) == want
    for a in [0, 1, p - 1, p, p + 1, 2 * p - 1] + [rng.randrange(0, 2 * p) for _ in range(40)]:
        assert GC.join(f.cond_sub_p(GC.split(a))) == a % p
        assert f.is_zero_mod_p(GC.split(a)) == (a % p == 0)


def _load_point(sim, F, base, k, P, rng, p, bounds=None):
    """affine oracle point -> XYZZ slots with a random ZZ (any representative), Montgomery form, each
    coordinate lifted by a random multiple of p inside its bound (the programs accept lazily reduced input)"""
    if P is None:
        zz = zzz = tuple([0] * k)
        x = tuple(rng.randrange(p) for _ in range(k))
        y = tuple(rng.randrange(p) for _ in range(k))
    else:
        z = tuple(rng.randrange(1, p) for _ in range(k))
        zz, zzz = F.sqr(z), F.mul(F.sqr(z), z)
        x, y = F.mul(P[0], zz), F.mul(P[1], zzz)
    for i, (coord, nm) in enumerate(((x, "X"), (y, "Y"), (zz, "ZZ"), (zzz, "ZZZ"))):
        for j, c in enumerate(coord):
            v = c * R % p
            if bounds is not None:
                v += p * rng.randrange(int(bounds[nm]))
            sim.put(base + i * k + j, v)


def _read_point(sim, F, base, k, p):
    rinv = pow(R, -1, p)
    c = [tuple(sim.get(base + i * k + j) * rinv % p for j in range(k)) for i in range(4)]
    x, y, zz, zzz = c
    if F.is_zero(zz):
        return None
    return (F.mul(x, F.inv(zz)), F.mul(y, F.inv(zzz)))


@pytest.mark.parametrize("gid", sorted(CURVES))
def test_programs_vs_oracle(gid):
    from util753 import sample_points
    C = CURVES[gid]
    F, k = C.F, C.F.k
    p = C.F.base.p
    f = GC.Field(p)
    progs = GC.build_group(gid)
    lay = GC.Layout(k)
    n_slots = max(pr.n_slots for pr in progs.values())
    rng = random.Random(0xC0D + gid)
    pts = sample_points(C, 4, 0xC0 + gid)
    one = R % p

    def fresh():
        s = GC.Sim(f, n_slots)
        s.put(lay.ONE, one)
        return s

    # doubling, also of infinity and after a previous program (lazily reduced accumulator)
    for P in pts[:2] + [None]:
        s = fresh()
        _load_point(s, F, lay.P, k, P, rng, p, GC.ACC_BOUND)
        s.run(progs["dbl"])
        assert _read_point(s, F, lay.P, k, p) == C.double(P)
        s.run(progs["dbl"])
        assert _read_point(s, F, lay.P, k, p) == C.double(C.double(P))
        for i in range(4 * k):     # the persistent bounds hold
            nm = ("X", "Y", "ZZ", "ZZZ")[i // k]
            assert s.get(lay.P + i) < GC.ACC_BOUND[nm] * p

    def is_zero_slots(s, slots):
        return all(f.is_zero_mod_p(s.slots[x]) for x in slots)

    # additions: generic, P + P (detected by the head), P + (-P)
    for tag in ("add", "madd"):
        head, tail = progs[tag + "_head"], progs[tag + "_tail"]
        for P, Q in ((pts[0], pts[1]), (pts[2], pts[3]), (pts[0], pts[0]), (pts[1], C.neg(pts[1]))):
            s = fresh()
            _load_point(s, F, lay.P, k, P, rng, p, GC.ACC_BOUND)
            if tag == "madd":
                for i, coord in enumerate((Q[0], Q[1], F.one(), F.one())):
                    for j, c in enumerate(coord):
                        s.put(lay.Q + i * k + j, c * R % p)
            else:
                _load_point(s, F, lay.Q, k, Q, rng, p)
            s.run(head)
            pz = is_zero_slots(s, head.test_slots["tP"])
            rz = is_zero_slots(s, head.test_slots["tR"])
            assert pz == (P[0] == Q[0])
            if pz:
                assert rz == (P == Q)
                continue
            s.run(tail)
            assert _read_point(s, F, lay.P, k, p) == C.add(P, Q)
            s.run(progs["dbl"])
            assert _read_point(s, F, lay.P, k, p) == C.double(C.add(P, Q))
    # conversions (outputs below 2 p: one conditional subtraction from canonical)
    s = fresh()
    _load_point(s, F, lay.P, k, pts[0], rng, p, GC.ACC_BOUND)
    s.run(progs["to_proj"])
    rinv = pow(R, -1, p)
    assert all(s.get(lay.P + i) < 2 * p for i in range(3 * k))
    X, Y, Z = (tuple(s.get(lay.P + i * k + j) * rinv % p for j in range(k)) for i in range(3))
    assert C.from_projective(X, Y, Z) == pts[0]
    s = fresh()
    _load_point(s, F, lay.P, k, pts[2], rng, p, GC.ACC_BOUND)
    s.run(progs["reduce"])
    assert all(s.get(lay.P + i) < 2 * p for i in range(4 * k))
    assert _read_point(s, F, lay.P, k, p) == pts[2]
    s = fresh()
    z = tuple(rng.randrange(1, p) for _ in range(k))
    for i, coord in enumerate((F.mul(pts[1][0], z), F.mul(pts[1][1], z), z)):
        for j, c in enumerate(coord):
            s.put(lay.Q + i * k + j, c * R % p)
    s.run(progs["from_proj"])
    assert _read_point(s, F, lay.Q, k, p) == pts[1]
