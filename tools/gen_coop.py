#!/usr/bin/env python3
"""Generator (and bit-exact Python model) of the WARP-COOPERATIVE group law used by the latency-bound
kernels of the MSM (window fold, reduction tail, partial-point fold; csrc/coop.cuh).

Why: one thread needs ~2.4 us per 753-bit Montgomery product (2352 IMAD issue slots on one scheduler),
so the reference's serial Horner fold (variable_base.rs:72-82: 753 doublings) costs ~31 ms on one thread
however many SMs idle beside it.  The cooperative form spreads ONE field product over the 8 lanes of an
"octet" (3 limbs of 32 bits per lane, digit-serial Montgomery in base 2^96, carries resolved with a
ballot) and runs up to four independent products of a curve formula on the four octets of a warp.  The
formulas are compiled HERE into straight-line micro-programs (levels of <= 4 field operations on
shared-memory slots) which csrc/coop.cuh interprets; this file is also their exact model:

  * Builder / Fk      formulas as a DAG over base-field operations (towers expanded: Karatsuba Fq2 /
                      Fq3 products, complex / Chung-Hasan squarings - fp2.rs:128-144, 387-401,
                      fp3.rs:165-185, 451-478; XYZZ dbl-2008-s-1 / add-2008-s / madd-2008-s)
  * linear algebra    sums, differences and small multiples are kept as LINEAR EXPRESSIONS over the products and
                      inputs and only materialised - as ONE fused operation of up to three terms,
                      d = ca A +- cb B +- cc C (+ 2^k p) - where a product or an output needs them: chains like
                      M = 3 X^2 + a ZZ^2, X3 = M^2 - 2 S, S - X3 = 3 S - M^2 collapse into single rows
  * schedule()        list scheduling into levels (linear operations first, then up to 4 products)
  * allocate()        slot allocation with in-place outputs
  * bounds            values are kept LAZILY reduced: a product of inputs < x p, y p is
                      < (1 + x y / 2^15) p because R / p >= 2^15 (R = 2^768, p < 2^753); sums add bounds;
                      a - b is computed as a + K p - b with K = 2^k >= bound(b).  The generator proves
                      every value < 2^768 and every product input bound x y <= 2^15.
  * emit()            csrc/coop_programs.inc (C tables)
  * Sim               executes a program limb by limb, lane by lane, exactly as coop.cuh does

Run: python tools/gen_coop.py            (rewrites ginger-lib_b200/csrc/coop_programs.inc)
Tested by tests/test_coop_model.py (model vs oracle) and, on the GPU, through the C ABI.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

NOP, MUL, LIN = 0, 1, 2
OPNAME = {NOP: "nop", MUL: "mul", LIN: "lin"}
CMAX = 15                              # largest |coefficient| of a fused linear operation
LANES, LPL = 8, 3                      # lanes per field element, limbs per lane
MASK32 = 0xFFFFFFFF
MASK96 = (1 << 96) - 1
R_BITS = 768
MAX_K_LOG = 11                         # K p tables for K = 2^0 .. 2^11
MUL_HEADROOM = 1 << 15                 # R / p >= 2^15 for both fields


# ---------------------------------------------------------------------------------------------
# formula DAG
# ---------------------------------------------------------------------------------------------
class Node:
    """op "in" (input slot), MUL (a, b) or LIN (terms = [(node, coef)], adds 2^k p when a coef is negative)"""
    __slots__ = ("op", "a", "b", "terms", "bound", "k", "idx", "slot", "level", "name")

    def __init__(self, op, a=None, b=None, terms=None, bound=1.0, k=0, name=""):
        self.op, self.a, self.b, self.terms, self.bound, self.k, self.name = op, a, b, terms, bound, k, name
        self.idx = -1
        self.slot = None
        self.level = -1

    def operands(self):
        if self.op == MUL:
            return [self.a, self.b]
        if self.op == LIN:
            return [n for n, _ in self.terms]
        return []


REDUCE_ABOVE = 181.0      # sqrt(2^15): any two values at or below it may be multiplied


def _klog(x):
    k = 0
    while (1 << k) < x:
        k += 1
    return k


class Val:
    """a linear expression sum_i coef_i * node_i over inputs and products (never over other sums)"""
    __slots__ = ("b", "e")

    def __init__(self, bld, expr):
        self.b = bld
        self.e = {i: c for i, c in expr.items() if c}

    @property
    def bound(self):
        nodes = self.b.nodes
        pos = sum(c * nodes[i].bound for i, c in self.e.items() if c > 0)
        neg = sum(-c * nodes[i].bound for i, c in self.e.items() if c < 0)
        return pos + ((1 << _klog(neg)) if neg > 0 else 0.0)


class Builder:
    def __init__(self, one_slot=None):
        self.nodes = []
        self.one_slot = one_slot
        self._one = None
        self._reduced = {}
        self._lin = {}

    def _add(self, n):
        n.idx = len(self.nodes)
        self.nodes.append(n)
        return n

    def _val(self, node):
        return Val(self, {node.idx: 1})

    def inp(self, slot, bound, name=""):
        n = self._add(Node("in", bound=float(bound), name=name))
        n.slot = slot
        return self._val(n)

    def one(self):
        if self._one is None:
            self._one = self.inp(self.one_slot, 1.0, "one")
        return self._one

    # ---- linear algebra: no operations are emitted here -------------------------------------------------
    def add(self, a, b):
        e = dict(a.e)
        for i, c in b.e.items():
            e[i] = e.get(i, 0) + c
        return Val(self, e)

    def sub(self, a, b):
        e = dict(a.e)
        for i, c in b.e.items():
            e[i] = e.get(i, 0) - c
        return Val(self, e)

    def small(self, a, c):
        """c * a for a small positive integer c (the curve / non-residue constants 2, 11, 13, 26, 121 of
        SURVEY.md appendix A)"""
        return Val(self, {i: c * x for i, x in a.e.items()})

    # ---- materialisation -----------------------------------------------------------------------------------
    def _reduce_node(self, n):
        """n -> n R R^-1 = n (mod p), below (1 + bound / 2^15) p: one Montgomery product by ONE"""
        if n.idx not in self._reduced:
            one = self.node_of(self.one())
            self._reduced[n.idx] = self._add(Node(MUL, n, one, bound=1.0 + n.bound / MUL_HEADROOM))
        return self._reduced[n.idx]

    def reduce(self, a):
        return self._val(self._reduce_node(self.node_of(a)))

    def _lin_node(self, terms):
        """one fused operation: terms = [(node, coef)], at most three, |coef| <= CMAX"""
        assert 1 <= len(terms) <= 3 and all(0 < abs(c) <= CMAX for _, c in terms)
        key = tuple(sorted((n.idx, c) for n, c in terms))
        if key in self._lin:
            return self._lin[key]
        while True:
            pos = sum(c * n.bound for n, c in terms if c > 0)
            neg = sum(-c * n.bound for n, c in terms if c < 0)
            k = _klog(neg) if neg > 0 else 0
            if k <= MAX_K_LOG and pos + (1 << k) < (1 << 14):
                break
            # bring the largest operand down first
            j = max(range(len(terms)), key=lambda t: abs(terms[t][1]) * terms[t][0].bound)
            assert terms[j][0].bound > 2.0, "cannot bound a linear operation"
            terms = [(self._reduce_node(n) if t == j else n, c) for t, (n, c) in enumerate(terms)]
        node = self._add(Node(LIN, terms=list(terms), bound=pos + ((1 << k) if neg > 0 else 0.0), k=k))
        self._lin[key] = node
        return node

    def node_of(self, v, force=False):
        """the node holding v: an input / product itself, or fused linear operations computing it"""
        nodes = self.nodes
        terms = sorted(((nodes[i], c) for i, c in v.e.items()), key=lambda t: (-(abs(t[1]) * t[0].bound), t[0].idx))
        assert terms, "the zero expression has no node"
        if len(terms) == 1 and terms[0][1] == 1 and not force:
            return terms[0][0]
        # coefficients beyond CMAX: scale the operand in steps (121 = 11 * 11, 26 * 5 = 13 * 10 ...)
        fixed = []
        for n, c in terms:
            while abs(c) > CMAX:
                f = next((d for d in range(CMAX, 1, -1) if abs(c) % d == 0), None)
                if f is None:                         # no small factor: c x = CMAX (q x) ... + r x handled as two terms
                    q, r = divmod(abs(c), CMAX)
                    t = self._lin_node([(n, CMAX)])
                    sgn = 1 if c > 0 else -1
                    if r:
                        fixed.append((n, sgn * r))
                    n, c = t, sgn * q
                else:
                    n = self._lin_node([(n, f)])
                    c = c // f
            fixed.append((n, c))
        terms = fixed
        # more than three terms: fold the three heaviest into a temporary first
        while len(terms) > 3:
            t = self._lin_node(terms[:3])
            terms = [(t, 1)] + terms[3:]
        return self._lin_node(terms)

    def mul(self, a, b):
        na, nb = self.node_of(a), self.node_of(b)
        same = na is nb
        # lazily reduced operands: bring the larger one down when the product would leave the headroom
        while na.bound * nb.bound > MUL_HEADROOM:
            if na.bound >= nb.bound:
                assert na.bound > 2.0
                na = self._reduce_node(na)
                if same:
                    nb = na
            else:
                nb = self._reduce_node(nb)
        return self._val(self._add(Node(MUL, na, nb, bound=1.0 + na.bound * nb.bound / MUL_HEADROOM)))


class Fk:
    """element of Fq, Fq2 = Fq[u]/(u^2 - nr) or Fq3 = Fq[u]/(u^3 - nr) as a tuple of linear expressions"""

    def __init__(self, bld, c, nr):
        self.b, self.c, self.nr = bld, tuple(c), nr

    @property
    def k(self):
        return len(self.c)

    def _w(self, c):
        return Fk(self.b, c, self.nr)

    def __add__(self, o):
        return self._w([self.b.add(x, y) for x, y in zip(self.c, o.c)])

    def __sub__(self, o):
        return self._w([self.b.sub(x, y) for x, y in zip(self.c, o.c)])

    def dbl(self):
        return self + self

    def __mul__(self, o):
        B, nr = self.b, self.nr
        a, b = self.c, o.c
        if self.k == 1:
            return self._w([B.mul(a[0], b[0])])
        if self.k == 2:     # Karatsuba, fp2.rs:387-401
            v0, v1 = B.mul(a[0], b[0]), B.mul(a[1], b[1])
            m = B.mul(B.add(a[0], a[1]), B.add(b[0], b[1]))
            c1 = B.sub(B.sub(m, v0), v1)
            c0 = B.add(v0, B.small(v1, nr))
            return self._w([c0, c1])
        # Karatsuba, fp3.rs:451-478
        v0, v1, v2 = B.mul(a[0], b[0]), B.mul(a[1], b[1]), B.mul(a[2], b[2])
        m12 = B.mul(B.add(a[1], a[2]), B.add(b[1], b[2]))
        m01 = B.mul(B.add(a[0], a[1]), B.add(b[0], b[1]))
        m02 = B.mul(B.add(a[0], a[2]), B.add(b[0], b[2]))
        c0 = B.add(v0, B.small(B.sub(B.sub(m12, v1), v2), nr))
        c1 = B.add(B.sub(B.sub(m01, v0), v1), B.small(v2, nr))
        c2 = B.add(B.sub(B.sub(m02, v0), v2), v1)
        return self._w([c0, c1, c2])

    def sqr(self):
        B, nr = self.b, self.nr
        a = self.c
        if self.k == 1:
            return self._w([B.mul(a[0], a[0])])
        if self.k == 2:     # complex squaring, fp2.rs:128-144
            ab = B.mul(a[0], a[1])
            t = B.mul(B.add(a[0], a[1]), B.add(a[0], B.small(a[1], nr)))
            c0 = B.sub(B.sub(t, ab), B.small(ab, nr))
            return self._w([c0, B.add(ab, ab)])
        # Chung-Hasan SQR2, fp3.rs:165-185
        s0 = B.mul(a[0], a[0])
        ab = B.mul(a[0], a[1])
        s1 = B.add(ab, ab)
        t = B.add(B.sub(a[0], a[1]), a[2])
        s2 = B.mul(t, t)
        bc = B.mul(a[1], a[2])
        s3 = B.add(bc, bc)
        s4 = B.mul(a[2], a[2])
        c0 = B.add(s0, B.small(s3, nr))
        c1 = B.add(s1, B.small(s4, nr))
        c2 = B.sub(B.sub(B.add(B.add(s1, s2), s3), s0), s4)
        return self._w([c0, c1, c2])


# curve descriptors: tower degree, non-residue, multiplication by the curve coefficient a
# (curves/mnt4753/g1.rs:18-33, g2.rs:112-118, curves/mnt6753/g1.rs:18-33, g2.rs:148-155)
def _mul_a_m4g1(x):
    return x.dbl()


def _mul_a_m6g1(x):
    return x._w([x.b.small(x.c[0], 11)])


def _mul_a_m4g2(x):
    return x._w([x.b.small(x.c[0], 26), x.b.small(x.c[1], 26)])


def _mul_a_m6g2(x):      # a' = 11 u^2: (c0, c1, c2) -> (121 c1, 121 c2, 11 c0)
    return x._w([x.b.small(x.c[1], 121), x.b.small(x.c[2], 121), x.b.small(x.c[0], 11)])


GROUPS = {
    0: dict(name="m4g1", k=1, nr=0, field=0, mul_a=_mul_a_m4g1),
    1: dict(name="m4g2", k=2, nr=13, field=0, mul_a=_mul_a_m4g2),
    2: dict(name="m6g1", k=1, nr=0, field=1, mul_a=_mul_a_m6g1),
    3: dict(name="m6g2", k=3, nr=11, field=1, mul_a=_mul_a_m6g2),
}

# persistent bounds (in units of p) of the coordinates of an accumulator between programs, and of a
# point freshly loaded from global memory (canonical)
ACC_BOUND = dict(X=24.0, Y=8.0, ZZ=2.0, ZZZ=2.0)
CANON = 1.0


class Layout:
    """slot map of one warp: accumulator P (4K slots: X, Y, ZZ, ZZZ), operand Q (4K), ONE (Montgomery
    one), then scratch"""

    def __init__(self, k):
        self.k = k
        self.P = 0
        self.Q = 4 * k
        self.ONE = 8 * k
        self.SCRATCH = 8 * k + 1


def _point(bld, g, base, bounds, tag):
    k, nr = g["k"], g["nr"]
    out = []
    for i, nm in enumerate(("X", "Y", "ZZ", "ZZZ")):
        out.append(Fk(bld, [bld.inp(base + i * k + j, bounds[nm], "%s.%s%d" % (tag, nm, j)) for j in range(k)], nr))
    return out


def formula_dbl(g):
    """P = 2 P (dbl-2008-s-1, general a): ZZ = 0 stays ZZ = 0, Y = 0 gives ZZ3 = 0"""
    lay = Layout(g["k"])
    bld = Builder(lay.ONE)
    X, Y, ZZ, ZZZ = _point(bld, g, lay.P, ACC_BOUND, "P")
    U = Y.dbl()
    V = U.sqr()
    W = U * V
    S = X * V
    XX = X.sqr()
    M = XX.dbl() + XX + g["mul_a"](ZZ.sqr())
    X3 = M.sqr() - S.dbl()
    Y3 = M * (S - X3) - W * Y
    ZZ3 = V * ZZ
    ZZZ3 = W * ZZZ
    return bld, lay, dict(X=X3, Y=Y3, ZZ=ZZ3, ZZZ=ZZZ3)


def formula_add_head(g, mixed):
    """first half of P += Q (add-2008-s; mixed: Q affine, ZZ2 = ZZZ2 = 1): leaves U1, S1, PP = P^2?, no -
    leaves Pd = U2 - U1 and Rd = S2 - S1 in scratch together with their images under a Montgomery
    product by ONE (< (1 + eps) p, so zero <=> 0 or p) for the exceptional-case tests"""
    lay = Layout(g["k"])
    bld = Builder(lay.ONE)
    X1, Y1, ZZ1, ZZZ1 = _point(bld, g, lay.P, ACC_BOUND, "P")
    qb = dict(X=CANON, Y=CANON, ZZ=CANON, ZZZ=CANON)
    X2, Y2, ZZ2, ZZZ2 = _point(bld, g, lay.Q, qb, "Q")
    one = bld.one()
    if mixed:
        U1, S1 = X1, Y1
    else:
        U1, S1 = X1 * ZZ2, Y1 * ZZZ2
    Pd = X2 * ZZ1 - U1
    Rd = Y2 * ZZZ1 - S1
    k, nr = g["k"], g["nr"]
    tP = Fk(bld, [bld.reduce(c) for c in Pd.c], nr)
    tR = Fk(bld, [bld.reduce(c) for c in Rd.c], nr)
    return bld, lay, dict(U1=U1, S1=S1, Pd=Pd, Rd=Rd, tP=tP, tR=tR)


def formula_add_tail(g, mixed, head_bounds):
    """second half: from U1, S1, Pd, Rd (scratch slots fixed by the head's allocation) to P"""
    lay = Layout(g["k"])
    bld = Builder(lay.ONE)
    k, nr = g["k"], g["nr"]
    X1, Y1, ZZ1, ZZZ1 = _point(bld, g, lay.P, ACC_BOUND, "P")
    qb = dict(X=CANON, Y=CANON, ZZ=CANON, ZZZ=CANON)
    X2, Y2, ZZ2, ZZZ2 = _point(bld, g, lay.Q, qb, "Q")

    def ext(name):
        slots, bounds = head_bounds[name]
        return Fk(bld, [bld.inp(s, b, name + str(j)) for j, (s, b) in enumerate(zip(slots, bounds))], nr)

    U1 = X1 if mixed else ext("U1")
    S1 = Y1 if mixed else ext("S1")
    Pd, Rd = ext("Pd"), ext("Rd")
    PP = Pd.sqr()
    PPP = Pd * PP
    Q = U1 * PP
    X3 = Rd.sqr() - PPP - Q.dbl()
    Y3 = Rd * (Q - X3) - S1 * PPP
    if mixed:
        ZZ3 = ZZ1 * PP
        ZZZ3 = ZZZ1 * PPP
    else:
        ZZ3 = ZZ1 * ZZ2 * PP
        ZZZ3 = ZZZ1 * ZZZ2 * PPP
    return bld, lay, dict(X=X3, Y=Y3, ZZ=ZZ3, ZZZ=ZZZ3)


def formula_to_projective(g):
    """XYZZ -> the reference's homogeneous (X ZZZ : Y ZZ : ZZ ZZZ), written over X, Y, ZZ of P"""
    lay = Layout(g["k"])
    bld = Builder(lay.ONE)
    X, Y, ZZ, ZZZ = _point(bld, g, lay.P, ACC_BOUND, "P")
    # (compiled with the limit 2 p, so that one conditional subtraction makes the coordinates canonical)
    return bld, lay, dict(X=X * ZZZ, Y=Y * ZZ, ZZ=ZZ * ZZZ)


def formula_reduce(g):
    """every coordinate of P below 2 p (one conditional subtraction away from canonical): X and Y through a
    product by ONE, ZZ and ZZZ already are"""
    lay = Layout(g["k"])
    bld = Builder(lay.ONE)
    X, Y, ZZ, ZZZ = _point(bld, g, lay.P, ACC_BOUND, "P")
    assert ACC_BOUND["ZZ"] <= 2.0 and ACC_BOUND["ZZZ"] <= 2.0
    outs = dict(X=X._w([bld.reduce(n) for n in X.c]), Y=Y._w([bld.reduce(n) for n in Y.c]))
    return bld, lay, outs


def formula_from_projective(g):
    """homogeneous (X : Y : Z) in Q (slots X, Y, ZZ of Q) -> XYZZ (X Z, Y Z^2, Z^2, Z^3) in Q"""
    lay = Layout(g["k"])
    bld = Builder(lay.ONE)
    k, nr = g["k"], g["nr"]
    X = Fk(bld, [bld.inp(lay.Q + j, CANON, "X%d" % j) for j in range(k)], nr)
    Y = Fk(bld, [bld.inp(lay.Q + k + j, CANON, "Y%d" % j) for j in range(k)], nr)
    Z = Fk(bld, [bld.inp(lay.Q + 2 * k + j, CANON, "Z%d" % j) for j in range(k)], nr)
    Z2 = Z.sqr()
    return bld, lay, dict(X=X * Z, Y=Y * Z2, ZZ=Z2, ZZZ=Z * Z2), lay.Q


# ---------------------------------------------------------------------------------------------
# scheduling and slot allocation
# ---------------------------------------------------------------------------------------------
def output_nodes(bld, outputs, limit=None):
    """materialise the output expressions; where a lazily reduced value would exceed the accumulator's
    persistent bound it goes through a product by ONE.  name -> list of nodes (one per tower coefficient)"""
    out = {}
    for nm, v in outputs.items():
        nodes = []
        for x in v.c:
            n = bld.node_of(x)
            if limit is not None and nm in limit and n.bound > limit[nm]:
                n = bld._reduce_node(n)
            nodes.append(n)
        out[nm] = nodes
    return out


def live_nodes(bld, out_nodes):
    keep = set()
    stack = [n for ns in out_nodes.values() for n in ns]
    while stack:
        n = stack.pop()
        if n.idx in keep:
            continue
        keep.add(n.idx)
        stack.extend(n.operands())
    return [n for n in bld.nodes if n.idx in keep]


def schedule(nodes):
    """levels of <= 4 operations: all ready linear operations first (cheap levels), then up to four
    products chosen by the longest path to an output"""
    height = {}
    users = {n.idx: [] for n in nodes}
    for n in nodes:
        for x in n.operands():
            users[x.idx].append(n)
    for n in reversed(nodes):
        h = 0
        for u in users[n.idx]:
            h = max(h, height[u.idx])
        height[n.idx] = h + (10 if n.op == MUL else 2 if n.op == LIN else 0)
    done = {n.idx for n in nodes if n.op == "in"}
    pending = [n for n in nodes if n.op != "in"]
    levels = []
    while pending:
        ready = [n for n in pending if all(x.idx in done for x in n.operands())]
        assert ready
        lin = [n for n in ready if n.op == LIN]
        take = sorted(lin if lin else ready, key=lambda n: -height[n.idx])[:4]
        for n in take:
            n.level = len(levels)
            done.add(n.idx)
            pending.remove(n)
        levels.append(take)
    return levels


def allocate(nodes, levels, out_nodes, out_base, lay):
    """slots for every node.  Inputs sit in their fixed slots; output coordinate `name` is pinned to the
    accumulator slot out_base + index; everything else takes scratch slots.  Within a level all operands
    are read before any result is written, so a slot whose last reader is in level t may be written in t.
    Returns the (slot, node) copies still to be made (pinned slot busy, or the output is an input) and the
    number of slots used."""
    k = lay.k
    order = {"X": 0, "Y": 1, "ZZ": 2, "ZZZ": 3}
    pinned = {}
    if out_base is not None:
        for name, ns in out_nodes.items():
            if name in order:
                for j, n in enumerate(ns):
                    pinned[n.idx] = out_base + order[name] * k + j
    last_use = {n.idx: -1 for n in nodes}
    for lv, ops in enumerate(levels):
        for n in ops:
            for x in n.operands():
                last_use[x.idx] = max(last_use[x.idx], lv)
    INF = 1 << 30
    for ns in out_nodes.values():
        for n in ns:
            last_use[n.idx] = INF
    occupied = {}     # slot -> node idx
    for n in nodes:
        if n.op == "in":
            occupied[n.slot] = n.idx
    scratch_next = [max([lay.SCRATCH] + [n.slot + 1 for n in nodes if n.op == "in"])]
    free_scratch = []
    copies = []

    def release(level):
        for s, i in list(occupied.items()):
            if last_use[i] <= level:
                del occupied[s]
                if s >= lay.SCRATCH:
                    free_scratch.append(s)

    def take_scratch():
        if free_scratch:
            free_scratch.sort()
            return free_scratch.pop(0)
        s = scratch_next[0]
        scratch_next[0] += 1
        return s

    for lv, ops in enumerate(levels):
        release(lv)
        for n in ops:
            want = pinned.get(n.idx)
            if want is not None and want not in occupied:
                n.slot = want
            else:
                n.slot = take_scratch()
                if want is not None:
                    copies.append((want, n))
            occupied[n.slot] = n.idx
    for n in nodes:                      # an output that IS an input but belongs elsewhere
        if n.op == "in" and n.idx in pinned and pinned[n.idx] != n.slot:
            copies.append((pinned[n.idx], n))
    n_slots = max([scratch_next[0]] + [n.slot + 1 for n in nodes if n.slot is not None])
    return copies, n_slots


HAS_MUL, HAS_LIN = 1 << 28, 1 << 29


def encode(op, d=0, terms=(), k=0):
    """two 32-bit words: op (3 bits) | dst << 3 | a << 10 | b << 17 | k << 24   and
    c | (|ca| << 7) | (sign a << 11) | (|cb| << 12) | (sign b << 16) | (|cc| << 17) | (sign c << 21);
    a product uses a and b only"""
    slots = [t[0] for t in terms] + [0] * (3 - len(terms))
    coefs = [t[1] for t in terms] + [0] * (3 - len(terms))
    assert all(0 <= x < 128 for x in slots + [d]) and 0 <= k < 16 and all(abs(c) <= CMAX for c in coefs)
    w0 = op | (d << 3) | (slots[0] << 10) | (slots[1] << 17) | (k << 24)
    w1 = slots[2]
    for j, c in enumerate(coefs):
        w1 |= (abs(c) << (7 + 5 * j)) | ((1 if c < 0 else 0) << (11 + 5 * j))
    return w0, w1


def assemble(levels, copies):
    """rows of 4 instructions (one per octet) of two words each; word 0 of a row carries its flags"""
    rows = []

    def row_of(instrs, flags):
        instrs = instrs + [encode(NOP)] * (4 - len(instrs))
        words = [w for ins in instrs for w in ins]
        words[0] |= flags
        return words
    for ops in levels:
        instrs = []
        for n in ops:
            if n.op == MUL:
                instrs.append(encode(MUL, n.slot, [(n.a.slot, 1), (n.b.slot, 1)]))
            else:
                instrs.append(encode(LIN, n.slot, [(x.slot, c) for x, c in n.terms], n.k))
        rows.append(row_of(instrs, HAS_MUL if ops[0].op == MUL else HAS_LIN))
    for i in range(0, len(copies), 4):
        rows.append(row_of([encode(LIN, d, [(n.slot, 1)]) for d, n in copies[i:i + 4]], HAS_LIN))
    return rows


class Program:
    def __init__(self, name, words, n_slots, info):
        self.name, self.words, self.n_slots, self.info = name, words, n_slots, info


def compile_formula(name, bld, lay, outputs, out_base, limit=None):
    out_nodes = output_nodes(bld, outputs, limit)
    nodes = live_nodes(bld, out_nodes)
    levels = schedule(nodes)
    for ops in levels:
        assert len({n.op for n in ops}) == 1
    copies, n_slots = allocate(nodes, levels, out_nodes, out_base, lay)
    if limit is not None:
        for nm, ns in out_nodes.items():
            if nm in limit:
                for n in ns:
                    assert n.bound <= limit[nm], "%s: output %s bound %g exceeds %g" % (name, nm, n.bound, limit[nm])
    for n in nodes:
        assert n.bound < (1 << 15), "value may exceed 2^768"
    info = dict(rows=len(levels) + (len(copies) + 3) // 4, mul_rows=sum(1 for ops in levels if ops[0].op == MUL),
                muls=sum(1 for n in nodes if n.op == MUL), lins=sum(1 for n in nodes if n.op == LIN), copies=len(copies))
    prog = Program(name, assemble(levels, copies), n_slots, info)
    prog.out_nodes = out_nodes
    return prog


def build_group(gid, with_mixed=True):
    g = GROUPS[gid]
    progs = {}
    bld, lay, outs = formula_dbl(g)
    progs["dbl"] = compile_formula("dbl", bld, lay, outs, lay.P, ACC_BOUND)
    bld, lay, outs = formula_reduce(g)
    progs["reduce"] = compile_formula("reduce", bld, lay, outs, lay.P, dict(X=2.0, Y=2.0))
    for mixed in ((False, True) if with_mixed else (False,)):
        tag = "madd" if mixed else "add"
        bld, lay, outs = formula_add_head(g, mixed)
        # the head's results stay in scratch for the tail: no pinned outputs
        head = compile_formula(tag + "_head", bld, lay, outs, None, dict(tP=2.0, tR=2.0))
        assert head.info["copies"] == 0
        on = head.out_nodes
        head.test_slots = dict(tP=[n.slot for n in on["tP"]], tR=[n.slot for n in on["tR"]])
        hb = {nm: ([n.slot for n in ns], [n.bound for n in ns]) for nm, ns in on.items()}
        progs[tag + "_head"] = head
        bld, lay, outs2 = formula_add_tail(g, mixed, hb)
        # the head's live scratch slots are inputs of the tail DAG: occupied until their last use
        progs[tag + "_tail"] = compile_formula(tag + "_tail", bld, lay, outs2, lay.P, ACC_BOUND)
    bld, lay, outs = formula_to_projective(g)
    progs["to_proj"] = compile_formula("to_proj", bld, lay, outs, lay.P, dict(X=2.0, Y=2.0, ZZ=2.0))
    bld, lay, outs, base = formula_from_projective(g)
    progs["from_proj"] = compile_formula("from_proj", bld, lay, outs, base)
    return progs


# ---------------------------------------------------------------------------------------------
# exact model of csrc/coop.cuh: values are 8 lanes x 3 limbs
# ---------------------------------------------------------------------------------------------
def split(x):
    assert 0 <= x < (1 << R_BITS)
    return [[(x >> (32 * (LPL * l + i))) & MASK32 for i in range(LPL)] for l in range(LANES)]


def join(v):
    return sum(v[l][i] << (32 * (LPL * l + i)) for l in range(LANES) for i in range(LPL))


def lane96(v3):
    return v3[0] | (v3[1] << 32) | (v3[2] << 64)


def to3(x):
    return [x & MASK32, (x >> 32) & MASK32, (x >> 64) & MASK32]


def resolve(low, carry):
    """low: per-lane 96-bit values, carry: per-lane small carry words out of each lane; returns the clean
    lanes of sum_l (low_l + carry_l 2^96) 2^(96 l) mod 2^768 and the carry out of the top lane - as
    coop.cuh: add the lower neighbour's carry, then one generate / propagate pass over the octet's ballot"""
    r, g, p = [], 0, 0
    for l in range(LANES):
        t = low[l] + (carry[l - 1] if l else 0)
        if t >> 96:
            g |= 1 << l
        t &= MASK96
        if t == MASK96:
            p |= 1 << l
        r.append(t)
    assert g & p == 0
    x, y = p | g, g
    cin = ((x + y) ^ x ^ y)          # bit l: carry into lane l
    top = ((x + y) >> LANES) & 1
    out = [(r[l] + ((cin >> l) & 1)) & MASK96 for l in range(LANES)]
    return out, top + carry[LANES - 1]


class Field:
    def __init__(self, p):
        self.p = p
        self.np96 = (-pow(p, -1, 1 << 96)) % (1 << 96)
        self.one = (1 << R_BITS) % p
        self.kp = [split(p << k) for k in range(MAX_K_LOG + 1)]
        self.pl = split(p)

    def mul(self, a, b):
        """digit-serial Montgomery product in base 2^96, one digit of b per round"""
        n = self.pl
        lo = [0] * LANES          # per-lane 96-bit low part
        hi = [0] * LANES          # per-lane small high word
        for j in range(LANES):
            B = lane96(b[j])
            t = [lo[l] + (hi[l] << 96) + lane96(a[l]) * B for l in range(LANES)]
            M = ((t[0] & MASK96) * self.np96) & MASK96
            t = [t[l] + lane96(n[l]) * M for l in range(LANES)]
            assert all(x < (1 << 224) for x in t)
            assert t[0] & MASK96 == 0
            new_lo, new_hi = [], []
            for l in range(LANES):
                recv = (t[l + 1] & MASK96) if l + 1 < LANES else 0
                s = recv + ((t[l] >> 96) & MASK96)
                new_lo.append(s & MASK96)
                new_hi.append((s >> 96) + (t[l] >> 192))
            lo, hi = new_lo, new_hi
            assert all(h < (1 << 32) for h in hi)
        out, top = resolve(lo, hi)
        assert top == 0, "Montgomery product exceeds 2^768"
        return [to3(x) for x in out]

    def add(self, a, b):
        s = [lane96(a[l]) + lane96(b[l]) for l in range(LANES)]
        out, top = resolve([x & MASK96 for x in s], [x >> 96 for x in s])
        assert top == 0
        return [to3(x) for x in out]

    def sub(self, a, b, k):
        """a + 2^k p - b as a + 2^k p + ~b + 1 (mod 2^768)"""
        kp = self.kp[k]
        s = [lane96(a[l]) + lane96(kp[l]) + (MASK96 ^ lane96(b[l])) + (1 if l == 0 else 0) for l in range(LANES)]
        out, top = resolve([x & MASK96 for x in s], [x >> 96 for x in s])
        assert top == 1, "a + K p - b must not be negative"
        return [to3(x) for x in out]

    def lin(self, terms, k):
        """sum_i coef_i x_i (+ 2^k p when a coefficient is negative), as coop.cuh's fused linear operation:
        every lane sums coef * limbs (complemented limbs for negative coefficients, whose +1s enter at lane
        0) into a wide value, the carries are resolved once; the carries out of 2^768 cancel"""
        has_neg = any(c < 0 for _, c in terms)
        inc = sum(-c for _, c in terms if c < 0)
        s = [0] * LANES
        for x, c in terms:
            for l in range(LANES):
                v = lane96(x[l])
                s[l] += abs(c) * ((MASK96 ^ v) if c < 0 else v)
        if has_neg:
            for l in range(LANES):
                s[l] += lane96(self.kp[k][l])
            s[0] += inc
        assert all(x < (1 << 128) for x in s)
        out, _ = resolve([x & MASK96 for x in s], [x >> 96 for x in s])
        return [to3(x) for x in out]

    def cond_sub_p(self, a):
        """a in [0, 2p) -> a mod p"""
        s = [lane96(a[l]) + (MASK96 ^ lane96(self.pl[l])) + (1 if l == 0 else 0) for l in range(LANES)]
        out, top = resolve([x & MASK96 for x in s], [x >> 96 for x in s])
        return [to3(x) for x in out] if top else a

    def is_zero_mod_p(self, a):
        """for a < 2p"""
        return join(a) in (0, self.p)


class Sim:
    """interpreter of assembled programs over slots holding lane-split values"""

    def __init__(self, field, n_slots):
        self.f = field
        self.slots = [split(0) for _ in range(n_slots)]

    def put(self, slot, x):
        self.slots[slot] = split(x)

    def get(self, slot):
        return join(self.slots[slot])

    def run(self, prog):
        f = self.f
        for row in prog.words:
            results = []
            for o in range(4):
                w0, w1 = row[2 * o], row[2 * o + 1]
                op, d, a, b, k = w0 & 7, (w0 >> 3) & 127, (w0 >> 10) & 127, (w0 >> 17) & 127, (w0 >> 24) & 15
                if op == NOP:
                    continue
                if op == MUL:
                    r = f.mul(self.slots[a], self.slots[b])
                else:
                    slots = (a, b, w1 & 127)
                    terms = []
                    for j in range(3):
                        mag, neg = (w1 >> (7 + 5 * j)) & 15, (w1 >> (11 + 5 * j)) & 1
                        if mag:
                            terms.append((self.slots[slots[j]], -mag if neg else mag))
                    r = f.lin(terms, k)
                    want = sum(c * join(x) for x, c in terms) + ((self.f.p << k) if any(c < 0 for _, c in terms) else 0)
                    assert 0 <= want < (1 << R_BITS) and join(r) == want, "linear operation out of range"
                results.append((d, r))
            for d, r in results:
                self.slots[d] = r


# ---------------------------------------------------------------------------------------------
# emission
# ---------------------------------------------------------------------------------------------
def field_moduli():
    sys.path.insert(0, ROOT)
    import importlib
    params = importlib.import_module("ginger-lib_b200.params")
    # field id 0 = mnt4753::Fq (base field of the MNT4 groups), 1 = mnt6753::Fq
    return {0: params.GROUP_BASE_MODULUS[0], 1: params.GROUP_BASE_MODULUS[2]}


def emit(path):
    mods = field_moduli()
    out = []
    out.append("// GENERATED by tools/gen_coop.py - do not edit.  Micro-programs of the warp-cooperative group law")
    out.append("// (csrc/coop.cuh): rows of four instructions, one per octet of a warp, two 32-bit words each:")
    out.append("//   word 0: op (3 bits: 0 nop, 1 product a * b, 2 linear) | dst << 3 | a << 10 | b << 17 | k << 24;")
    out.append("//   word 1: c | |ca| << 7 | sign a << 11 | |cb| << 12 | sign b << 16 | |cc| << 17 | sign c << 21;")
    out.append("//   linear: dst = ca a + cb b + cc c (+ 2^k p when a coefficient is negative);")
    out.append("//   bit 28 / 29 of a row's first word: the row holds products / linear operations.")
    out.append("#pragma once")
    out.append("namespace g753 {")
    out.append("// Programs live in __constant__ memory behind one accessor type each: the interpreter indexes them with its")
    out.append("// (warp-uniform) row counter, so a row's flags are a uniform register and its branches are uniform.")
    for fid, p in mods.items():
        f = Field(p)
        out.append("// field %d: K p for K = 2^0 .. 2^%d (24 limbs each), -p^-1 mod 2^96" % (fid, MAX_K_LOG))
        rows = []
        for k in range(MAX_K_LOG + 1):
            v = p << k
            rows.append("{" + ", ".join("0x%08xu" % ((v >> (32 * i)) & MASK32) for i in range(24)) + "}")
        out.append("static __device__ const uint32_t COOP_KP_%d[%d][24] = {\n  %s};" % (fid, MAX_K_LOG + 1, ",\n  ".join(rows)))
        out.append("static __device__ const uint32_t COOP_NP96_%d[3] = {%s};" % (
            fid, ", ".join("0x%08xu" % ((f.np96 >> (32 * i)) & MASK32) for i in range(3))))
    summary = []
    for gid in sorted(GROUPS):
        progs = build_group(gid, with_mixed=False)      # the device code uses the full addition only
        nm = GROUPS[gid]["name"]
        slots = max(p.n_slots for p in progs.values())
        for pn, pr in progs.items():
            flat = ", ".join("0x%08xu" % w for row in pr.words for w in row)
            assert all(len(row) == 8 for row in pr.words)
            out.append("// %s %s: %s" % (nm, pn, pr.info))
            cname = "COOP_%s_%s" % (nm.upper(), pn.upper())
            out.append("static __constant__ uint32_t %s[] = {%s};" % (cname, flat))
            out.append("struct %s_P { static constexpr unsigned ROWS = %d; static __device__ __forceinline__ uint32_t w(unsigned i) "
                       "{ return %s[i]; } };" % (cname, len(pr.words), cname))
            summary.append((nm, pn, pr.info))
        for tag in ("add",):
            ts = progs[tag + "_head"].test_slots
            for which in ("tP", "tR"):
                out.append("static __device__ const unsigned char COOP_%s_%s_%s[] = {%s};" % (
                    nm.upper(), tag.upper(), which.upper(), ", ".join(str(x) for x in ts[which])))
        out.append("static const unsigned COOP_%s_SLOTS = %d;" % (nm.upper(), slots))
    out.append("}  // namespace g753")
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")
    return summary


if __name__ == "__main__":
    dst = os.path.join(ROOT, "ginger-lib_b200", "csrc", "coop_programs.inc")
    for nm, pn, info in emit(dst):
        print("%-5s %-10s %s" % (nm, pn, info))
    print("wrote", dst)
