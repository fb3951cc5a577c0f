"""CPU oracle (Python big-int) for ginger-lib's Groth16 prover hot path on MNT4-753 / MNT6-753.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or executed by
the product (ginger-lib_b200/); only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may use it, and only as the checker.

This file is the *definitional* oracle: it computes in canonical integers mod p and
affine coordinates with a true modular inverse, so that every result is the unique
canonical answer the reference must also produce (normalised affine points,
fully-reduced field elements).  The faster C++ restatement of the reference's
*algorithms* (Montgomery u64x12 limbs, homogeneous projective formulas, unsigned
window Pippenger, serial/parallel radix-2 FFT) lives in oracle/ref753.cpp and is
cross-checked against this file.

Parity pinning: tests/test_oracle_kat.py replays the reference's own known-answer
tests (tests/golden/reference_kat.json, extracted by tests/golden/make_golden.py)
against this module: field add/sub/mul/square in raw Montgomery limbs
(algebra/src/fields/mnt4753/tests.rs:271-730, mnt6753/tests.rs), Fq2/Fq3 KATs,
G1/G2 add/double/scalar-mul/normalisation KATs (algebra/src/curves/mnt4753/tests.rs:624-1596,
mnt6753/tests.rs:805-2028) and every constant in fields/*/fq.rs, curves/*/g{1,2}.rs.
There is no 753-bit MSM/FFT vector in the reference; at that
level the answer is pinned by definition (MSM = sum s_i P_i, FFT = DFT over <omega>).
The mixed-radix DFT has no reference counterpart at all: "parity unpinned".

Reference files followed (relative to /root/reference):
  algebra/src/fields/models/fp_768.rs      Montgomery R = 2^768, from_repr/into_repr :627-667
  algebra/src/fields/models/fp2.rs:387-401 Fp2 mul, fp3.rs:451-478 Fp3 mul
  algebra/src/curves/models/short_weierstrass_projective.rs:444-519,574-678
  algebra/src/msm/variable_base.rs:10-90   Pippenger semantics (zip-truncation, 0/1 scalars)
  algebra/src/fft/domain.rs:65-179,305-358 EvaluationDomain, serial_fft
"""

# --------------------------------------------------------------------------------------
# prime fields
# --------------------------------------------------------------------------------------

LIMBS64 = 12
R_BITS = 768
R_INT = 1 << R_BITS
MASK64 = (1 << 64) - 1


class PrimeField:
    """One of the two 753-bit prime fields of the MNT4/MNT6 cycle.

    mnt4753::Fq == mnt6753::Fr  (algebra/src/fields/mnt4753/fq.rs:18-31, mnt6753/fr.rs:1)
    mnt6753::Fq == mnt4753::Fr  (algebra/src/fields/mnt6753/fq.rs:17-30, mnt4753/fr.rs:1)
    Constants are derived here from p alone and compared with the reference's literals in
    tests/test_oracle_kat.py::test_field_constants.
    """

    def __init__(self, name, p, two_adicity):
        self.name = name
        self.p = p
        self.bits = p.bit_length()
        self.R = R_INT % p                         # FpParameters::R
        self.R2 = (R_INT * R_INT) % p              # FpParameters::R2
        self.Rinv = pow(R_INT, -1, p)
        self.inv64 = (-pow(p, -1, 1 << 64)) % (1 << 64)   # FpParameters::INV
        self.inv32 = self.inv64 & 0xFFFFFFFF
        self.generator = 17                        # FpParameters::GENERATOR (canonical)
        self.two_adicity = two_adicity
        self.t = (p - 1) >> two_adicity
        assert (p - 1) == self.t << two_adicity and self.t & 1
        self.root_of_unity = pow(self.generator, self.t, p)   # 2^s-th primitive root

    # representation changes -----------------------------------------------------------
    def to_mont(self, x):
        return (x * self.R) % self.p

    def from_mont(self, xm):
        return (xm * self.Rinv) % self.p

    def mont_mul(self, am, bm):
        """Fp768::mul_assign on raw Montgomery limbs: a*b*R^-1 mod p (fp_768.rs:1009-1185)."""
        return (am * bm * self.Rinv) % self.p

    # canonical arithmetic --------------------------------------------------------------
    def add(self, a, b):
        return (a + b) % self.p

    def sub(self, a, b):
        return (a - b) % self.p

    def mul(self, a, b):
        return (a * b) % self.p

    def neg(self, a):
        return (-a) % self.p

    def inv(self, a):
        return pow(a, -1, self.p)

    def sqrt(self, a):
        """Any square root or None (Tonelli-Shanks); used only by input generators."""
        p = self.p
        a %= p
        if a == 0:
            return 0
        if pow(a, (p - 1) // 2, p) != 1:
            return None
        s, t = self.two_adicity, self.t
        z = self.root_of_unity
        x = pow(a, (t + 1) // 2, p)
        b = pow(a, t, p)
        m = s
        while b != 1:
            i, b2 = 0, b
            while b2 != 1:
                b2 = b2 * b2 % p
                i += 1
            g = pow(z, 1 << (m - i - 1), p)
            x = x * g % p
            z = g * g % p
            b = b * z % p
            m = i
        return x


P_MNT4_FQ = int(
    "1c4c62d92c41110229022eee2cdadb7f997505b8fafed5eb7e8f96c97d87307fdb925e8a0ed8d99d124d9a15af79db"
    "117e776f218059db80f0da5cb537e38685acce9767254a4638810719ac425f0e39d54522cdd119f5e9063de245e8001", 16)
P_MNT6_FQ = int(
    "1c4c62d92c41110229022eee2cdadb7f997505b8fafed5eb7e8f96c97d87307fdb925e8a0ed8d99d124d9a15af79db"
    "26c5c28c859a99b3eebca9429212636b9dff97634993aa4d6c381bc3f0057974ea099170fa13a4fd90776e240000001", 16)

MNT4_FQ = PrimeField("mnt4753_fq", P_MNT4_FQ, 15)   # == mnt6753::Fr
MNT6_FQ = PrimeField("mnt6753_fq", P_MNT6_FQ, 30)   # == mnt4753::Fr
MNT4_FR = MNT6_FQ
MNT6_FR = MNT4_FQ
FIELDS = {"mnt4753_fq": MNT4_FQ, "mnt6753_fq": MNT6_FQ, "mnt4753_fr": MNT4_FR, "mnt6753_fr": MNT6_FR}


# limb / byte helpers ------------------------------------------------------------------

def int_to_limbs64(x):
    return [(x >> (64 * i)) & MASK64 for i in range(LIMBS64)]


def limbs64_to_int(l):
    return sum(int(v) << (64 * i) for i, v in enumerate(l))


def int_to_bytes96(x):
    """ToBytes of a BigInteger768: 12 little-endian u64 (bytes.rs:70-78)."""
    return x.to_bytes(96, "little")


def bytes96_to_int(b):
    return int.from_bytes(b, "little")


# --------------------------------------------------------------------------------------
# extension fields:  elements are tuples of canonical ints
# --------------------------------------------------------------------------------------

class ExtField:
    """Fq[u]/(u^k - nonresidue) for k in {1, 2, 3}; k == 1 is the prime field itself.

    Fq2: nonresidue 13 (algebra/src/fields/mnt4753/fq2.rs:17-32, mul fp2.rs:387-401)
    Fq3: nonresidue 11 (algebra/src/fields/mnt6753/fq3.rs:17-31, mul fp3.rs:451-478)
    """

    def __init__(self, base, k, nonresidue):
        self.base = base
        self.k = k
        self.nr = nonresidue
        self.p = base.p

    def zero(self):
        return (0,) * self.k

    def one(self):
        return (1,) + (0,) * (self.k - 1)

    def from_int(self, x):
        return (x % self.p,) + (0,) * (self.k - 1)

    def is_zero(self, a):
        return all(c == 0 for c in a)

    def add(self, a, b):
        return tuple((x + y) % self.p for x, y in zip(a, b))

    def sub(self, a, b):
        return tuple((x - y) % self.p for x, y in zip(a, b))

    def neg(self, a):
        return tuple((-x) % self.p for x in a)

    def mul(self, a, b):
        p, k, nr = self.p, self.k, self.nr
        if k == 1:
            return ((a[0] * b[0]) % p,)
        res = [0] * (2 * k - 1)
        for i in range(k):
            for j in range(k):
                res[i + j] += a[i] * b[j]
        for i in range(2 * k - 2, k - 1, -1):
            res[i - k] += nr * res[i]
        return tuple(r % p for r in res[:k])

    def sqr(self, a):
        return self.mul(a, a)

    def mul_base(self, a, s):
        return tuple((x * s) % self.p for x in a)

    def inv(self, a):
        p, k, nr = self.p, self.k, self.nr
        if k == 1:
            return (pow(a[0], -1, p),)
        if k == 2:
            # 1/(a0 + a1 u) = (a0 - a1 u)/(a0^2 - nr a1^2)
            d = pow((a[0] * a[0] - nr * a[1] * a[1]) % p, -1, p)
            return ((a[0] * d) % p, (-a[1] * d) % p)
        # k == 3: solve by the adjugate of the multiplication matrix
        a0, a1, a2 = a
        t0 = (a0 * a0 - nr * a1 * a2) % p
        t1 = (nr * a2 * a2 - a0 * a1) % p
        t2 = (a1 * a1 - a0 * a2) % p
        d = pow((a0 * t0 + nr * (a2 * t1 + a1 * t2)) % p, -1, p)
        return ((t0 * d) % p, (t1 * d) % p, (t2 * d) % p)


FQ_MNT4 = ExtField(MNT4_FQ, 1, None)
FQ2_MNT4 = ExtField(MNT4_FQ, 2, 13)
FQ_MNT6 = ExtField(MNT6_FQ, 1, None)
FQ3_MNT6 = ExtField(MNT6_FQ, 3, 11)


# --------------------------------------------------------------------------------------
# short Weierstrass curves  y^2 = x^3 + a x + b  over an ExtField; affine, None == infinity
# --------------------------------------------------------------------------------------

class Curve:
    def __init__(self, name, F, a, b, scalar_field, gen=None):
        self.name = name
        self.F = F
        self.a = a
        self.b = b
        self.r = scalar_field.p          # prime subgroup order
        self.gen = gen

    def on_curve(self, P):
        if P is None:
            return True
        F = self.F
        x, y = P
        rhs = F.add(F.add(F.mul(F.sqr(x), x), F.mul(self.a, x)), self.b)
        return F.sqr(y) == rhs

    def neg(self, P):
        if P is None:
            return None
        return (P[0], self.F.neg(P[1]))

    def double(self, P):
        if P is None:
            return None
        F = self.F
        x, y = P
        if F.is_zero(y):
            return None
        xx = F.sqr(x)
        num = F.add(F.add(F.add(xx, xx), xx), self.a)
        lam = F.mul(num, F.inv(F.add(y, y)))
        x3 = F.sub(F.sub(F.sqr(lam), x), x)
        y3 = F.sub(F.mul(lam, F.sub(x, x3)), y)
        return (x3, y3)

    def add(self, P, Q):
        if P is None:
            return Q
        if Q is None:
            return P
        F = self.F
        if P[0] == Q[0]:
            if P[1] == Q[1]:
                return self.double(P)
            return None
        lam = F.mul(F.sub(Q[1], P[1]), F.inv(F.sub(Q[0], P[0])))
        x3 = F.sub(F.sub(F.sqr(lam), P[0]), Q[0])
        y3 = F.sub(F.mul(lam, F.sub(P[0], x3)), P[1])
        return (x3, y3)

    def mul(self, P, k):
        acc = None
        if k < 0:
            P, k = self.neg(P), -k
        for bit in bin(k)[2:] if k else "":
            acc = self.double(acc)
            if bit == "1":
                acc = self.add(acc, P)
        return acc

    def from_projective(self, X, Y, Z):
        """GroupProjective -> GroupAffine: homogeneous x = X/Z, y = Y/Z
        (short_weierstrass_projective.rs:663-678)."""
        F = self.F
        if F.is_zero(Z):
            return None
        zi = F.inv(Z)
        return (F.mul(X, zi), F.mul(Y, zi))


_B4 = int("1373684a8c9dcae7a016ac5d7748d3313cd8e39051c596560835df0c9e50a5b59b882a92c78dc537e51a16703ec9855c"
          "77fc3d8bb21c8d68bb8cfb9db4b8c8fba773111c36c8b1b4e8f1ece940ef9eaad265458e06372009c9a0491678ef4", 16)
_B6 = int("7da285e70863c79d56446237ce2e1468d14ae9bb64b2bb01b10e60a5d5dfe0a25714b7985993f62f03b22a9a3c737a1a"
          "1e0fcf2c43d7bf847957c34cca1e3585f9a80a95f401867c4e80f4747fde5aba7505ba6fcf2485540b13dfc8468a", 16)

# G1:  MNT4 y^2 = x^3 + 2x + b   (algebra/src/curves/mnt4753/g1.rs:18-50)
#      MNT6 y^2 = x^3 + 11x + b  (algebra/src/curves/mnt6753/g1.rs:18-52)
# G2:  MNT4 twist over Fq2: a' = a*13 = (26, 0), b' = b*13*u = (0, 13 b)   (mnt4753/g2.rs:20-75)
#      MNT6 twist over Fq3: a' = a*u^2 = (0, 0, 11), b' = b*11 = (11 b, 0, 0)   (mnt6753/g2.rs:19-95)
MNT4_G1 = Curve("mnt4753_g1", FQ_MNT4, (2,), (_B4,), MNT4_FR)
MNT6_G1 = Curve("mnt6753_g1", FQ_MNT6, (11,), (_B6,), MNT6_FR)
MNT4_G2 = Curve("mnt4753_g2", FQ2_MNT4, (26, 0), (0, (13 * _B4) % P_MNT4_FQ), MNT4_FR)
MNT6_G2 = Curve("mnt6753_g2", FQ3_MNT6, (0, 0, 11), ((11 * _B6) % P_MNT6_FQ, 0, 0), MNT6_FR)
CURVES = {c.name: c for c in (MNT4_G1, MNT4_G2, MNT6_G1, MNT6_G2)}


# --------------------------------------------------------------------------------------
# MSM
# --------------------------------------------------------------------------------------

def msm_naive(curve, bases, scalars):
    """sum_i s_i * P_i over zip(bases, scalars) - the definition the reference's own MSM
    test checks against (algebra/src/msm/variable_base.rs:102-131)."""
    acc = None
    for P, s in zip(bases, scalars):
        acc = curve.add(acc, curve.mul(P, s))
    return acc


def ref_window_size(n_scalars):
    """c of variable_base.rs:14-18 (uses scalars.len())."""
    import math
    if n_scalars < 32:
        return 3
    return int(math.ceil(2.0 / 3.0 * math.log2(float(n_scalars)) + 2.0))


def msm_pippenger_ref(curve, bases, scalars, num_bits=753):
    """The reference's unsigned-window bucket method, step for step
    (variable_base.rs:10-83), in affine arithmetic.  Semantics preserved: zip-truncation,
    zero scalars skipped, scalar == 1 handled once in window 0, infinity bases no-ops."""
    c = ref_window_size(len(scalars))
    window_sums = []
    for w_start in range(0, num_bits, c):
        res = None
        buckets = [None] * ((1 << c) - 1)
        for s, P in zip(scalars, bases):
            if s == 0:
                continue
            if s == 1:
                if w_start == 0:
                    res = curve.add(res, P)
                continue
            d = (s >> w_start) % (1 << c)
            if d != 0:
                buckets[d - 1] = curve.add(buckets[d - 1], P)
        running = None
        for b in reversed(buckets):
            running = curve.add(running, b)
            res = curve.add(res, running)
        window_sums.append(res)
    total = None
    for s in reversed(window_sums[1:]):
        total = curve.add(total, s)
        for _ in range(c):
            total = curve.double(total)
    return curve.add(total, window_sums[0])


# --------------------------------------------------------------------------------------
# EvaluationDomain / FFT   (algebra/src/fft/domain.rs)
# --------------------------------------------------------------------------------------

class EvaluationDomain:
    """Mirror of EvaluationDomain<F>::new (domain.rs:65-94) in canonical integers."""

    def __init__(self, field, num_coeffs):
        size = 1
        while size < num_coeffs:
            size <<= 1
        log_n = size.bit_length() - 1
        if log_n >= field.two_adicity:            # domain.rs:70-72
            raise ValueError("domain too large for this field")
        self.field = field
        self.size = size
        self.log_size_of_group = log_n
        g = field.root_of_unity
        for _ in range(log_n, field.two_adicity):
            g = g * g % field.p
        self.group_gen = g
        self.group_gen_inv = pow(g, -1, field.p)
        self.size_inv = pow(size, -1, field.p)
        self.generator = field.generator
        self.generator_inv = pow(field.generator, -1, field.p)

    @staticmethod
    def try_new(field, num_coeffs):
        try:
            return EvaluationDomain(field, num_coeffs)
        except ValueError:
            return None

    def _resize(self, v):
        v = list(v)[:self.size]                   # Vec::resize truncates or zero-pads
        return v + [0] * (self.size - len(v))

    def fft(self, coeffs):
        return serial_fft(self._resize(coeffs), self.group_gen, self.log_size_of_group, self.field.p)

    def ifft(self, evals):
        p = self.field.p
        out = serial_fft(self._resize(evals), self.group_gen_inv, self.log_size_of_group, p)
        return [x * self.size_inv % p for x in out]

    def coset_fft(self, coeffs):
        # distribute_powers is applied to the *unresized* vector (domain.rs:163-166, 140-152)
        return self.fft(distribute_powers(coeffs, self.generator, self.field.p))

    def coset_ifft(self, evals):
        return distribute_powers(self.ifft(evals), self.generator_inv, self.field.p)


def distribute_powers(v, g, p):
    out, u = [], 1
    for x in v:
        out.append(x * u % p)
        u = u * g % p
    return out


def serial_fft(a, omega, log_n, p):
    """domain.rs:315-358: bit-reversal then log_n DIT stages; natural-order output."""
    n = len(a)
    assert n == 1 << log_n
    a = list(a)
    for k in range(n):
        rk = int(format(k, "0%db" % log_n)[::-1], 2) if log_n else 0
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    m = 1
    for _ in range(log_n):
        w_m = pow(omega, n // (2 * m), p)
        for k in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                t = a[k + j + m] * w % p
                a[k + j + m] = (a[k + j] - t) % p
                a[k + j] = (a[k + j] + t) % p
                w = w * w_m % p
        m *= 2
    return a


def dft_naive(a, omega, p):
    """out[i] = sum_j a[j] omega^(ij): the definition (also the only oracle for the
    mixed-radix sizes, which the reference does not implement: parity unpinned)."""
    n = len(a)
    pw = [1] * n
    for i in range(1, n):
        pw[i] = pw[i - 1] * omega % p
    return [sum(a[j] * pw[(i * j) % n] for j in range(n)) % p for i in range(n)]


def mixed_radix_omega(field, N):
    assert (field.p - 1) % N == 0
    return pow(field.generator, (field.p - 1) // N, field.p)


# --------------------------------------------------------------------------------------
# deterministic input generators: SplitMix64 streams
# --------------------------------------------------------------------------------------

class SplitMix64:
    def __init__(self, seed):
        self.s = seed & MASK64

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)


def random_field_element(rng, field):
    """12 random words, top REPR_SHAVE_BITS=15 bits masked, rejected if >= p - the sampling
    rule of algebra/src/fields/macros.rs:11-28."""
    while True:
        x = 0
        for i in range(LIMBS64):
            x |= rng.next() << (64 * i)
        x &= (1 << 753) - 1
        if x < field.p:
            return x


def random_g1_point(rng, curve):
    """x-sampling; valid for the cofactor-1 G1 groups (curves/mnt4753/g1.rs:53)."""
    F = curve.F
    fld = F.base
    while True:
        x = random_field_element(rng, fld)
        rhs = (x * x * x + curve.a[0] * x + curve.b[0]) % fld.p
        y = fld.sqrt(rhs)
        if y is None:
            continue
        if y & 1:
            y = fld.p - y
        return ((x,), (y,))


# --------------------------------------------------------------------------------------
# Groth16 prover algebra  (proof-systems/src/groth16/{r1cs_to_qap.rs, prover.rs, mod.rs})
# --------------------------------------------------------------------------------------

def witness_map(field, a_evals, b_evals, c_evals, d1, d2, d3):
    """R1CStoQAP::witness_map from the evaluated constraints onwards (r1cs_to_qap.rs:121-166), in
    canonical integers.  a_evals / b_evals / c_evals are the domain_size evaluation vectors the
    reference fills at :111-119 and :141-152 (padding included).  Returns h, domain_size + 1 values.

    The reference initialises h to zeros and then MULTIPLIES h_i by (d2 a_i + d1 b_i) (:125-130), so
    those entries stay zero; restated literally."""
    p = field.p
    n = len(a_evals)
    assert len(b_evals) == n and len(c_evals) == n
    dom = EvaluationDomain(field, n)
    assert dom.size == n
    a = dom.ifft(a_evals)                                    # :121
    b = dom.ifft(b_evals)                                    # :122
    h = [0 * ((d2 * ai + d1 * bi) % p) % p for ai, bi in zip(a, b)]   # :124-130  (0 *= ...)
    d1d2 = d1 * d2 % p
    h[0] = (h[0] - d3) % p                                   # :131
    h[0] = (h[0] - d1d2) % p                                 # :133
    h.append(d1d2)                                           # :134
    a = dom.coset_fft(a)                                     # :136
    b = dom.coset_fft(b)                                     # :137
    ab = [x * y % p for x, y in zip(a, b)]                   # :139  mul_polynomials_in_evaluation_domain
    c = dom.coset_fft(dom.ifft(c_evals))                     # :154-155
    ab = [(x - y) % p for x, y in zip(ab, c)]                # :157-159
    z_inv = pow((pow(dom.generator, n, p) - 1) % p, -1, p)   # domain.rs:245-256
    ab = [x * z_inv % p for x in ab]
    ab = dom.coset_ifft(ab)                                  # :162
    for i in range(n - 1):                                   # :164-167
        h[i] = (h[i] + ab[i]) % p
    return h


class Groth16Key:
    """The fields of groth16::Parameters the prover reads (mod.rs:313-371), as affine oracle points
    (None = infinity).  g1 / g2 are the curves of E::G1Affine / E::G2Affine."""

    def __init__(self, g1, g2, alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2, a_query, b_g1_query,
                 b_g2_query, h_query, l_query):
        self.g1, self.g2 = g1, g2
        self.alpha_g1, self.beta_g1, self.beta_g2 = alpha_g1, beta_g1, beta_g2
        self.delta_g1, self.delta_g2 = delta_g1, delta_g2
        self.a_query, self.b_g1_query, self.b_g2_query = a_query, b_g1_query, b_g2_query
        self.h_query, self.l_query = h_query, l_query


def groth16_create_proof(key, num_inputs, full_assignment, h, r, s, msm=msm_naive):
    """prover.rs:241-345 from the witness map's outputs onwards: the nine MSMs (zip-truncating, as
    variable_base.rs:36) and the assembly of (A, B, C), normalised to affine.
    full_assignment: canonical ints, inputs first (index 0 is the constant one); h: canonical ints."""
    g1, g2 = key.g1, key.g2
    inp = full_assignment[1:num_inputs]                       # :241-246
    aux = full_assignment[num_inputs:]                        # :248-253
    h_in, h_aux = h[:num_inputs], h[num_inputs:]              # :256-267

    def acc(curve, *pts):
        t = None
        for q in pts:
            t = curve.add(t, q)
        return t

    g_a = acc(g1, g1.mul(key.delta_g1, r), key.a_query[0], msm(g1, key.a_query[1:num_inputs], inp),
              msm(g1, key.a_query[num_inputs:], aux), key.alpha_g1)                      # :270-283
    g1_b = acc(g1, g1.mul(key.delta_g1, s), key.b_g1_query[0], msm(g1, key.b_g1_query[1:num_inputs], inp),
               msm(g1, key.b_g1_query[num_inputs:], aux), key.beta_g1)                   # :286-299
    g2_b = acc(g2, g2.mul(key.delta_g2, s), key.b_g2_query[0], msm(g2, key.b_g2_query[1:num_inputs], inp),
               msm(g2, key.b_g2_query[num_inputs:], aux), key.beta_g2)                   # :302-315
    rs_delta = g1.mul(g1.mul(key.delta_g1, r), s)
    g_c = acc(g1, g1.mul(g_a, s), g1.mul(g1_b, r), g1.neg(rs_delta), msm(g1, key.l_query, aux),
              msm(g1, key.h_query[:num_inputs], h_in), msm(g1, key.h_query[num_inputs:], h_aux))  # :318-337
    return g_a, g2_b, g_c


# --------------------------------------------------------------------------------------
# GM17 prover (proof-systems/src/gm17): the SAP witness map and create_proof, restated
# --------------------------------------------------------------------------------------

def sap_witness_map(field, a_evals, c_evals, d1, d2):
    """R1CStoSAP::witness_map from the evaluated constraints onwards (gm17/r1cs_to_sap.rs:191-245).
    a_evals / c_evals are the domain_size vectors the reference fills at :163-186 and :205-232
    (padding included).  Returns h, domain_size + 1 canonical integers."""
    p = field.p
    n = len(a_evals)
    assert len(c_evals) == n
    dom = EvaluationDomain(field, n)
    assert dom.size == n
    a = dom.ifft(a_evals)                                    # :191
    d1_double = 2 * d1 % p                                   # :193
    h = [d1_double * ai % p for ai in a]                     # :194-195
    d1d1 = d1 * d1 % p
    h[0] = (h[0] - d2) % p                                   # :196
    h[0] = (h[0] - d1d1) % p                                 # :198
    h.append(d1d1)                                           # :199
    a = dom.coset_fft(a)                                     # :201
    aa = [x * x % p for x in a]                              # :203
    c = dom.coset_fft(dom.ifft(c_evals))                     # :234-235
    aa = [(x - y) % p for x, y in zip(aa, c)]                # :237
    z_inv = pow((pow(dom.generator, n, p) - 1) % p, -1, p)   # :239, domain.rs:245-256
    aa = dom.coset_ifft([x * z_inv % p for x in aa])         # :240
    for i in range(n - 1):                                   # :242-245 (h[..domain_size - 1])
        h[i] = (h[i] + aa[i]) % p
    return h


class GM17Key:
    """The fields of gm17::Parameters the prover reads (gm17/mod.rs:138-149), as affine oracle points
    (None = infinity)."""

    def __init__(self, g1, g2, a_query, b_query, c_query_1, c_query_2, g_gamma_z, h_gamma_z, g_ab_gamma_z,
                 g_gamma2_z2, g_gamma2_z_t):
        self.g1, self.g2 = g1, g2
        self.a_query, self.b_query, self.c_query_1, self.c_query_2 = a_query, b_query, c_query_1, c_query_2
        self.g_gamma_z, self.h_gamma_z, self.g_ab_gamma_z = g_gamma_z, h_gamma_z, g_ab_gamma_z
        self.g_gamma2_z2, self.g_gamma2_z_t = g_gamma2_z2, g_gamma2_z_t


def gm17_create_proof(key, num_inputs, full_assignment, h, d1, d2, r, fr, msm=msm_naive):
    """gm17/prover.rs:235-354 from the witness map's outputs onwards.  full_assignment: the SAP
    assignment (inputs, aux and the extra variables of r1cs_to_sap.rs:127-151), canonical ints;
    h: canonical ints; fr: the scalar field (r^2, 2 d1 r are field products, :299-301)."""
    g1, g2, p = key.g1, key.g2, fr.p
    ni = num_inputs
    inp, aux = full_assignment[1:ni], full_assignment[ni:]    # :235-247
    h_in, h_aux = h[:ni], h[ni:]                              # :250-261

    def acc(curve, *pts):
        t = None
        for q in pts:
            t = curve.add(t, q)
        return t

    g_a = acc(g1, g1.mul(key.g_gamma_z, r), key.a_query[0], g1.mul(key.g_gamma_z, d1),
              msm(g1, key.a_query[1:ni], inp), msm(g1, key.a_query[ni:], aux))            # :264-278
    g_b = acc(g2, g2.mul(key.h_gamma_z, r), key.b_query[0], g2.mul(key.h_gamma_z, d1),
              msm(g2, key.b_query[1:ni], inp), msm(g2, key.b_query[ni:], aux))            # :280-296
    r_2, r2 = 2 * r % p, r * r % p                            # :299-301
    d1_r_2 = d1 * r_2 % p
    c1_acc = msm(g1, key.c_query_1, aux)                      # :303-306 (get_c_query_1(0))
    c2_acc = g1.add(msm(g1, key.c_query_2[1:ni], inp), msm(g1, key.c_query_2[ni:], aux))  # :308-315
    g_acc = g1.add(msm(g1, key.g_gamma2_z_t[:ni], h_in), msm(g1, key.g_gamma2_z_t[ni:], h_aux))  # :317-325
    g_c = acc(g1, c1_acc, g1.mul(key.g_gamma2_z2, r2), g1.mul(key.g_ab_gamma_z, r), g1.mul(key.g_ab_gamma_z, d1),
              g1.mul(key.c_query_2[0], r), g1.mul(key.g_gamma2_z2, d1_r_2), g1.mul(c2_acc, r),
              g1.mul(key.g_gamma2_z_t[0], d2), g_acc)                                    # :327-344
    return g_a, g_b, g_c
