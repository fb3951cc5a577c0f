"""TEST INFRASTRUCTURE ONLY: the ate pairing of MNT4-753 / MNT6-753 and the Groth16 key generator / verifier,
restated from the reference in plain Python integers - the INDEPENDENT arbiter of a proof (the reference's
own acceptance test is generate -> prove -> verify, proof-systems/src/groth16/test.rs:216-301).  Nothing in
the product imports this file.

Restated, with the reference lines followed (relative to /root/reference):
  algebra/src/curves/models/mnt4/mod.rs:84-151, mnt6/mod.rs (twin)   ate_precompute_g1 / ate_precompute_g2:
        affine doubling / addition steps on the twist along the signed digits (NAF) of the loop count,
        recording (r_y, gamma, gamma_x) per step
  .../mnt4/mod.rs:154-222                                            ate_miller_loop: f <- f^2 * g_RR(P),
        f <- f * g_RQ(P), line values  y_P twist^2 + (gamma_x - gamma twist x_P -+ y) Y; unitary inverse
        when the loop count is negative
  .../mnt4/mod.rs:224-275, mnt6/mod.rs:224-270                       final_exponentiation: first chunk
        f^(q^2 - 1) (MNT4) / f^((q^3 - 1)(q + 1)) (MNT6) by Frobenius maps, last chunk f^(m_1 q + m_0)
  proof-systems/src/groth16/verifier.rs:18-44                        verify_proof
  proof-systems/src/groth16/generator.rs:150-330, r1cs_to_qap.rs:13-67   generate_parameters (toxic waste
        given), instance_map_with_evaluation
  proof-systems/src/groth16/examples/snark-scalability/constraints.rs:19-91   the benchmark circuit's matrices

Representation: the embedding field F_q^k (k = 4 / 6) is ONE polynomial ring F_q[W] / (W^k - alpha) - the
reference's towers F_q2[Y] / (Y^2 - X) over F_q[X] / (X^2 - alpha) (resp. F_q3) are the same field with
X = W^2, Y = W: a tower element ((a0, a1), (b0, b1)) is a0 + b0 W + a1 W^2 + b1 W^3.  The reference's
sparse products (mul_by_023, cyclotomic_exp) compute the same values as the plain operations used here.

Pinning: tests/test_oracle_pairing.py checks the loop count, its NAF and the final-exponent split derived
here against the reference's literals (tests/golden/reference_pairing.json, reference_params.json),
bilinearity, non-degeneracy and e(P, Q)^r = 1 on the reference's generators.
"""
from oracle import g753 as O


class Fqk:
    """F_q[W] / (W^k - alpha), elements as length-k lists of ints"""

    def __init__(self, p, k, alpha):
        self.p, self.k, self.alpha = p, k, alpha
        self.one = [1] + [0] * (k - 1)
        self.zero = [0] * k

    def mul(self, a, b):
        p, k, al = self.p, self.k, self.alpha
        t = [0] * (2 * k - 1)
        for i, x in enumerate(a):
            if x:
                for j, y in enumerate(b):
                    t[i + j] += x * y
        return [(t[i] + (al * t[i + k] if i + k < 2 * k - 1 else 0)) % p for i in range(k)]

    def sqr(self, a):
        return self.mul(a, a)

    def pow(self, a, e):
        r = self.one
        for bit in bin(e)[2:]:
            r = self.sqr(r)
            if bit == "1":
                r = self.mul(r, a)
        return r

    def inv(self, a):
        # a^-1 = a^(q^k - 2); small k, test-only: the generic power is fast enough
        return self.pow(a, self.p ** self.k - 2)

    def frobenius(self, a, power):
        """x -> x^(q^power): W^(q^power) = W * alpha^((q^power - 1) / k), coefficient-wise"""
        p, k = self.p, self.k
        c = pow(self.alpha, (p ** power - 1) // k, p)
        out, cj = [], 1
        for j in range(k):
            out.append(a[j] * cj % p)
            cj = cj * c % p
        return out

    def conj_half(self, a):
        """the unitary inverse of the reference: (c0, c1) -> (c0, -c1) over the degree-k/2 subfield = negate
        the odd powers of W"""
        return [(-x) % self.p if j & 1 else x for j, x in enumerate(a)]


def naf(n):
    """non-adjacent form, least significant digit first"""
    out = []
    while n:
        if n & 1:
            d = 2 - (n % 4)
            n -= d
        else:
            d = 0
        out.append(d)
        n >>= 1
    return out


class Engine:
    """one pairing engine of the cycle"""

    def __init__(self, name, g1, g2, fr):
        self.name, self.g1, self.g2, self.fr = name, g1, g2, fr
        self.p = g1.F.base.p
        self.r = g1.r
        self.ke = g2.F.k                     # degree of the twist field: 2 (MNT4) / 3 (MNT6)
        self.k = 2 * self.ke                 # embedding degree
        self.alpha = g2.F.nr
        self.Fk = Fqk(self.p, self.k, self.alpha)
        self.Fe = g2.F
        # trace of Frobenius: #E(F_q) = q + 1 - t = r (cofactor one); loop count |t - 1|
        t = self.p + 1 - self.r
        self.loop_count = abs(t - 1)
        self.loop_count_neg = (t - 1) < 0
        digits = naf(self.loop_count)
        self.wnaf = digits[:-1]             # without the most significant digit (always 1), LSB first
        # (q^k - 1) / r = first chunk * (m_1 q + m_0), |m_0| <= q / 2
        if self.ke == 2:
            last = (self.p ** 2 + 1) // self.r
            assert (self.p ** 2 + 1) % self.r == 0
        else:
            last = (self.p ** 2 - self.p + 1) // self.r
            assert (self.p ** 2 - self.p + 1) % self.r == 0
        m1, m0 = divmod(last, self.p)
        if m0 > self.p // 2:
            m1, m0 = m1 + 1, m0 - self.p
        self.m1, self.m0 = m1, m0
        # twist = X: the element (0, 1[, 0]) of the twist field; the twist curve's a' = a * twist^2
        self.twist = tuple([0, 1] + [0] * (self.ke - 2))
        self.twist_sq = self.Fe.sqr(self.twist)

    # ---- tower <-> F_q[W] -----------------------------------------------------------------------------
    def embed(self, c0, c1):
        """(c0, c1) with c0, c1 in the twist field, c0 + c1 Y  ->  coefficients in W (X = W^2, Y = W)"""
        out = [0] * self.k
        for j in range(self.ke):
            out[2 * j] = c0[j] % self.p
            out[2 * j + 1] = c1[j] % self.p
        return out

    # ---- mnt4/mod.rs:84-151 ---------------------------------------------------------------------------
    def precompute_g2(self, Q):
        Fe = self.Fe
        a2 = self.g2.a
        coeffs = []
        sx, sy = Q
        for n in reversed(self.wnaf):
            sx2 = Fe.sqr(sx)
            num = Fe.add(Fe.add(Fe.add(sx2, sx2), sx2), a2)
            gamma = Fe.mul(num, Fe.inv(Fe.add(sy, sy)))
            gamma_x = Fe.mul(gamma, sx)
            nx = Fe.sub(Fe.sqr(gamma), Fe.add(sx, sx))
            ny = Fe.sub(Fe.mul(gamma, Fe.sub(sx, nx)), sy)
            coeffs.append((sy, gamma, gamma_x))
            sx, sy = nx, ny
            if n != 0:
                inv = Fe.inv(Fe.sub(sx, Q[0]))
                num = Fe.sub(sy, Q[1]) if n > 0 else Fe.add(sy, Q[1])
                gamma = Fe.mul(num, inv)
                gamma_x = Fe.mul(gamma, Q[0])
                nx = Fe.sub(Fe.sqr(gamma), Fe.add(sx, Q[0]))
                ny = Fe.sub(Fe.mul(gamma, Fe.sub(sx, nx)), sy)
                coeffs.append((sy, gamma, gamma_x))
                sx, sy = nx, ny
        return coeffs

    # ---- mnt4/mod.rs:154-222 ----------------------------------------------------------------------------
    def miller_loop(self, P, Q):
        Fe, Fk = self.Fe, self.Fk
        px, py = P[0][0], P[1][0]
        py_twist_sq = tuple(c * py % self.p for c in self.twist_sq)
        coeffs = self.precompute_g2(Q)
        f = Fk.one
        idx = 0
        for n in reversed(self.wnaf):
            f = Fk.sqr(f)
            c = coeffs[idx]
            idx += 1
            gtx = tuple(v * px % self.p for v in Fe.mul(c[1], self.twist))
            f = Fk.mul(f, self.embed(py_twist_sq, Fe.sub(Fe.sub(c[2], gtx), c[0])))
            if n != 0:
                c = coeffs[idx]
                idx += 1
                gtx = tuple(v * px % self.p for v in Fe.mul(c[1], self.twist))
                t = Fe.sub(c[2], gtx)
                c1 = Fe.sub(t, Q[1]) if n > 0 else Fe.add(t, Q[1])
                f = Fk.mul(f, self.embed(py_twist_sq, c1))
        if self.loop_count_neg:
            f = Fk.conj_half(f)
        return f

    # ---- mnt4/mod.rs:224-275, mnt6/mod.rs:224-270 ----------------------------------------------------------
    def final_exponentiation(self, f):
        Fk = self.Fk
        finv = Fk.inv(f)

        def first(elt, elt_inv):
            if self.ke == 2:
                return Fk.mul(Fk.frobenius(elt, 2), elt_inv)                    # elt^(q^2 - 1)
            e = Fk.mul(Fk.frobenius(elt, 3), elt_inv)                           # elt^(q^3 - 1)
            return Fk.mul(Fk.frobenius(e, 1), e)                                # ... ^(q + 1)
        a, ainv = first(f, finv), first(finv, f)
        w1 = Fk.pow(Fk.frobenius(a, 1), self.m1)
        w0 = Fk.pow(ainv, -self.m0) if self.m0 < 0 else Fk.pow(a, self.m0)
        return Fk.mul(w1, w0)

    def pairing(self, P, Q):
        if P is None or Q is None:
            return self.Fk.one
        return self.final_exponentiation(self.miller_loop(P, Q))

    def multi_pairing(self, pairs):
        """final_exponentiation(prod miller_loop) - PairingEngine::miller_loop over several pairs"""
        f = self.Fk.one
        for P, Q in pairs:
            if P is not None and Q is not None:
                f = self.Fk.mul(f, self.miller_loop(P, Q))
        return self.final_exponentiation(f)


_ENGINES = {}


def engine(name):
    if name not in _ENGINES:
        if name == "mnt4":
            _ENGINES[name] = Engine("mnt4", O.MNT4_G1, O.MNT4_G2, O.MNT4_FR)
        else:
            _ENGINES[name] = Engine("mnt6", O.MNT6_G1, O.MNT6_G2, O.MNT6_FR)
    return _ENGINES[name]


# ---------------------------------------------------------------------------------------------------------
# Groth16 key generation and verification (generator.rs:150-330, verifier.rs:18-44)
# ---------------------------------------------------------------------------------------------------------
def benchmark_circuit_matrices(num_constraints):
    """the R1CS of examples/snark-scalability/constraints.rs:19-91 as sparse rows over the variables
    [one, a, b, aux_0, ...] (num_inputs = 3): at[j], bt[j], ct[j] = {variable index: coefficient}"""
    ni = 3
    at, bt, ct = [], [], []
    a_var, b_var = 1, 2
    pushed = {1: 2}                       # `assignments.push((a_val, a_var))` twice (constraints.rs:27-31)
    nxt = ni
    for i in range(num_constraints - 1):
        c_var = nxt
        nxt += 1
        if i % 2:
            at.append({a_var: 1})
            bt.append({b_var: 1})
        else:
            row = {}
            for v in (a_var, b_var):
                row[v] = row.get(v, 0) + 1
            at.append(row)
            bt.append({0: 1})
        ct.append({c_var: 1})
        pushed[c_var] = pushed.get(c_var, 0) + 1
        a_var, b_var = b_var, c_var
    at.append(dict(pushed))
    bt.append(dict(pushed))
    ct.append({nxt: 1})
    return at, bt, ct, ni, nxt + 1 - ni


def lagrange_coefficients(F, n, tau):
    """EvaluationDomain::evaluate_all_lagrange_coefficients (domain.rs:183-220) for tau outside the domain"""
    p = F.p
    dom = O.EvaluationDomain(F, n)
    zt = (pow(tau, n, p) - 1) % p
    assert zt != 0
    l = zt * dom.size_inv % p
    out, g = [], 1
    for _ in range(n):
        out.append(l * g % p * pow((tau - g) % p, -1, p) % p)
        g = g * dom.group_gen % p
    return out, zt


def generate_parameters(eng, at, bt, ct, num_inputs, num_aux, alpha, beta, gamma, delta, tau, g1_gen, g2_gen):
    """generate_parameters (generator.rs:150-330) with the toxic waste given.  Returns the proving key as an
    oracle Groth16Key plus the verifying key (alpha_g1_beta_g2, gamma_g2, delta_g2, gamma_abc_g1)."""
    F = eng.fr
    p = F.p
    C1, C2 = eng.g1, eng.g2
    nc = len(at)
    n = 1
    while n < nc + (num_inputs - 1) + 1:
        n <<= 1
    u, zt = lagrange_coefficients(F, n, tau)
    nv = num_inputs + num_aux
    a, b, c = [0] * nv, [0] * nv, [0] * nv
    for i in range(num_inputs):                       # r1cs_to_qap.rs:38-40
        a[i] = u[nc + i]
    for j in range(nc):
        for row, dst in ((at[j], a), (bt[j], b), (ct[j], c)):
            for v, coeff in row.items():
                dst[v] = (dst[v] + u[j] * coeff) % p
    ginv, dinv = pow(gamma, -1, p), pow(delta, -1, p)
    gamma_abc = [(beta * a[i] + alpha * b[i] + c[i]) * ginv % p for i in range(num_inputs)]
    l = [(beta * a[i] + alpha * b[i] + c[i]) * dinv % p for i in range(nv)]
    mul1 = lambda s: C1.mul(g1_gen, s % C1.r)
    mul2 = lambda s: C2.mul(g2_gen, s % C2.r)
    key = O.Groth16Key(C1, C2, mul1(alpha), mul1(beta), mul2(beta), mul1(delta), mul2(delta),
                       [mul1(x) for x in a], [mul1(x) for x in b], [mul2(x) for x in b],
                       [mul1(zt * dinv % p * pow(tau, i, p)) for i in range(n - 1)],
                       [mul1(x) for x in l[num_inputs:]])
    vk = dict(alpha_g1_beta_g2=eng.pairing(key.alpha_g1, key.beta_g2), gamma_g2=mul2(gamma), delta_g2=mul2(delta),
              gamma_abc_g1=[mul1(x) for x in gamma_abc])
    return key, vk, n


def verify_proof(eng, vk, proof, public_inputs):
    """verifier.rs:18-44: e(A, B) * e(g_ic, -gamma) * e(C, -delta) == e(alpha, beta); proof = (A, B, C) affine
    oracle points, public_inputs without the leading one"""
    C1, C2 = eng.g1, eng.g2
    assert len(public_inputs) + 1 == len(vk["gamma_abc_g1"])
    g_ic = vk["gamma_abc_g1"][0]
    for x, base in zip(public_inputs, vk["gamma_abc_g1"][1:]):
        g_ic = C1.add(g_ic, C1.mul(base, x % C1.r))
    A, B, C = proof
    test = eng.multi_pairing([(A, B), (g_ic, C2.neg(vk["gamma_g2"])), (C, C2.neg(vk["delta_g2"]))])
    return test == vk["alpha_g1_beta_g2"]
