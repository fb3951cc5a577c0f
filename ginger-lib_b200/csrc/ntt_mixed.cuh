// Mixed-radix transform for domain sizes N = 2^a * m (m odd, N | p - 1) - BASELINE config 4.
//
// The reference snapshot has no mixed-radix domain, so this follows the
// definition: omega = GENERATOR^((p-1)/N), out[k] = sum_i in[i] omega^(i k), natural order, coset
// generator 17 as in EvaluationDomain (domain.rs:140-179).  PARITY UNPINNED: the only oracle is the
// direct DFT (oracle/g753.py dft_naive).
//
// Four-step split n1 = m, n2 = 2^a:  i = i1 n2 + i2,  k = k1 + m k2
//   X[k1 + m k2] = sum_{i2} (omega^m)^(i2 k2) * omega^(i2 k1) * sum_{i1} x[i1 n2 + i2] zeta^(i1 k1),  zeta = omega^n2
//   step 1  k_small_dft:   length-m DFTs down the columns, straight from the definition (m is small)
//   step 2  twiddle omega^(i2 k1): the `pre` table of step 3
//   step 3  m batched radix-2 transforms of length 2^a (k_ntt_pass; omega^m is the radix-2 root)
//   step 4  k_permute3:    Z[k1][k2] -> out[k1 + m k2]
#pragma once
#include "ntt.cuh"
#include "ntt_dist.cuh"

namespace g753 {

constexpr unsigned MIXED_MAX_M = 1024;

struct MixedTables {
  size_t N = 0;
  unsigned a = 0, m = 1;
  Fq* zeta = nullptr;       // [2][m]: zeta^j, zeta^-j
  Fq* tw = nullptr;         // [2][m][n2]: omega^(i2 k1), omega^-(i2 k1)
  Fq* coset = nullptr;      // [N] g^i
  Fq* coset_inv = nullptr;  // [N] g^-i / N
  Fq* consts = nullptr;     // [0] omega, [1] omega^-1, [2] N^-1, [3] zeta, [4] zeta^-1, [5] g, [6] g^-1
  void release() {
    dev_free(zeta);
    dev_free(tw);
    dev_free(coset);
    dev_free(coset_inv);
    dev_free(consts);
    zeta = tw = coset = coset_inv = consts = nullptr;
  }
};

// exp = (p - 1) / N as 24 limbs (host-computed); consts as documented in MixedTables
template <int FID>
__global__ void k_mixed_setup(const uint32_t* __restrict__ exp, uint64_t N, uint64_t n2, Fq* consts) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Fq g, gi;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    g.l[i] = G753_FC(FID).gen[i];
    gi.l[i] = G753_FC(FID).gen_inv[i];
  }
  uint32_t e[NL];
  for (int i = 0; i < NL; i++) e[i] = exp[i];
  Fq w = fq_pow<FID>(g, e);
  Fq wi = fq_inv<FID>(w);
  // N in Montgomery form: N * R = (N as raw integer) * R^2 / R
  Fq nraw = fq_zero<FID>();
  nraw.l[0] = (uint32_t)N;
  nraw.l[1] = (uint32_t)(N >> 32);
  Fq nm = fq_to_mont<FID>(nraw);
  consts[0] = w;
  consts[1] = wi;
  consts[2] = fq_inv<FID>(nm);
  consts[3] = fq_pow_u64<FID>(w, n2);
  consts[4] = fq_pow_u64<FID>(wi, n2);
  consts[5] = g;
  consts[6] = gi;
}

// column DFTs of an m x n2 row-major matrix: out[k1][i2] = sum_{i1} in[i1][i2] * zt[(i1 k1) mod m]
template <int FID>
__global__ void __launch_bounds__(128)
k_small_dft(const Fq* __restrict__ in, Fq* __restrict__ out, const Fq* __restrict__ zt, unsigned m, size_t n2,
            const Fq* __restrict__ pre) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)m * n2) return;
  const unsigned k1 = (unsigned)(t / n2);
  const size_t i2 = t % n2;
  Fq acc = fq_zero<FID>();
  unsigned e = 0;  // (i1 k1) mod m
  for (unsigned i1 = 0; i1 < m; i1++) {
    Fq v = in[(size_t)i1 * n2 + i2];
    if (pre != nullptr) v = fq_mul<FID>(v, pre[(size_t)i1 * n2 + i2]);
    if (e != 0) v = fq_mul<FID>(v, zt[e]);
    acc = fq_add<FID>(acc, v);
    e += k1;
    if (e >= m) e -= m;
  }
  out[t] = acc;
}

// The same column DFT of length m = q1 q2 in two stages (Cooley-Tukey on the odd part: i1 = a q2 + b,
// k1 = c + q1 d):
//   stage A:  T[b][c] = zeta^(b c) * sum_a in[a q2 + b] (zeta^q2)^(a c)        q1 + 1 products per element
//   stage C:  out[c + q1 d] = sum_b T[b][c] (zeta^q1)^(b d)                     q2 products per element
// instead of m: 25 points = 5 x 5 cost ~10 products per element instead of 25 (trivial factors are skipped).
template <int FID>
__global__ void __launch_bounds__(128)
k_small_dft_a(const Fq* __restrict__ in, Fq* __restrict__ tmp, const Fq* __restrict__ zt, unsigned m, unsigned q1,
              unsigned q2, size_t n2, const Fq* __restrict__ pre) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)m * n2) return;
  const size_t i2 = t % n2;
  const unsigned bc = (unsigned)(t / n2), b = bc / q1, c = bc % q1;   // tmp layout [b][c][i2]
  Fq acc = fq_zero<FID>();
  const unsigned step = (q2 * c) % m;
  unsigned e = 0;  // (a q2 c) mod m
  for (unsigned a = 0; a < q1; a++) {
    const size_t src = (size_t)(a * q2 + b) * n2 + i2;
    Fq v = in[src];
    if (pre != nullptr) v = fq_mul<FID>(v, pre[src]);
    if (e != 0) v = fq_mul<FID>(v, zt[e]);
    acc = fq_add<FID>(acc, v);
    e += step;
    if (e >= m) e -= m;
  }
  const unsigned tw = (b * c) % m;
  if (tw != 0) acc = fq_mul<FID>(acc, zt[tw]);
  tmp[t] = acc;
}
template <int FID>
__global__ void __launch_bounds__(128)
k_small_dft_c(const Fq* __restrict__ tmp, Fq* __restrict__ out, const Fq* __restrict__ zt, unsigned m, unsigned q1,
              unsigned q2, size_t n2) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)m * n2) return;
  const size_t i2 = t % n2;
  const unsigned k1 = (unsigned)(t / n2), c = k1 % q1, d = k1 / q1;   // out row k1 = c + q1 d
  Fq acc = fq_zero<FID>();
  const unsigned step = (q1 * d) % m;
  unsigned e = 0;  // (b q1 d) mod m
  for (unsigned b = 0; b < q2; b++) {
    Fq v = tmp[(size_t)(b * q1 + c) * n2 + i2];
    if (e != 0) v = fq_mul<FID>(v, zt[e]);
    acc = fq_add<FID>(acc, v);
    e += step;
    if (e >= m) e -= m;
  }
  out[t] = acc;
}
// m = q1 q2 with q1 the largest divisor of m not above sqrt(m) (1 for a prime: single stage)
static inline unsigned mixed_q1(unsigned m) {
  unsigned best = 1;
  for (unsigned q = 2; q * q <= m; q++)
    if (m % q == 0) best = q;
  return best;
}

static inline bool div_small_768(const uint32_t* num, uint64_t d, uint32_t* quo) {
  // quo = num / d for a 768-bit num and d < 2^32; returns true when the division is exact
  uint64_t rem = 0;
  for (int i = NL - 1; i >= 0; i--) {
    uint64_t cur = (rem << 32) | num[i];
    quo[i] = (uint32_t)(cur / d);
    rem = cur % d;
  }
  return rem == 0;
}

// N = 2^a * m with m odd <= MIXED_MAX_M, N | p - 1, a <= two-adicity.  Returns false otherwise.
static inline bool mixed_split(int field, uint64_t N, unsigned* a_out, unsigned* m_out, uint32_t* exp) {
  if (N == 0 || N >= (1ull << 31)) return false;
  unsigned a = 0;
  uint64_t m = N;
  while ((m & 1) == 0) {
    m >>= 1;
    a++;
  }
  if (m > MIXED_MAX_M || a > G753_FIELD_CONSTANTS[field].two_adicity || a > NTT_MAX_LOG) return false;
  uint32_t pm1[NL];
  for (int i = 0; i < NL; i++) pm1[i] = G753_FIELD_CONSTANTS[field].p[i];
  pm1[0] -= 1;  // p is odd
  uint32_t q[NL];
  if (!div_small_768(pm1, N, q)) return false;
  if (a_out) *a_out = a;
  if (m_out) *m_out = (unsigned)m;
  if (exp)
    for (int i = 0; i < NL; i++) exp[i] = q[i];
  return true;
}

}  // namespace g753
