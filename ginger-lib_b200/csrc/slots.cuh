// Field / tower arithmetic on operands that live in shared-memory "slots".
//
// Why: a 753-bit element is 24 registers; a mixed addition keeps ~8 of them live and the
// multiplier itself needs ~100 registers, so the curve formulas cannot stay in the register
// file.  Letting the compiler spill puts ~1.7 KB/thread of stack through L1 (measured: the
// first version of the bucket kernels ran at ~20% of the integer roofline).  Here every curve
// operand has a fixed home in shared memory instead:
//
//   slot s, 16-byte chunk k (0..5), thread t   ->   g_slots[(s*6 + k) * T + t]
//
// so a warp's access to one chunk of one slot is 512 contiguous bytes (conflict-free
// LDS.128 / STS.128), and the formulas become short sequences of out-of-line
// "slot <- slot op slot" calls whose only arguments are small integers.  A product costs
// 12 LDS.128 + 6 STS.128 next to ~1200 IMAD.WIDE: the memory side is noise, the register
// file holds only the multiplier's working set, and the footprint (8 slots = 768 B/thread
// for a G1 mixed addition) is fixed by construction.
//
// Replaces the same reference code as fq.cuh / fqk.cuh (fields/models/fp_768.rs, fp2.rs,
// fp3.rs); the arithmetic itself is fq.cuh's.
#pragma once
#include "device.cuh"
#include "fq.cuh"

namespace g753 {

#if defined(G753_HOST_EMUL)
// test-only host build: one static buffer stands in for the block's shared memory
static uint4 g_slots[(240 * 1024) / sizeof(uint4)];
#else
extern __shared__ uint4 g_slots[];
#endif

constexpr int SLOT_CHUNKS = NL / 4;  // 6 x 16 bytes per Fq

template <int T>
G753_D uint4* slot_ptr(int s) {
  return g_slots + s * (SLOT_CHUNKS * T) + threadIdx.x;
}
template <int T>
G753_D Fq s_ld(int s) {
  const uint4* p = slot_ptr<T>(s);
  Fq r;
#pragma unroll
  for (int k = 0; k < SLOT_CHUNKS; k++) {
    uint4 v = p[k * T];
    r.l[4 * k] = v.x;
    r.l[4 * k + 1] = v.y;
    r.l[4 * k + 2] = v.z;
    r.l[4 * k + 3] = v.w;
  }
  return r;
}
template <int T>
G753_D void s_st(int s, const Fq& a) {
  uint4* p = slot_ptr<T>(s);
#pragma unroll
  for (int k = 0; k < SLOT_CHUNKS; k++) {
    uint4 v;
    v.x = a.l[4 * k];
    v.y = a.l[4 * k + 1];
    v.z = a.l[4 * k + 2];
    v.w = a.l[4 * k + 3];
    p[k * T] = v;
  }
}
// global <-> register, 16-byte vector accesses (Fq is 16-byte aligned)
G753_D Fq g_ld(const Fq* g) {
  const uint4* p = (const uint4*)g;
  Fq r;
#pragma unroll
  for (int k = 0; k < SLOT_CHUNKS; k++) {
    uint4 v = p[k];
    r.l[4 * k] = v.x;
    r.l[4 * k + 1] = v.y;
    r.l[4 * k + 2] = v.z;
    r.l[4 * k + 3] = v.w;
  }
  return r;
}
G753_D void g_st(Fq* g, const Fq& a) {
  uint4* p = (uint4*)g;
#pragma unroll
  for (int k = 0; k < SLOT_CHUNKS; k++) {
    uint4 v;
    v.x = a.l[4 * k];
    v.y = a.l[4 * k + 1];
    v.z = a.l[4 * k + 2];
    v.w = a.l[4 * k + 3];
    p[k] = v;
  }
}

// ---- Fq on slots (d may alias a or b everywhere: operands are read before d is written) ----
template <int FID, int T>
G753_NI void s_mul(int d, int a, int b) {
  s_st<T>(d, fq_mul<FID>(s_ld<T>(a), s_ld<T>(b)));
}
template <int FID, int T>
G753_NI void s_sqr(int d, int a) {
  s_st<T>(d, fq_sqr<FID>(s_ld<T>(a)));
}
template <int FID, int T>
G753_NI void s_add(int d, int a, int b) {
  s_st<T>(d, fq_add<FID>(s_ld<T>(a), s_ld<T>(b)));
}
template <int FID, int T>
G753_NI void s_sub(int d, int a, int b) {
  s_st<T>(d, fq_sub<FID>(s_ld<T>(a), s_ld<T>(b)));
}
template <int FID, int T>
G753_NI void s_dbl(int d, int a) {
  s_st<T>(d, fq_dbl<FID>(s_ld<T>(a)));
}
template <int FID, int T>
G753_NI void s_neg(int d, int a) {
  s_st<T>(d, fq_neg<FID>(s_ld<T>(a)));
}
template <int FID, int T, unsigned KK>
G753_NI void s_mul_small(int d, int a) {
  s_st<T>(d, fq_mul_small<FID, KK>(s_ld<T>(a)));
}
template <int FID, int T>
G753_NI void s_inv(int d, int a) {  // Fermat; the inverse is unique, so the limbs equal fp_768.rs:551-605's
  s_st<T>(d, fq_inv<FID>(s_ld<T>(a)));
}
template <int T>
G753_D void s_copy(int d, int a) {
  if (d == a) return;
  const uint4* p = slot_ptr<T>(a);
  uint4* q = slot_ptr<T>(d);
#pragma unroll
  for (int k = 0; k < SLOT_CHUNKS; k++) q[k * T] = p[k * T];
}
template <int T>
G753_D bool s_is_zero(int a) {
  const uint4* p = slot_ptr<T>(a);
  uint32_t t = 0;
#pragma unroll
  for (int k = 0; k < SLOT_CHUNKS; k++) {
    uint4 v = p[k * T];
    t |= v.x | v.y | v.z | v.w;
  }
  return t == 0;
}
template <int T>
G753_D void s_set_zero(int d) {
  uint4* q = slot_ptr<T>(d);
  uint4 z;
  z.x = z.y = z.z = z.w = 0;
#pragma unroll
  for (int k = 0; k < SLOT_CHUNKS; k++) q[k * T] = z;
}
template <int FID, int T>
G753_D void s_set_one(int d) {
  s_st<T>(d, fq_one<FID>());
}
template <int T>
G753_D void s_ldg(int d, const Fq* g) {
  const uint4* p = (const uint4*)g;
  uint4* q = slot_ptr<T>(d);
#pragma unroll
  for (int k = 0; k < SLOT_CHUNKS; k++) q[k * T] = p[k];
}
template <int T>
G753_D void s_stg(Fq* g, int a) {
  uint4* p = (uint4*)g;
  const uint4* q = slot_ptr<T>(a);
#pragma unroll
  for (int k = 0; k < SLOT_CHUNKS; k++) p[k] = q[k * T];
}

// ---- towers: an element is K consecutive slots; `t` is the first of NTMP scratch slots ----
template <int FID, int T_>
struct Tw1 {
  static constexpr int K = 1, NTMP = 0, FIELD = FID, T = T_;
  static G753_D void mul(int d, int a, int b, int) { s_mul<FID, T>(d, a, b); }
  static G753_D void sqr(int d, int a, int) { s_sqr<FID, T>(d, a); }
  static G753_D void inv(int d, int a, int) { s_inv<FID, T>(d, a); }
  static G753_D void add(int d, int a, int b) { s_add<FID, T>(d, a, b); }
  static G753_D void sub(int d, int a, int b) { s_sub<FID, T>(d, a, b); }
  static G753_D void dbl(int d, int a) { s_dbl<FID, T>(d, a); }
  static G753_D void neg(int d, int a) { s_neg<FID, T>(d, a); }
  static G753_D void copy(int d, int a) { s_copy<T>(d, a); }
  static G753_D bool is_zero(int a) { return s_is_zero<T>(a); }
  static G753_D void set_zero(int d) { s_set_zero<T>(d); }
  static G753_D void set_one(int d) { s_set_one<FID, T>(d); }
  static G753_D void ldg(int d, const Fq* g) { s_ldg<T>(d, g); }
  static G753_D void stg(Fq* g, int a) { s_stg<T>(g, a); }
};

template <int FID, int T_, unsigned NR>
struct Tw2 {
  static constexpr int K = 2, NTMP = 3, FIELD = FID, T = T_;
  // Karatsuba (fp2.rs:387-401), 3 products
  static G753_NI void mul(int d, int a, int b, int t) {
    s_mul<FID, T>(t, a, b);
    s_mul<FID, T>(t + 1, a + 1, b + 1);
    s_add<FID, T>(t + 2, a, a + 1);
    s_add<FID, T>(d + 1, b, b + 1);  // a1 / b1 are consumed by now, so d may alias a or b
    s_mul<FID, T>(d + 1, t + 2, d + 1);
    s_sub<FID, T>(d + 1, d + 1, t);
    s_sub<FID, T>(d + 1, d + 1, t + 1);
    s_mul_small<FID, T, NR>(t + 1, t + 1);
    s_add<FID, T>(d, t, t + 1);
  }
  // complex squaring (fp2.rs:128-144), 2 products
  static G753_NI void sqr(int d, int a, int t) {
    s_mul<FID, T>(t, a, a + 1);
    s_add<FID, T>(t + 1, a, a + 1);
    s_mul_small<FID, T, NR>(t + 2, a + 1);
    s_add<FID, T>(t + 2, a, t + 2);
    s_mul<FID, T>(t + 1, t + 1, t + 2);
    s_mul_small<FID, T, NR>(t + 2, t);
    s_dbl<FID, T>(d + 1, t);
    s_sub<FID, T>(d, t + 1, t);
    s_sub<FID, T>(d, d, t + 2);
  }
  // fp2.rs:83-127: (a0 - a1 u) / (a0^2 - NR a1^2)
  static G753_NI void inv(int d, int a, int t) {
    s_sqr<FID, T>(t, a);
    s_sqr<FID, T>(t + 1, a + 1);
    s_mul_small<FID, T, NR>(t + 1, t + 1);
    s_sub<FID, T>(t, t, t + 1);
    s_inv<FID, T>(t, t);
    s_mul<FID, T>(d, a, t);
    s_mul<FID, T>(d + 1, a + 1, t);
    s_neg<FID, T>(d + 1, d + 1);
  }
  static G753_D void add(int d, int a, int b) {
    s_add<FID, T>(d, a, b);
    s_add<FID, T>(d + 1, a + 1, b + 1);
  }
  static G753_D void sub(int d, int a, int b) {
    s_sub<FID, T>(d, a, b);
    s_sub<FID, T>(d + 1, a + 1, b + 1);
  }
  static G753_D void dbl(int d, int a) {
    s_dbl<FID, T>(d, a);
    s_dbl<FID, T>(d + 1, a + 1);
  }
  static G753_D void neg(int d, int a) {
    s_neg<FID, T>(d, a);
    s_neg<FID, T>(d + 1, a + 1);
  }
  static G753_D void copy(int d, int a) {
    s_copy<T>(d, a);
    s_copy<T>(d + 1, a + 1);
  }
  static G753_D bool is_zero(int a) { return s_is_zero<T>(a) && s_is_zero<T>(a + 1); }
  static G753_D void set_zero(int d) {
    s_set_zero<T>(d);
    s_set_zero<T>(d + 1);
  }
  static G753_D void set_one(int d) {
    s_set_one<FID, T>(d);
    s_set_zero<T>(d + 1);
  }
  static G753_D void ldg(int d, const Fq* g) {
    s_ldg<T>(d, g);
    s_ldg<T>(d + 1, g + 1);
  }
  static G753_D void stg(Fq* g, int a) {
    s_stg<T>(g, a);
    s_stg<T>(g + 1, a + 1);
  }
};

template <int FID, int T_, unsigned NR>
struct Tw3 {
  static constexpr int K = 3, NTMP = 6, FIELD = FID, T = T_;
  // Karatsuba (fp3.rs:451-478), 6 products
  static G753_NI void mul(int d, int a, int b, int t) {
    s_mul<FID, T>(t, a, b);                  // v0
    s_mul<FID, T>(t + 1, a + 1, b + 1);      // v1
    s_mul<FID, T>(t + 2, a + 2, b + 2);      // v2
    s_add<FID, T>(t + 3, a + 1, a + 2);
    s_add<FID, T>(t + 4, b + 1, b + 2);
    s_mul<FID, T>(t + 3, t + 3, t + 4);
    s_sub<FID, T>(t + 3, t + 3, t + 1);
    s_sub<FID, T>(t + 3, t + 3, t + 2);
    s_mul_small<FID, T, NR>(t + 3, t + 3);   // c0 - v0
    s_add<FID, T>(t + 4, a, a + 1);
    s_add<FID, T>(t + 5, b, b + 1);
    s_mul<FID, T>(t + 4, t + 4, t + 5);
    s_sub<FID, T>(t + 4, t + 4, t);
    s_sub<FID, T>(t + 4, t + 4, t + 1);
    s_mul_small<FID, T, NR>(t + 5, t + 2);
    s_add<FID, T>(t + 4, t + 4, t + 5);      // c1
    s_add<FID, T>(t + 5, a, a + 2);
    s_add<FID, T>(d + 1, b, b + 2);          // a1 / b1 consumed: d may alias a or b
    s_mul<FID, T>(t + 5, t + 5, d + 1);
    s_sub<FID, T>(t + 5, t + 5, t);
    s_sub<FID, T>(t + 5, t + 5, t + 2);
    s_add<FID, T>(d + 2, t + 5, t + 1);      // c2
    s_add<FID, T>(d, t, t + 3);
    s_copy<T>(d + 1, t + 4);
  }
  // Chung-Hasan SQR2 (fp3.rs:165-185)
  static G753_NI void sqr(int d, int a, int t) {
    s_sqr<FID, T>(t, a);                     // s0
    s_mul<FID, T>(t + 1, a, a + 1);
    s_dbl<FID, T>(t + 1, t + 1);             // s1
    s_sub<FID, T>(t + 2, a, a + 1);
    s_add<FID, T>(t + 2, t + 2, a + 2);
    s_sqr<FID, T>(t + 2, t + 2);             // s2
    s_mul<FID, T>(t + 3, a + 1, a + 2);
    s_dbl<FID, T>(t + 3, t + 3);             // s3
    s_sqr<FID, T>(t + 4, a + 2);             // s4
    s_add<FID, T>(t + 2, t + 2, t + 1);
    s_add<FID, T>(t + 2, t + 2, t + 3);
    s_sub<FID, T>(t + 2, t + 2, t);
    s_sub<FID, T>(d + 2, t + 2, t + 4);      // c2
    s_mul_small<FID, T, NR>(t + 5, t + 3);
    s_add<FID, T>(d, t, t + 5);              // c0
    s_mul_small<FID, T, NR>(t + 5, t + 4);
    s_add<FID, T>(d + 1, t + 1, t + 5);      // c1
  }
  // fp3.rs:107-164
  static G753_NI void inv(int d, int a, int t) {
    s_sqr<FID, T>(t, a);
    s_mul<FID, T>(t + 3, a + 1, a + 2);
    s_mul_small<FID, T, NR>(t + 3, t + 3);
    s_sub<FID, T>(t, t, t + 3);              // t0 = a0^2 - NR a1 a2
    s_sqr<FID, T>(t + 1, a + 2);
    s_mul_small<FID, T, NR>(t + 1, t + 1);
    s_mul<FID, T>(t + 3, a, a + 1);
    s_sub<FID, T>(t + 1, t + 1, t + 3);      // t1 = NR a2^2 - a0 a1
    s_sqr<FID, T>(t + 2, a + 1);
    s_mul<FID, T>(t + 3, a, a + 2);
    s_sub<FID, T>(t + 2, t + 2, t + 3);      // t2 = a1^2 - a0 a2
    s_mul<FID, T>(t + 3, a + 2, t + 1);
    s_mul<FID, T>(t + 4, a + 1, t + 2);
    s_add<FID, T>(t + 3, t + 3, t + 4);
    s_mul_small<FID, T, NR>(t + 3, t + 3);
    s_mul<FID, T>(t + 4, a, t);
    s_add<FID, T>(t + 3, t + 3, t + 4);      // norm
    s_inv<FID, T>(t + 3, t + 3);
    s_mul<FID, T>(d, t, t + 3);
    s_mul<FID, T>(d + 1, t + 1, t + 3);
    s_mul<FID, T>(d + 2, t + 2, t + 3);
  }
  static G753_D void add(int d, int a, int b) {
    for (int i = 0; i < 3; i++) s_add<FID, T>(d + i, a + i, b + i);
  }
  static G753_D void sub(int d, int a, int b) {
    for (int i = 0; i < 3; i++) s_sub<FID, T>(d + i, a + i, b + i);
  }
  static G753_D void dbl(int d, int a) {
    for (int i = 0; i < 3; i++) s_dbl<FID, T>(d + i, a + i);
  }
  static G753_D void neg(int d, int a) {
    for (int i = 0; i < 3; i++) s_neg<FID, T>(d + i, a + i);
  }
  static G753_D void copy(int d, int a) {
    for (int i = 0; i < 3; i++) s_copy<T>(d + i, a + i);
  }
  static G753_D bool is_zero(int a) { return s_is_zero<T>(a) && s_is_zero<T>(a + 1) && s_is_zero<T>(a + 2); }
  static G753_D void set_zero(int d) {
    for (int i = 0; i < 3; i++) s_set_zero<T>(d + i);
  }
  static G753_D void set_one(int d) {
    s_set_one<FID, T>(d);
    s_set_zero<T>(d + 1);
    s_set_zero<T>(d + 2);
  }
  static G753_D void ldg(int d, const Fq* g) {
    for (int i = 0; i < 3; i++) s_ldg<T>(d + i, g + i);
  }
  static G753_D void stg(Fq* g, int a) {
    for (int i = 0; i < 3; i++) s_stg<T>(g + i, a + i);
  }
};

}  // namespace g753
