// 753-bit prime-field arithmetic for sm_100a: 24 x 32-bit limbs, Montgomery radix R = 2^768.
//
// Replaces (reference, relative to /root/reference/algebra/src):
//   fields/models/fp_768.rs:1009-1185  Fp768::mul_assign      -> fq_mul
//   fields/models/fp_768.rs:339-548    Fp768::square_in_place -> fq_sqr
//   fields/models/fp_768.rs:929-949    add_assign/sub_assign  -> fq_add / fq_sub
//   fields/models/fp_768.rs:870-883    neg                    -> fq_neg
//   fields/models/fp_768.rs:303-309    double_in_place        -> fq_dbl
//   fields/models/fp_768.rs:551-605    inverse                -> fq_inv (Fermat; same value)
// Limbs are the reference's own in-memory form (BigInteger768 = 12 x u64 little endian,
// biginteger/mod.rs:20) re-read as 24 x u32, so raw buffers cross the C ABI unconverted.
// All results are canonical (< p), as the reference's are (fp_768.rs:39-48).
//
// The multiplier is an interleaved (CIOS-style) Montgomery product laid out so that every
// 32x32->64 partial product is one mad.lo.cc/madc.hi.cc pair on a single carry chain: the
// products of the even-indexed limbs of `a` land on columns (j, j+1) and form one chain,
// those of the odd-indexed limbs form a second chain one column up.  The two accumulators
// swap roles after each one-limb Montgomery shift, so no partial product ever has to be
// re-aligned.  ptxas turns each lo/hi pair into IMAD.WIDE.U32 with carry-in/out (checked
// with cuobjdump, see DESIGN.md "Field multiplier").
//
// The same source compiles for the host (tests/host_emul) with the carry flag emulated in
// a thread-local variable: that build exists only to test the limb logic without a GPU and
// is never linked into the product library.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define G753_HD __host__ __device__ __forceinline__
#define G753_D __device__ __forceinline__
#define G753_NI __host__ __device__ __noinline__
#else
#define G753_HD inline
#define G753_D inline
#define G753_NI
#ifndef __align__
#define __align__(n) alignas(n)
#endif
#endif

namespace g753 {

constexpr int NL = 24;  // 32-bit limbs per base-field element (768 / 32)

struct FieldConstants {
  uint32_t p[NL];        // modulus
  uint32_t one[NL];      // R mod p      (Montgomery form of 1)
  uint32_t r2[NL];       // R^2 mod p    (to-Montgomery multiplier)
  uint32_t r3[NL];       // R^3 mod p    (Montgomery inverse fix-up)
  uint32_t gen[NL];      // 17 * R       (multiplicative generator, FpParameters::GENERATOR)
  uint32_t gen_inv[NL];  // 17^-1 * R
  uint32_t root[NL];     // 2^s-th root of unity * R (FpParameters::ROOT_OF_UNITY)
  uint32_t p_minus_2[NL];
  uint32_t curve_b[NL];  // G1 coefficient b * R of the curve whose base field this is
  uint32_t inv32;        // -p^-1 mod 2^32
  uint32_t two_adicity;
};

#include "constants.inc"
static const FieldConstants G753_FIELD_CONSTANTS[2] = G753_FIELD_CONSTANTS_INIT;

#if defined(__CUDACC__)
// statically initialised, one copy per translation unit (no cudaMemcpyToSymbol at start-up)
static __constant__ FieldConstants d_fc[2] = G753_FIELD_CONSTANTS_INIT;
#endif

#if defined(__CUDA_ARCH__)
#define G753_FC(FID) (d_fc[FID])
#else
#define G753_FC(FID) (G753_FIELD_CONSTANTS[FID])
static thread_local uint32_t g_host_cf = 0;  // emulated CC.CF (host test build only)
#endif

// ------------------------------------------------------------------------------------
// carry-chain primitives.  Each maps to exactly one PTX instruction on the device.
// ------------------------------------------------------------------------------------
G753_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
G753_HD uint32_t mul_hi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

#if defined(__CUDA_ARCH__)
#define G753_ASM3(name, ins)                                                         \
  G753_HD uint32_t name(uint32_t a, uint32_t b) {                                    \
    uint32_t r;                                                                      \
    asm volatile(ins " %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));                    \
    return r;                                                                        \
  }
#define G753_ASM4(name, ins)                                                         \
  G753_HD uint32_t name(uint32_t a, uint32_t b, uint32_t c) {                        \
    uint32_t r;                                                                      \
    asm volatile(ins " %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));        \
    return r;                                                                        \
  }
G753_ASM3(add_cc, "add.cc.u32")
G753_ASM3(addc_cc, "addc.cc.u32")
G753_ASM3(addc, "addc.u32")
G753_ASM3(sub_cc, "sub.cc.u32")
G753_ASM3(subc_cc, "subc.cc.u32")
G753_ASM3(subc, "subc.u32")
G753_ASM4(mad_lo_cc, "mad.lo.cc.u32")
G753_ASM4(madc_lo_cc, "madc.lo.cc.u32")
G753_ASM4(mad_hi_cc, "mad.hi.cc.u32")
G753_ASM4(madc_hi_cc, "madc.hi.cc.u32")
G753_ASM4(madc_hi, "madc.hi.u32")
G753_ASM4(madc_lo, "madc.lo.u32")
#undef G753_ASM3
#undef G753_ASM4
#else
G753_HD uint32_t add_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a + b;
  g_host_cf = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
G753_HD uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a + b + g_host_cf;
  g_host_cf = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
G753_HD uint32_t addc(uint32_t a, uint32_t b) { return a + b + g_host_cf; }
G753_HD uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a - b;
  g_host_cf = (uint32_t)((t >> 32) & 1);  // borrow (PTX keeps the borrow in CC.CF for sub)
  return (uint32_t)t;
}
G753_HD uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a - b - g_host_cf;
  g_host_cf = (uint32_t)((t >> 32) & 1);
  return (uint32_t)t;
}
G753_HD uint32_t subc(uint32_t a, uint32_t b) { return a - b - g_host_cf; }
G753_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint64_t t = (uint64_t)(uint32_t)(a * b) + c;
  g_host_cf = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
G753_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint64_t t = (uint64_t)(uint32_t)(a * b) + c + g_host_cf;
  g_host_cf = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
G753_HD uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint64_t t = (((uint64_t)a * b) >> 32) + c;
  g_host_cf = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
G753_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint64_t t = (((uint64_t)a * b) >> 32) + c + g_host_cf;
  g_host_cf = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
G753_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
  return (uint32_t)((((uint64_t)a * b) >> 32) + c + g_host_cf);
}
G753_HD uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { return a * b + c + g_host_cf; }
#endif

// ------------------------------------------------------------------------------------
// element type
// ------------------------------------------------------------------------------------
struct __align__(16) Fq {
  uint32_t l[NL];
};

template <int FID>
G753_HD Fq fq_zero() {
  Fq r;
#pragma unroll
  for (int i = 0; i < NL; i++) r.l[i] = 0;
  return r;
}
template <int FID>
G753_HD Fq fq_one() {
  Fq r;
#pragma unroll
  for (int i = 0; i < NL; i++) r.l[i] = G753_FC(FID).one[i];
  return r;
}
G753_HD bool fq_is_zero(const Fq& a) {
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) t |= a.l[i];
  return t == 0;
}
G753_HD bool fq_eq(const Fq& a, const Fq& b) {
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) t |= a.l[i] ^ b.l[i];
  return t == 0;
}

// r = (x >= p) ? x - p : x, for x < 2p
template <int FID>
G753_HD void fq_reduce_once(uint32_t* x) {
  uint32_t t[NL];
  t[0] = sub_cc(x[0], G753_FC(FID).p[0]);
#pragma unroll
  for (int i = 1; i < NL; i++) t[i] = subc_cc(x[i], G753_FC(FID).p[i]);
  uint32_t borrow = subc(0, 0);  // 0 - 0 - CF : 0xffffffff when x < p
#pragma unroll
  for (int i = 0; i < NL; i++) x[i] = borrow ? x[i] : t[i];
}

template <int FID>
G753_HD Fq fq_add(const Fq& a, const Fq& b) {
  Fq r;
  r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
  for (int i = 1; i < NL - 1; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
  r.l[NL - 1] = addc(a.l[NL - 1], b.l[NL - 1]);  // 2p < 2^768: no carry out
  fq_reduce_once<FID>(r.l);
  return r;
}

template <int FID>
G753_HD Fq fq_sub(const Fq& a, const Fq& b) {
  Fq r;
  r.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
  for (int i = 1; i < NL; i++) r.l[i] = subc_cc(a.l[i], b.l[i]);
  uint32_t mask = subc(0, 0);  // all ones when a < b
  r.l[0] = add_cc(r.l[0], G753_FC(FID).p[0] & mask);
#pragma unroll
  for (int i = 1; i < NL - 1; i++) r.l[i] = addc_cc(r.l[i], G753_FC(FID).p[i] & mask);
  r.l[NL - 1] = addc(r.l[NL - 1], G753_FC(FID).p[NL - 1] & mask);
  return r;
}

template <int FID>
G753_HD Fq fq_dbl(const Fq& a) {
  return fq_add<FID>(a, a);
}

template <int FID>
G753_HD Fq fq_neg(const Fq& a) {
  if (fq_is_zero(a)) return a;
  Fq r;
  r.l[0] = sub_cc(G753_FC(FID).p[0], a.l[0]);
#pragma unroll
  for (int i = 1; i < NL - 1; i++) r.l[i] = subc_cc(G753_FC(FID).p[i], a.l[i]);
  r.l[NL - 1] = subc(G753_FC(FID).p[NL - 1], a.l[NL - 1]);
  return r;
}

// k * a for a small compile-time k (curve / non-residue constants 2, 11, 13, 26, 121 are
// tiny once de-Montgomerised - SURVEY.md appendix A - so the reference's full-width
// constant multiplications become short addition chains with the same value).
template <int FID, unsigned K>
G753_HD Fq fq_mul_small(const Fq& a) {
  static_assert(K >= 1 && K < 256, "small constant");
  Fq acc = a;
  int top = 0;
  for (int i = 7; i >= 0; i--)
    if ((K >> i) & 1) {
      top = i;
      break;
    }
#pragma unroll
  for (int i = top - 1; i >= 0; i--) {
    acc = fq_dbl<FID>(acc);
    if ((K >> i) & 1) acc = fq_add<FID>(acc, a);
  }
  return acc;
}

// ------------------------------------------------------------------------------------
// Montgomery multiplication
// ------------------------------------------------------------------------------------
// acc[j], acc[j+1] = a[j] * bi   for j = 0, 2, ..., NL-2   (a may be offset by one limb)
G753_HD void row_mul(uint32_t* acc, const uint32_t* a, uint32_t bi) {
#pragma unroll
  for (int j = 0; j < NL; j += 2) {
    acc[j] = mul_lo(a[j], bi);
    acc[j + 1] = mul_hi(a[j], bi);
  }
}
// acc[0..NL) += sum_j a[j] * bi << (32 j), j even: one carry chain; carry out left in CC.CF
G753_HD void row_mad(uint32_t* acc, const uint32_t* a, uint32_t bi) {
  acc[0] = mad_lo_cc(a[0], bi, acc[0]);
  acc[1] = madc_hi_cc(a[0], bi, acc[1]);
#pragma unroll
  for (int j = 2; j < NL; j += 2) {
    acc[j] = madc_lo_cc(a[j], bi, acc[j]);
    acc[j + 1] = madc_hi_cc(a[j], bi, acc[j + 1]);
  }
}
// acc[j] = a[j]*bi (lo/hi as above) + acc[j+2] + CF : multiply-accumulate that also drops the
// accumulator two limbs (the one-limb Montgomery shift seen from the "odd" array)
G753_HD void row_mad_shift(uint32_t* acc, const uint32_t* a, uint32_t bi) {
#pragma unroll
  for (int j = 0; j < NL - 2; j += 2) {
    acc[j] = madc_lo_cc(a[j], bi, acc[j + 2]);
    acc[j + 1] = madc_hi_cc(a[j], bi, acc[j + 3]);
  }
  acc[NL - 2] = madc_lo_cc(a[NL - 2], bi, 0);
  acc[NL - 1] = madc_hi(a[NL - 1 - 1], bi, 0);
}

// One outer iteration: E holds columns 0..NL-1; O holds, at index k, column k-1 (index 0 dead).
// On exit E[0] == 0 and (E, O) are to be read swapped (O -> columns 0.., E index k -> column k-1).
template <int FID, bool FIRST>
G753_HD void mont_step(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi) {
  if (FIRST) {
    row_mul(O, a + 1, bi);
    row_mul(E, a, bi);
  } else {
    E[0] = add_cc(E[0], O[1]);
    row_mad_shift(O, a + 1, bi);
    row_mad(E, a, bi);
    O[NL - 1] = addc(O[NL - 1], 0);
  }
  uint32_t m = mul_lo(E[0], G753_FC(FID).inv32);
  row_mad(O, G753_FC(FID).p + 1, m);
  row_mad(E, G753_FC(FID).p, m);
  O[NL - 1] = addc(O[NL - 1], 0);
}

template <int FID>
G753_HD Fq fq_mul(const Fq& a, const Fq& b) {
  uint32_t even[NL], odd[NL];
  mont_step<FID, true>(even, odd, a.l, b.l[0]);
#pragma unroll
  for (int i = 1; i < NL; i += 2) {
    mont_step<FID, false>(odd, even, a.l, b.l[i]);
    if (i + 1 < NL) mont_step<FID, false>(even, odd, a.l, b.l[i + 1]);
  }
  // NL is even: after the last step `even` holds columns 0.. and odd[k] column k-1
  Fq r;
  r.l[0] = add_cc(even[0], odd[1]);
#pragma unroll
  for (int i = 1; i < NL - 1; i++) r.l[i] = addc_cc(even[i], odd[i + 1]);
  r.l[NL - 1] = addc(even[NL - 1], 0);
  fq_reduce_once<FID>(r.l);
  return r;
}

// Two products under ONE Montgomery reduction: (a b + c d) / R mod p, the lazy reduction of
// sums of products (a0 b0 + NR a1 b1 of an Fq2 product, R (Q - X3) - Y1 PPP of a mixed addition).
// 2 x 576 + 600 limb-MACs instead of 2 x 1176.  Same interleaved accumulators as fq_mul with a second
// row of partial products per step; the running total stays below 4 p 2^32 < 2^(32 (NL + 1)), so no
// chain carries out of its array (the argument of mont_step), and the result
// (a b + c d + M p) / R < p (1 + 2 p / R) < 2 p needs the same single conditional subtraction.
// The multipliers b, d are consumed one limb per step, so callers may stream them from shared
// memory (s_mul2, slots.cuh) and keep only a and c in registers.
template <int FID, bool FIRST>
G753_HD void mont2_step(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi, const uint32_t* c, uint32_t di) {
  if (FIRST) {
    row_mul(O, a + 1, bi);
    row_mul(E, a, bi);
  } else {
    E[0] = add_cc(E[0], O[1]);
    row_mad_shift(O, a + 1, bi);
    row_mad(E, a, bi);
    O[NL - 1] = addc(O[NL - 1], 0);
  }
  row_mad(O, c + 1, di);
  row_mad(E, c, di);
  O[NL - 1] = addc(O[NL - 1], 0);
  uint32_t m = mul_lo(E[0], G753_FC(FID).inv32);
  row_mad(O, G753_FC(FID).p + 1, m);
  row_mad(E, G753_FC(FID).p, m);
  O[NL - 1] = addc(O[NL - 1], 0);
}
// Three products under ONE reduction: (a b + c d + e f) / R mod p - one coefficient of an Fq3 product
// (fp3.rs:433-478 computes c0 = a0 b0 + nr (a1 b2 + a2 b1) etc.).  3 x 576 + 600 limb-MACs.  The running
// total stays below (a + c + e + p) 2^32 < 4 p 2^32 < 2^(32 (NL + 1)) like mont2_step's, and the result
// (a b + c d + e f + M p) / R < p (1 + 3 p / R) < 2 p needs the same single conditional subtraction.
template <int FID, bool FIRST>
G753_HD void mont3_step(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi, const uint32_t* c, uint32_t di,
                        const uint32_t* e, uint32_t fi) {
  if (FIRST) {
    row_mul(O, a + 1, bi);
    row_mul(E, a, bi);
  } else {
    E[0] = add_cc(E[0], O[1]);
    row_mad_shift(O, a + 1, bi);
    row_mad(E, a, bi);
    O[NL - 1] = addc(O[NL - 1], 0);
  }
  row_mad(O, c + 1, di);
  row_mad(E, c, di);
  O[NL - 1] = addc(O[NL - 1], 0);
  row_mad(O, e + 1, fi);
  row_mad(E, e, fi);
  O[NL - 1] = addc(O[NL - 1], 0);
  uint32_t m = mul_lo(E[0], G753_FC(FID).inv32);
  row_mad(O, G753_FC(FID).p + 1, m);
  row_mad(E, G753_FC(FID).p, m);
  O[NL - 1] = addc(O[NL - 1], 0);
}
// the accumulators after the last step -> canonical element (NL is even: `even` holds the columns)
template <int FID>
G753_HD Fq mont_finish(const uint32_t* even, const uint32_t* odd) {
  Fq r;
  r.l[0] = add_cc(even[0], odd[1]);
#pragma unroll
  for (int i = 1; i < NL - 1; i++) r.l[i] = addc_cc(even[i], odd[i + 1]);
  r.l[NL - 1] = addc(even[NL - 1], 0);
  fq_reduce_once<FID>(r.l);
  return r;
}
template <int FID>
G753_HD Fq fq_mul2(const Fq& a, const Fq& b, const Fq& c, const Fq& d) {
  uint32_t even[NL], odd[NL];
  mont2_step<FID, true>(even, odd, a.l, b.l[0], c.l, d.l[0]);
#pragma unroll
  for (int i = 1; i < NL; i += 2) {
    mont2_step<FID, false>(odd, even, a.l, b.l[i], c.l, d.l[i]);
    if (i + 1 < NL) mont2_step<FID, false>(even, odd, a.l, b.l[i + 1], c.l, d.l[i + 1]);
  }
  return mont_finish<FID>(even, odd);
}

template <int FID>
G753_HD Fq fq_mul3(const Fq& a, const Fq& b, const Fq& c, const Fq& d, const Fq& e, const Fq& f) {
  uint32_t even[NL], odd[NL];
  mont3_step<FID, true>(even, odd, a.l, b.l[0], c.l, d.l[0], e.l, f.l[0]);
#pragma unroll
  for (int i = 1; i < NL; i += 2) {
    mont3_step<FID, false>(odd, even, a.l, b.l[i], c.l, d.l[i], e.l, f.l[i]);
    if (i + 1 < NL) mont3_step<FID, false>(even, odd, a.l, b.l[i + 1], c.l, d.l[i + 1], e.l, f.l[i + 1]);
  }
  return mont_finish<FID>(even, odd);
}

// Dedicated squaring (replaces Fp768::square_in_place, fp_768.rs:339-548: the reference also
// computes the cross products once, doubles them and adds the squares before reducing).
//   a^2 = 2 * sum_{i<j} a_i a_j 2^(32 (i+j)) + sum_i a_i^2 2^(64 i):   276 + 24 products instead of 576,
// then a separate word-by-word Montgomery reduction (576 products): 876 limb-MACs against the 1152
// of fq_mul, i.e. 25 % fewer instructions on the binding IMAD pipe.  Every product is still one
// mad.lo.cc / madc.hi.cc pair on a carry chain: within one row the products of equal parity of
// (i + j) occupy consecutive column pairs.  Carries leaving a chain are counted in small per-column
// counters and folded in with one addition chain at the end (IADD3s run on the idle ALU pipe).
template <int FID>
G753_HD Fq fq_sqr(const Fq& a) {
  uint32_t T[2 * NL];      // columns 0 .. 47 of the product
  uint32_t K[2 * NL + 2];  // deferred carries into columns (tiny counts)
#pragma unroll
  for (int c = 0; c < 2 * NL; c++) T[c] = 0;
#pragma unroll
  for (int c = 0; c < 2 * NL + 2; c++) K[c] = 0;
  // ---- cross products, each once -------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < NL - 1; i++) {
#pragma unroll
    for (int par = 1; par <= 2; par++) {       // j = i + par, i + par + 2, ...
      if (i + par >= NL) continue;
      int last = i + par;
      T[i + last] = mad_lo_cc(a.l[i], a.l[last], T[i + last]);
      T[i + last + 1] = madc_hi_cc(a.l[i], a.l[last], T[i + last + 1]);
#pragma unroll
      for (int j = i + par + 2; j < NL; j += 2) {
        T[i + j] = madc_lo_cc(a.l[i], a.l[j], T[i + j]);
        T[i + j + 1] = madc_hi_cc(a.l[i], a.l[j], T[i + j + 1]);
        last = j;
      }
      K[i + last + 2] += addc(0, 0);
    }
  }
  // ---- T = 2 T + 2 K + squares ----------------------------------------------------------------
  T[0] = add_cc(T[0], T[0]);
#pragma unroll
  for (int c = 1; c < 2 * NL; c++) T[c] = addc_cc(T[c], T[c]);
  T[0] = add_cc(T[0], K[0] << 1);
#pragma unroll
  for (int c = 1; c < 2 * NL; c++) T[c] = addc_cc(T[c], K[c] << 1);
  T[0] = mad_lo_cc(a.l[0], a.l[0], T[0]);
  T[1] = madc_hi_cc(a.l[0], a.l[0], T[1]);
#pragma unroll
  for (int i = 1; i < NL; i++) {
    T[2 * i] = madc_lo_cc(a.l[i], a.l[i], T[2 * i]);
    T[2 * i + 1] = madc_hi_cc(a.l[i], a.l[i], T[2 * i + 1]);
  }
  // ---- Montgomery reduction: 24 rows m_i * p, two parity chains per row --------------------------
#pragma unroll
  for (int c = 0; c < 2 * NL + 2; c++) K[c] = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    const uint32_t m = mul_lo(T[i], G753_FC(FID).inv32);
    T[i] = mad_lo_cc(m, G753_FC(FID).p[0], T[i]);
    T[i + 1] = madc_hi_cc(m, G753_FC(FID).p[0], T[i + 1]);
#pragma unroll
    for (int j = 2; j < NL; j += 2) {
      T[i + j] = madc_lo_cc(m, G753_FC(FID).p[j], T[i + j]);
      T[i + j + 1] = madc_hi_cc(m, G753_FC(FID).p[j], T[i + j + 1]);
    }
    K[i + NL] += addc(0, 0);
    T[i + 1] = mad_lo_cc(m, G753_FC(FID).p[1], T[i + 1]);
    T[i + 2] = madc_hi_cc(m, G753_FC(FID).p[1], T[i + 2]);
#pragma unroll
    for (int j = 3; j < NL; j += 2) {
      T[i + j] = madc_lo_cc(m, G753_FC(FID).p[j], T[i + j]);
      if (i + j + 1 < 2 * NL)
        T[i + j + 1] = madc_hi_cc(m, G753_FC(FID).p[j], T[i + j + 1]);
      // i = 23, j = 23: the high word would land in column 48; it is zero because T + M p < 2^1536
    }
    if (i + NL + 1 < 2 * NL + 2) K[i + NL + 1] += addc(0, 0);
  }
  Fq r;
  r.l[0] = add_cc(T[NL], K[NL]);
#pragma unroll
  for (int c = 1; c < NL - 1; c++) r.l[c] = addc_cc(T[NL + c], K[NL + c]);
  r.l[NL - 1] = addc(T[2 * NL - 1], K[2 * NL - 1]);
  fq_reduce_once<FID>(r.l);
  return r;
}

// Out-of-line entry points: the curve code calls the multiplier hundreds of times per group
// operation; one shared body per field keeps kernels small (and ptxas time sane) at the cost
// of a call per product (~5% of its ~1400 instructions).
template <int FID>
G753_NI void fq_mul_ni(Fq& r, const Fq& a, const Fq& b) {
  r = fq_mul<FID>(a, b);
}
template <int FID>
G753_HD Fq fq_mulc(const Fq& a, const Fq& b) {
  Fq r;
  fq_mul_ni<FID>(r, a, b);
  return r;
}
template <int FID>
G753_HD Fq fq_sqrc(const Fq& a) {
  Fq r;
  fq_mul_ni<FID>(r, a, a);
  return r;
}

// a^e for a 768-bit exponent given as limbs (used for Fermat inversion; not on the hot loop)
template <int FID>
G753_NI Fq fq_pow(const Fq& a, const uint32_t* e) {
  Fq r = fq_one<FID>();
  bool started = false;
  for (int i = NL * 32 - 1; i >= 0; i--) {
    if (started) r = fq_sqrc<FID>(r);
    if ((e[i >> 5] >> (i & 31)) & 1) {
      r = started ? fq_mulc<FID>(r, a) : a;
      started = true;
    }
  }
  return r;
}

// ------------------------------------------------------------------------------------
// Inversion.  The reference uses a binary extended Euclid (fp_768.rs:551-605); the inverse is
// unique, so any correct algorithm gives identical limbs.  Here: the Bernstein-Yang "safegcd"
// divsteps in the batched form libsecp256k1's modinv32 made popular - 30 divsteps at a time are
// decided on the low 30 bits of (f, g) alone and collected into a 2x2 matrix that is then applied to
// the full-width (f, g) and, modulo p, to (d, e).  Branch-free, a fixed 62 x 30 = 1860 divsteps
// (bound for 768-bit inputs: (45907 * 768 + 26313) / 19929 = 1771), about 30 multiplications' worth
// of instructions instead of the ~1130 products of a Fermat exponentiation.
// Numbers are 26 signed limbs of 30 bits.
// ------------------------------------------------------------------------------------
constexpr int SG_LIMBS = 26;
constexpr int32_t SG_M30 = (int32_t)(0xffffffffu >> 2);
struct Sg30 {
  int32_t v[SG_LIMBS];
};
struct SgTrans {
  int32_t u, v, q, r;
};

// limb i (30 bits) of a 768-bit little-endian u32 array
G753_HD int32_t sg_limb30(const uint32_t* x, int i) {
  const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
  uint64_t two = 0;
  if (w < NL) two = x[w];
  if (w + 1 < NL) two |= (uint64_t)x[w + 1] << 32;
  return (int32_t)((uint32_t)(two >> sh) & (uint32_t)SG_M30);
}

G753_HD int32_t sg_divsteps_30(int32_t zeta, uint32_t f0, uint32_t g0, SgTrans& t) {
  // u, v, q, r: the transition matrix, signed in [-2^30, 2^30], kept modulo 2^32
  uint32_t u = 1, v = 0, q = 0, r = 1;
  uint32_t f = f0, g = g0;
  for (int i = 0; i < 30; ++i) {
    uint32_t mask1 = (uint32_t)(zeta >> 31);   // zeta < 0
    uint32_t mask2 = 0u - (g & 1u);            // g odd
    uint32_t x = (f ^ mask1) - mask1, y = (u ^ mask1) - mask1, z = (v ^ mask1) - mask1;
    g += x & mask2;
    q += y & mask2;
    r += z & mask2;
    mask1 &= mask2;
    zeta = (int32_t)(((uint32_t)zeta ^ mask1) - 1u);
    f += g & mask1;
    u += q & mask1;
    v += r & mask1;
    g >>= 1;
    u <<= 1;
    v <<= 1;
  }
  t.u = (int32_t)u;
  t.v = (int32_t)v;
  t.q = (int32_t)q;
  t.r = (int32_t)r;
  return zeta;
}

// (f, g) <- t * (f, g) / 2^30  (exact)
G753_HD void sg_update_fg(Sg30& f, Sg30& g, const SgTrans& t) {
  const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
  int64_t cf = u * f.v[0] + v * g.v[0];
  int64_t cg = q * f.v[0] + r * g.v[0];
  cf >>= 30;
  cg >>= 30;
#pragma unroll
  for (int i = 1; i < SG_LIMBS; ++i) {
    const int64_t fi = f.v[i], gi = g.v[i];
    cf += u * fi + v * gi;
    cg += q * fi + r * gi;
    f.v[i - 1] = (int32_t)cf & SG_M30;
    cf >>= 30;
    g.v[i - 1] = (int32_t)cg & SG_M30;
    cg >>= 30;
  }
  f.v[SG_LIMBS - 1] = (int32_t)cf;
  g.v[SG_LIMBS - 1] = (int32_t)cg;
}

// (d, e) <- t * (d, e) / 2^30 mod p, both kept in (-2p, p)
template <int FID>
G753_HD void sg_update_de(Sg30& d, Sg30& e, const SgTrans& t) {
  const int32_t u = t.u, v = t.v, q = t.q, r = t.r;
  const uint32_t* P = G753_FC(FID).p;
  const uint32_t pinv30 = (0u - G753_FC(FID).inv32) & (uint32_t)SG_M30;  // p^-1 mod 2^30
  const int32_t sd = d.v[SG_LIMBS - 1] >> 31, se = e.v[SG_LIMBS - 1] >> 31;
  int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);
  int64_t cd = (int64_t)u * d.v[0] + (int64_t)v * e.v[0];
  int64_t ce = (int64_t)q * d.v[0] + (int64_t)r * e.v[0];
  md -= (int32_t)((pinv30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)SG_M30);
  me -= (int32_t)((pinv30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)SG_M30);
  cd += (int64_t)sg_limb30(P, 0) * md;
  ce += (int64_t)sg_limb30(P, 0) * me;
  cd >>= 30;
  ce >>= 30;
#pragma unroll
  for (int i = 1; i < SG_LIMBS; ++i) {
    const int64_t di = d.v[i], ei = e.v[i], pi = sg_limb30(P, i);
    cd += (int64_t)u * di + (int64_t)v * ei + pi * md;
    ce += (int64_t)q * di + (int64_t)r * ei + pi * me;
    d.v[i - 1] = (int32_t)cd & SG_M30;
    cd >>= 30;
    e.v[i - 1] = (int32_t)ce & SG_M30;
    ce >>= 30;
  }
  d.v[SG_LIMBS - 1] = (int32_t)cd;
  e.v[SG_LIMBS - 1] = (int32_t)ce;
}

// r in (-2p, p) -> sign-adjusted r in [0, p); sign < 0 requests negation
template <int FID>
G753_HD void sg_normalize(Sg30& r, int32_t sign) {
  const uint32_t* P = G753_FC(FID).p;
  int32_t cond_add = r.v[SG_LIMBS - 1] >> 31;
  const int32_t cond_neg = sign >> 31;
#pragma unroll
  for (int i = 0; i < SG_LIMBS; ++i) {
    int32_t x = r.v[i] + (sg_limb30(P, i) & cond_add);
    r.v[i] = (x ^ cond_neg) - cond_neg;
  }
#pragma unroll
  for (int i = 0; i < SG_LIMBS - 1; ++i) {
    r.v[i + 1] += r.v[i] >> 30;
    r.v[i] &= SG_M30;
  }
  cond_add = r.v[SG_LIMBS - 1] >> 31;
#pragma unroll
  for (int i = 0; i < SG_LIMBS; ++i) r.v[i] += sg_limb30(P, i) & cond_add;
#pragma unroll
  for (int i = 0; i < SG_LIMBS - 1; ++i) {
    r.v[i + 1] += r.v[i] >> 30;
    r.v[i] &= SG_M30;
  }
}

// x^-1 mod p as plain integers (x canonical, < p); 0 -> 0
template <int FID>
G753_NI void fq_modinv_raw(Fq& out, const Fq& x) {
  Sg30 d, e, f, g;
#pragma unroll
  for (int i = 0; i < SG_LIMBS; ++i) {
    d.v[i] = 0;
    e.v[i] = i == 0 ? 1 : 0;
    f.v[i] = sg_limb30(G753_FC(FID).p, i);
    g.v[i] = sg_limb30(x.l, i);
  }
  int32_t zeta = -1;
  for (int it = 0; it < 62; ++it) {
    SgTrans t;
    zeta = sg_divsteps_30(zeta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
    sg_update_de<FID>(d, e, t);
    sg_update_fg(f, g, t);
  }
  sg_normalize<FID>(d, f.v[SG_LIMBS - 1]);
  // 26 x 30 bits -> 24 x 32 bits
#pragma unroll
  for (int w = 0; w < NL; ++w) {
    const int bit = 32 * w, i = bit / 30, sh = bit % 30;
    uint64_t acc = (uint64_t)(uint32_t)d.v[i] >> sh;
    if (i + 1 < SG_LIMBS) acc |= (uint64_t)(uint32_t)d.v[i + 1] << (30 - sh);
    if (i + 2 < SG_LIMBS && 60 - sh < 32) acc |= (uint64_t)(uint32_t)d.v[i + 2] << (60 - sh);
    out.l[w] = (uint32_t)acc;
  }
}

// a^-1 (Montgomery form in, Montgomery form out): (aR)^-1 = a^-1 R^-1, times R^3 / R = a^-1 R
template <int FID>
G753_HD Fq fq_inv(const Fq& a) {
  Fq t;
  fq_modinv_raw<FID>(t, a);
  Fq r3;
#pragma unroll
  for (int i = 0; i < NL; i++) r3.l[i] = G753_FC(FID).r3[i];
  return fq_mul<FID>(t, r3);
}

template <int FID>
G753_HD Fq fq_to_mont(const Fq& a) {
  Fq r2;
#pragma unroll
  for (int i = 0; i < NL; i++) r2.l[i] = G753_FC(FID).r2[i];
  return fq_mul<FID>(a, r2);
}
template <int FID>
G753_HD Fq fq_from_mont(const Fq& a) {
  Fq one;
#pragma unroll
  for (int i = 0; i < NL; i++) one.l[i] = (i == 0) ? 1u : 0u;
  return fq_mul<FID>(a, one);
}

}  // namespace g753
