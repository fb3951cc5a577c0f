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

// Fq2 products of the two-lane accumulation tower: 0 = two rounds of plain products, 1 = one
// two-product call per lane (lazy reduction), 2 = squarings through the same body as well
#ifndef G753_FQ2_LAZY
#define G753_FQ2_LAZY 2
#endif
// Products on slots through a ROLLED multiplier (fq_mul_rolled below): ~9 KB of code per product instead
// of ~45 KB, so that more than 8 warps per SM stop out-running instruction fetch.
// Opt-in: measured on a B200 it removes the fetch stalls but is ~2 % slower than the default at equal work
// (229.1 against 224.3 ms at 2^22); bit-exact either way (tests/test_pipeline_emul.py runs both).
#ifndef G753_ROLLED
#define G753_ROLLED 0
#endif

namespace g753 {

#if defined(G753_HOST_EMUL)
// test-only host build: one static buffer stands in for the block's shared memory
static uint4 g_slots[(240 * 1024) / sizeof(uint4)];
#else
extern __shared__ uint4 g_slots[];
#endif

constexpr int SLOT_CHUNKS = NL / 4;  // 6 x 16 bytes per Fq

// Thread <-> slot-column layout of a block.  NC columns (one curve operation each) of TP lanes:
//   TP = 1: one thread per column (prime-field curves);
//   TP = 2 / 8: the lanes of a column share its slots and split the independent base-field
//   products of an Fq2 / Fq3 operation between them (roles), so the extension-field curves put
//   2x / 8x more warps on an SM for the same shared-memory footprint.  Within a warp the role is
//   the slow index (lane = role * (32 / TP) + column-in-warp): a quarter-warp touches
//   consecutive columns of one slot, keeping LDS.128 / STS.128 conflict-free.
// A "slot type" T below is either a plain int-like column count (legacy spelling Lay<T, 1>) or Lay.
//   TP = 3 (Fq3 accumulation): ten columns per warp on lanes 0 .. 29, one coefficient per lane; lanes 30
//   and 31 are idle (role() == TP: kernels return at once for them, see idle()).
template <int NC_, int TP_>
struct Lay {
  static constexpr int NC = NC_, TP = TP_, CPW = 32 / TP_;  // CPW: columns per warp
  static constexpr int THREADS = TP_ == 1 ? NC_ : NC_ / CPW * 32;   // whole warps (= NC * TP when TP divides 32)
  static_assert(TP_ == 1 || NC_ % CPW == 0, "columns per block must fill whole warps");
  static G753_D int col() {
    if (TP == 1) return (int)threadIdx.x;
    return (int)(threadIdx.x >> 5) * CPW + (int)((threadIdx.x & 31) % CPW);
  }
  static G753_D int role() {
    if (TP == 1) return 0;
    return (int)((threadIdx.x & 31) / CPW);
  }
  // lanes that belong to no column (TP = 3: lanes 30, 31 of every warp)
  static G753_D bool idle() { return (32 % TP) != 0 && role() >= TP; }
  // reconverge the lanes of this column and order their slot accesses
  static G753_D void sync() {
#if defined(__CUDA_ARCH__)
    if (TP > 1) {
      unsigned m = 0;
#pragma unroll
      for (int r = 0; r < TP; r++) m |= 1u << (r * CPW);
      __syncwarp(m << ((threadIdx.x & 31) % CPW));
    }
#endif
  }
  // index of the column over the whole grid: the work item of this lane's column
  static G753_D unsigned item() { return blockIdx.x * NC + (unsigned)col(); }
};

template <class L>
G753_D uint4* slot_ptr(int s) {
  return g_slots + s * (SLOT_CHUNKS * L::NC) + L::col();
}
template <class T>
G753_D Fq s_ld(int s) {
  const uint4* p = slot_ptr<T>(s);
  Fq r;
#pragma unroll
  for (int k = 0; k < SLOT_CHUNKS; k++) {
    uint4 v = p[k * T::NC];
    r.l[4 * k] = v.x;
    r.l[4 * k + 1] = v.y;
    r.l[4 * k + 2] = v.z;
    r.l[4 * k + 3] = v.w;
  }
  return r;
}
template <class T>
G753_D void s_st(int s, const Fq& a) {
  uint4* p = slot_ptr<T>(s);
#pragma unroll
  for (int k = 0; k < SLOT_CHUNKS; k++) {
    uint4 v;
    v.x = a.l[4 * k];
    v.y = a.l[4 * k + 1];
    v.z = a.l[4 * k + 2];
    v.w = a.l[4 * k + 3];
    p[k * T::NC] = v;
  }
}
// global <-> register, 16-byte vector accesses (Fq is 16-byte aligned)
G753_HD Fq g_ld(const Fq* g) {
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

// a * slot[b] with the 24 Montgomery steps as a loop of six iterations of four (one 16-byte chunk of
// the multiplier per iteration, read from the slot when it is needed): the loop body is all the code
// there is.  Same accumulators and step as fq_mul; the first step runs on zeroed accumulators
// instead of the multiply-only form (same instruction count).
template <int FID, class T>
G753_D Fq fq_mul_rolled(const Fq& a, const uint4* pb) {
  uint32_t even[NL], odd[NL];
#pragma unroll
  for (int i = 0; i < NL; i++) even[i] = odd[i] = 0;
  uint4 nb = pb[0];
#pragma unroll 1
  for (int k = 0; k < SLOT_CHUNKS; k++) {
    const uint4 vb = nb;
    if (k + 1 < SLOT_CHUNKS) nb = pb[(k + 1) * T::NC];  // the next chunk arrives during these four steps
    mont_step<FID, false>(even, odd, a.l, vb.x);
    mont_step<FID, false>(odd, even, a.l, vb.y);
    mont_step<FID, false>(even, odd, a.l, vb.z);
    mont_step<FID, false>(odd, even, a.l, vb.w);
  }
  return mont_finish<FID>(even, odd);
}
// slot[a] * slot[b] / slot[a]^2 by the build's multiplier
template <int FID, class T>
G753_D Fq s_prod(const Fq& a, int b) {
#if G753_ROLLED
  return fq_mul_rolled<FID, T>(a, slot_ptr<T>(b));
#else
  return fq_mul<FID>(a, s_ld<T>(b));
#endif
}

// ---- Fq on slots (d may alias a or b everywhere: operands are read before d is written) ----
template <int FID, class T>
G753_NI void s_mul(int d, int a, int b) {
  s_st<T>(d, s_prod<FID, T>(s_ld<T>(a), b));
}
template <int FID, class T>
G753_NI void s_sqr(int d, int a) {
#if G753_ROLLED
  s_st<T>(d, fq_mul_rolled<FID, T>(s_ld<T>(a), slot_ptr<T>(a)));  // one multiplier body in the hot code
#else
  s_st<T>(d, fq_sqr<FID>(s_ld<T>(a)));
#endif
}
template <int FID, class T>
G753_NI void s_add(int d, int a, int b) {
  s_st<T>(d, fq_add<FID>(s_ld<T>(a), s_ld<T>(b)));
}
template <int FID, class T>
G753_NI void s_sub(int d, int a, int b) {
  s_st<T>(d, fq_sub<FID>(s_ld<T>(a), s_ld<T>(b)));
}
template <int FID, class T>
G753_NI void s_dbl(int d, int a) {
  s_st<T>(d, fq_dbl<FID>(s_ld<T>(a)));
}
template <int FID, class T>
G753_NI void s_neg(int d, int a) {
  s_st<T>(d, fq_neg<FID>(s_ld<T>(a)));
}
template <int FID, class T, unsigned KK>
G753_NI void s_mul_small(int d, int a) {
  s_st<T>(d, fq_mul_small<FID, KK>(s_ld<T>(a)));
}
template <int FID, class T>
G753_NI void s_inv(int d, int a) {  // Fermat; the inverse is unique, so the limbs equal fp_768.rs:551-605's
  s_st<T>(d, fq_inv<FID>(s_ld<T>(a)));
}
template <class T>
G753_D void s_copy(int d, int a) {
  if (d == a) return;
  const uint4* p = slot_ptr<T>(a);
  uint4* q = slot_ptr<T>(d);
#pragma unroll
  for (int k = 0; k < SLOT_CHUNKS; k++) q[k * T::NC] = p[k * T::NC];
}
template <class T>
G753_D bool s_is_zero(int a) {
  const uint4* p = slot_ptr<T>(a);
  uint32_t t = 0;
#pragma unroll
  for (int k = 0; k < SLOT_CHUNKS; k++) {
    uint4 v = p[k * T::NC];
    t |= v.x | v.y | v.z | v.w;
  }
  return t == 0;
}
template <class T>
G753_D void s_set_zero(int d) {
  uint4* q = slot_ptr<T>(d);
  uint4 z;
  z.x = z.y = z.z = z.w = 0;
#pragma unroll
  for (int k = 0; k < SLOT_CHUNKS; k++) q[k * T::NC] = z;
}
template <int FID, class T>
G753_D void s_set_one(int d) {
  s_st<T>(d, fq_one<FID>());
}
template <class T>
G753_D void s_ldg(int d, const Fq* g) {
  const uint4* p = (const uint4*)g;
  uint4* q = slot_ptr<T>(d);
#pragma unroll
  for (int k = 0; k < SLOT_CHUNKS; k++) q[k * T::NC] = p[k];
}
template <class T>
G753_D void s_stg(Fq* g, int a) {
  uint4* p = (uint4*)g;
  const uint4* q = slot_ptr<T>(a);
#pragma unroll
  for (int k = 0; k < SLOT_CHUNKS; k++) p[k] = q[k * T::NC];
}

// N tower elements global -> slots with all the loads in flight at once (one trip to L2 / HBM instead of
// N): element n of tower M goes from g[n] to slot d[n].  Lane roles as in M::ldg.
template <class M, int N>
G753_D void t_ldg_many(const int (&d)[N], const Fq* const (&g)[N]) {
  typedef typename M::T L;
  auto coefficient = [&](int c) {
    uint4 v[N][SLOT_CHUNKS];
#pragma unroll
    for (int n = 0; n < N; n++) {
      const uint4* p = (const uint4*)(g[n] + c);
#pragma unroll
      for (int k = 0; k < SLOT_CHUNKS; k++) v[n][k] = p[k];
    }
#pragma unroll
    for (int n = 0; n < N; n++) {
      uint4* q = slot_ptr<L>(d[n] + c);
#pragma unroll
      for (int k = 0; k < SLOT_CHUNKS; k++) q[k * L::NC] = v[n][k];
    }
  };
  if (L::TP == 1) {   // one lane per column holds every coefficient
    for (int c = 0; c < M::K; c++) coefficient(c);
  } else {
    const int r = L::role();
    if (r < M::K) coefficient(r);
  }
  L::sync();
}

// ---- two products under one reduction on slots ---------------------------------------------------
// returns a * slot[b] + c * slot[e] (fq_mul2).  a and c are in registers; the multipliers are
// streamed from their slots, one 16-byte chunk every four steps.
template <int FID, class T>
G753_D Fq s_mul2_stream(const Fq& a, int b, const Fq& c, int e) {
  const uint4* pb = slot_ptr<T>(b);
  const uint4* pe = slot_ptr<T>(e);
  uint32_t even[NL], odd[NL];
#if G753_ROLLED
#pragma unroll
  for (int i = 0; i < NL; i++) even[i] = odd[i] = 0;
  uint4 nb = pb[0], ne = pe[0];
#pragma unroll 1
  for (int k = 0; k < SLOT_CHUNKS; k++) {
    const uint4 vb = nb, ve = ne;
    if (k + 1 < SLOT_CHUNKS) {
      nb = pb[(k + 1) * T::NC];
      ne = pe[(k + 1) * T::NC];
    }
    mont2_step<FID, false>(even, odd, a.l, vb.x, c.l, ve.x);
    mont2_step<FID, false>(odd, even, a.l, vb.y, c.l, ve.y);
    mont2_step<FID, false>(even, odd, a.l, vb.z, c.l, ve.z);
    mont2_step<FID, false>(odd, even, a.l, vb.w, c.l, ve.w);
  }
#else
#pragma unroll
  for (int k = 0; k < SLOT_CHUNKS; k++) {
    const uint4 vb = pb[k * T::NC], ve = pe[k * T::NC];
    const uint32_t bb[4] = {vb.x, vb.y, vb.z, vb.w}, ee[4] = {ve.x, ve.y, ve.z, ve.w};
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (k == 0 && u == 0) mont2_step<FID, true>(even, odd, a.l, bb[0], c.l, ee[0]);
      else if (u & 1) mont2_step<FID, false>(odd, even, a.l, bb[u], c.l, ee[u]);
      else mont2_step<FID, false>(even, odd, a.l, bb[u], c.l, ee[u]);
    }
  }
#endif
  return mont_finish<FID>(even, odd);
}
// ---- the six-slot mixed addition of the prime-field curves (EcS::madd6_g) -----------------------
// madd-2008-s on a point in slots X, Y, ZZ, ZZZ = P .. P + 3 with only two temporaries t0, t1 = W, W + 1
// (six slots per thread instead of eight: three 128-thread blocks = 12 warps per SM instead of 8).
// Two things make that possible: PP = P^2 never gets a slot - it stays in registers as the common
// factor of the three products that use it - and the subtraction that follows a product is folded
// into the same step.  The whole addition is a short micro-program run by ONE function with ONE
// inlined product and ONE inlined squaring: a B200 fetches instructions fast enough only while the
// hot code of an SM stays around 100 KB (measured: the same formulas spread over five fused
// functions, each with its own 45 KB multiplier body, ran at 63 % instead of 85 % pipe utilisation
// with 6 stall cycles per issue waiting for instructions, profiles/r01_bucket_acc_2p22_v19).
//
//   phase 0:  t0 = x2 ZZ - X (P),  t1 = +-(y2 ZZZ) - Y (R);  returns bit 0 = (P == 0), bit 1 = (R == 0)
//   phase 1:  f = t0^2 (PP);  X *= f (Q);  ZZ *= f;  t0 *= f (PPP);  ZZZ *= t0;  Y *= t0;
//             t0 = t1^2 - t0;  X = t0 - 2X (X3), t0 = Q - X3;  Y = t1 t0 - Y (Y3)
//   (G753_ROLLED: Y *= t0 is skipped and the last step is Y = t1 (Q - X3) + (-Y) t0, two products under one
//   reduction: 600 limb-MACs saved, which pays for the two squarings that go through the product body)
template <int FID, class T>
G753_NI unsigned s_madd6(int P, int W, const Fq* q, bool negq, int phase) {
  enum { AG = 1, SQR = 2, KEEP = 4, TOF = 8, SUBC = 16, NEGADD = 32, XPOST = 64 };
  const int X = P, Y = P + 1, ZZ = P + 2, ZZZ = P + 3, t0 = W, t1 = W + 1;
  Fq f;  // retained factor
#pragma unroll
  for (int i = 0; i < NL; i++) f.l[i] = 0;
  unsigned zero = 0;
  const int first = phase == 0 ? 0 : 2, last = phase == 0 ? 2 : 10;
#pragma unroll 1
  for (int s = first; s < last; s++) {
    int d = 0, a = 0, b = 0, c = 0;
    unsigned fl = 0;
    switch (s) {
      case 0: d = t0; b = ZZ; c = X; fl = AG | SUBC; break;
      case 1: d = t1; b = ZZZ; c = Y; fl = AG | (negq ? NEGADD : SUBC); break;
      case 2: a = t0; fl = SQR | TOF; break;
      case 3: d = X; a = X; fl = KEEP; break;
      case 4: d = ZZ; a = ZZ; fl = KEEP; break;
      case 5: d = t0; a = t0; fl = KEEP; break;
      case 6: d = ZZZ; a = ZZZ; b = t0; break;
      case 7: d = Y; a = Y; b = t0; break;
      case 8: d = t0; a = t1; c = t0; fl = SQR | SUBC | XPOST; break;
      default: d = Y; a = t1; b = t0; c = Y; fl = SUBC; break;
    }
    Fq r;
#if G753_ROLLED
    // with the rolled bodies code size is no concern, so the last two products share one reduction:
    // Y3 = R (Q - X3) + (-Y1) PPP (fq_mul2).  Y1 and PPP stay in their slots until then: step 7 is
    // skipped, step 8 leaves Q - X3 in f instead of overwriting PPP.
    if (s == 7) continue;
    if (s == 9) {
      s_st<T>(Y, s_mul2_stream<FID, T>(f, t1, fq_neg<FID>(s_ld<T>(Y)), t0));
      continue;
    }
    // every product has one operand in a slot: the rolled multiplier streams that one and keeps the
    // other (the global operand, the retained factor f, or a copy of the slot) in registers
    if (fl & KEEP) {
      r = fq_mul_rolled<FID, T>(f, slot_ptr<T>(a));
    } else {
      const Fq A = (fl & AG) ? g_ld(q + s) : s_ld<T>(a);
      r = fq_mul_rolled<FID, T>(A, slot_ptr<T>((fl & SQR) ? a : b));
    }
#else
    const Fq A = (fl & AG) ? g_ld(q + s) : s_ld<T>(a);
    if (fl & SQR) {
      r = fq_sqr<FID>(A);
    } else {
      if (!(fl & KEEP)) f = s_ld<T>(b);
      r = fq_mul<FID>(A, f);
    }
#endif
    if (fl & TOF) {
      f = r;
      continue;
    }
    if (fl & (SUBC | NEGADD)) {
      const Fq cc = s_ld<T>(c);
      r = (fl & NEGADD) ? fq_neg<FID>(fq_add<FID>(r, cc)) : fq_sub<FID>(r, cc);
    }
    if (fl & XPOST) {  // r = R^2 - PPP: X3 = r - 2Q into X, Q - X3 into d
      const Fq qq = s_ld<T>(X);
      r = fq_sub<FID>(fq_sub<FID>(r, qq), qq);
      s_st<T>(X, r);
      r = fq_sub<FID>(qq, r);
#if G753_ROLLED
      f = r;
      continue;
#endif
    }
    s_st<T>(d, r);
    if (s < 2 && fq_is_zero(r)) zero |= 1u << s;
  }
  return zero;
}

// ---- two products under one reduction on slots: s_mul2 (s_mul2_stream is defined above s_madd6) ----
// d = slot[a] * slot[b] + f(slot[c]) * slot[e], f by `mode`: 0 identity, 1 negation, 2 times NR (the
// non-residue of an Fq2 product).  The lanes of a column reconverge before d is written (d may be an
// operand of the other lanes' products) and after.
template <int FID, class T, unsigned NR>
G753_NI void s_mul2(int d, int a, int b, int c, int e, int mode) {
  Fq cc = s_ld<T>(c);
  if (mode == 1) cc = fq_neg<FID>(cc);
  if (mode == 2) cc = fq_mul_small<FID, NR>(cc);
  const Fq r = s_mul2_stream<FID, T>(s_ld<T>(a), b, cc, e);
  T::sync();
  s_st<T>(d, r);
  T::sync();
}

// ---- three products under one reduction on slots: x0 * slot[y0] + x1 * slot[y1] + x2 * slot[y2] (fq_mul3);
// the multiplicands are in registers, the multipliers are streamed from their slots 16 bytes at a time
template <int FID, class T>
G753_D Fq s_mul3_stream(const Fq& x0, int y0, const Fq& x1, int y1, const Fq& x2, int y2) {
  const uint4* p0 = slot_ptr<T>(y0);
  const uint4* p1 = slot_ptr<T>(y1);
  const uint4* p2 = slot_ptr<T>(y2);
  uint32_t even[NL], odd[NL];
#pragma unroll
  for (int k = 0; k < SLOT_CHUNKS; k++) {
    const uint4 v0 = p0[k * T::NC], v1 = p1[k * T::NC], v2 = p2[k * T::NC];
    const uint32_t b0[4] = {v0.x, v0.y, v0.z, v0.w}, b1[4] = {v1.x, v1.y, v1.z, v1.w}, b2[4] = {v2.x, v2.y, v2.z, v2.w};
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (k == 0 && u == 0) mont3_step<FID, true>(even, odd, x0.l, b0[0], x1.l, b1[0], x2.l, b2[0]);
      else if (u & 1) mont3_step<FID, false>(odd, even, x0.l, b0[u], x1.l, b1[u], x2.l, b2[u]);
      else mont3_step<FID, false>(even, odd, x0.l, b0[u], x1.l, b1[u], x2.l, b2[u]);
    }
  }
  return mont_finish<FID>(even, odd);
}

// ---- towers: an element is K consecutive slots; `t` is the first of NTMP scratch slots ----
template <int FID, class T_>
struct Tw1 {
  typedef T_ T;
  static constexpr int K = 1, NTMP = 0, FIELD = FID;
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

template <int FID, class T_, unsigned NR>
struct Tw2 {
  typedef T_ T;
  static constexpr int K = 2, NTMP = 3, FIELD = FID;
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

template <int FID, class T_, unsigned NR>
struct Tw3 {
  typedef T_ T;
  static constexpr int K = 3, NTMP = 6, FIELD = FID;
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


#if !defined(G753_HOST_EMUL)
// ---- cooperative towers (device only): the TP lanes of a column split the independent base-field
// products of one extension-field operation.  Lanes that run the SAME function on different
// slots proceed in lockstep; role-specific steps are serialised by the SIMT hardware, so every
// phase below is written as "one call, role-dependent slot numbers" wherever it matters
// (the products).  Every operation ends with L::sync(): its result is visible to all lanes of the
// column before the next one starts.  d may alias a or b: d is written only after the last read
// of the operands.  The TEST-ONLY host build keeps the one-thread towers above.
template <int FID, class L, unsigned NR>
struct Tw2C {
  typedef L T;
  static constexpr int K = 2, NTMP = 4, FIELD = FID;
  static_assert(L::TP == 2 || L::TP == 4, "Fq2 runs on two or four lanes");
  // TP = 4: Karatsuba (fp2.rs:387-401), products a0 b0, a1 b1, (a0 + a1)(b0 + b1) on lanes 0, 1, 2.
  // TP = 2: schoolbook in two rounds, a0 b0 | a1 b1 then a0 b1 | a1 b0 - no idle lane in a product
  //         round and one addition less; lane time 2 x 2.4 against 4 x 1.55 products.
  static G753_NI void mul(int d, int a, int b, int t) {
    const int r = L::role();
    if (L::TP == 2) {
#if G753_FQ2_LAZY
      // one lane per coefficient, each a sum of two products under one reduction (s_mul2):
      // a0 b0 + (NR a1) b1 | a0 b1 + a1 b0 - 1752 limb-MACs of lane time instead of 2 x 1176
      s_mul2<FID, L, NR>(d + r, a, b + r, a + 1, b + 1 - r, r == 0 ? 2 : 0);
#else
      s_mul<FID, L>(t + r, a + r, b + r);              // a0 b0 | a1 b1
      s_mul<FID, L>(t + 2 + r, a + r, b + 1 - r);      // a0 b1 | a1 b0
      L::sync();
      if (r == 0) {
        s_mul_small<FID, L, NR>(t + 1, t + 1);
        s_add<FID, L>(d, t, t + 1);
      } else {
        s_add<FID, L>(d + 1, t + 2, t + 3);
      }
      L::sync();
#endif
      return;
    }
    if (r == 2 || r == 3) s_add<FID, L>(t + r, r == 2 ? a : b, r == 2 ? a + 1 : b + 1);
    L::sync();
    if (r < 3) s_mul<FID, L>(t + r, r == 0 ? a : r == 1 ? a + 1 : t + 2, r == 0 ? b : r == 1 ? b + 1 : t + 3);
    L::sync();
    if (r == 1) {
      s_sub<FID, L>(d + 1, t + 2, t);
      s_sub<FID, L>(d + 1, d + 1, t + 1);
    } else if (r == 0) {
      s_mul_small<FID, L, NR>(t + 3, t + 1);
      s_add<FID, L>(d, t, t + 3);
    }
    L::sync();
  }
  // complex squaring (fp2.rs:128-144): a0 a1 and (a0 + a1)(a0 + NR a1) on lanes 0, 1
  static G753_NI void sqr(int d, int a, int t) {
    const int r = L::role();
#if G753_FQ2_LAZY >= 2
    if (L::TP == 2) {
      // through the same two-product body as mul (one multiplier body in the accumulation kernel's
      // hot code): a0^2 + (NR a1) a1 | a0 a1 + a1 a0
      s_mul2<FID, L, NR>(d + r, a, a + r, a + 1, a + 1 - r, r == 0 ? 2 : 0);
      return;
    }
#endif
    if (r == 0) s_add<FID, L>(t + 1, a, a + 1);
    if (r == 1) {
      s_mul_small<FID, L, NR>(t + 2, a + 1);
      s_add<FID, L>(t + 2, a, t + 2);
    }
    L::sync();
    if (r < 2) s_mul<FID, L>(t + r, r == 0 ? a : t + 1, r == 0 ? a + 1 : t + 2);
    L::sync();
    if (r == 0) {
      s_dbl<FID, L>(d + 1, t);
    } else if (r == 1) {
      s_mul_small<FID, L, NR>(t + 2, t);
      s_sub<FID, L>(t + 1, t + 1, t);
      s_sub<FID, L>(d, t + 1, t + 2);
    }
    L::sync();
  }
  static G753_NI void inv(int d, int a, int t) {
    if (L::role() == 0) Tw2<FID, L, NR>::inv(d, a, t);
    L::sync();
  }
  static G753_D void add(int d, int a, int b) {
    const int r = L::role();
    if (r < K) s_add<FID, L>(d + r, a + r, b + r);
    L::sync();
  }
  static G753_D void sub(int d, int a, int b) {
    const int r = L::role();
    if (r < K) s_sub<FID, L>(d + r, a + r, b + r);
    L::sync();
  }
  static G753_D void dbl(int d, int a) {
    const int r = L::role();
    if (r < K) s_dbl<FID, L>(d + r, a + r);
    L::sync();
  }
  static G753_D void neg(int d, int a) {
    const int r = L::role();
    if (r < K) s_neg<FID, L>(d + r, a + r);
    L::sync();
  }
  static G753_D void copy(int d, int a) {
    const int r = L::role();
    if (r < K) s_copy<L>(d + r, a + r);
    L::sync();
  }
  static G753_D bool is_zero(int a) { return s_is_zero<L>(a) && s_is_zero<L>(a + 1); }
  static G753_D void set_zero(int d) {
    const int r = L::role();
    if (r < K) s_set_zero<L>(d + r);
    L::sync();
  }
  static G753_D void set_one(int d) {
    const int r = L::role();
    if (r == 0) s_set_one<FID, L>(d);
    if (r == 1) s_set_zero<L>(d + 1);
    L::sync();
  }
  static G753_D void ldg(int d, const Fq* g) {
    const int r = L::role();
    if (r < K) s_ldg<L>(d + r, g + r);
    L::sync();
  }
  static G753_D void stg(Fq* g, int a) {
    const int r = L::role();
    if (r < K) s_stg<L>(g + r, a + r);
  }
  // coefficient-wise multiplication by small constants (curve coefficient a of the twists)
  template <unsigned C0, unsigned C1>
  static G753_D void mul_small2(int d, int a) {
    const int r = L::role();
    if (r == 0) s_mul_small<FID, L, C0>(d, a);
    if (r == 1) s_mul_small<FID, L, C1>(d + 1, a + 1);
    L::sync();
  }
};

template <int FID, class L, unsigned NR>
struct Tw3C {
  typedef L T;
  static constexpr int K = 3, NTMP = 9, FIELD = FID;
  static_assert(L::TP == 4 || L::TP == 8, "Fq3 runs on four or eight lanes");
  // Karatsuba (fp3.rs:451-478): v0 = a0 b0, v1 = a1 b1, v2 = a2 b2 and the cross products
  // m12 = (a1+a2)(b1+b2), m01 = (a0+a1)(b0+b1), m02 = (a0+a2)(b0+b2).
  //   TP = 8: all six products in one round (lanes 0..5);
  //   TP = 4: v0, v1, v2, m12 in the first round, m01, m02 (lanes 1, 2) in the second.
  // scratch: t+0..2 = v0, v1, v2; t+3.. = the operand sums of m12, m01, m02 (product over the first)
  static G753_NI void mul(int d, int a, int b, int t) {
    const int r = L::role();
    const int j = L::TP == 8 ? r - 3 : (r == 3 ? 0 : r);   // which cross product this lane prepares
    if (L::TP == 8 ? (r >= 3 && r < 6) : (r >= 1)) {
      const int x = j == 0 ? 1 : 0, y = j == 1 ? 1 : 2;     // j = 0: (1,2)  1: (0,1)  2: (0,2)
      s_add<FID, L>(t + 3 + 2 * j, a + x, a + y);
      s_add<FID, L>(t + 4 + 2 * j, b + x, b + y);
    }
    L::sync();
    if (L::TP == 8) {
      if (r < 6) s_mul<FID, L>(r < 3 ? t + r : t + 3 + 2 * j, r < 3 ? a + r : t + 3 + 2 * j, r < 3 ? b + r : t + 4 + 2 * j);
    } else {
      s_mul<FID, L>(r < 3 ? t + r : t + 3, r < 3 ? a + r : t + 3, r < 3 ? b + r : t + 4);
      if (r == 1 || r == 2) s_mul<FID, L>(t + 3 + 2 * r, t + 3 + 2 * r, t + 4 + 2 * r);
    }
    L::sync();
    // m12 = t+3, m01 = t+5, m02 = t+7; one output coordinate per lane 0..2
    if (r < 3) {
      const int m = t + 3 + 2 * r;
      s_sub<FID, L>(m, m, r == 0 ? t + 1 : t);                // m12 - v1 | m01 - v0 | m02 - v0
      s_sub<FID, L>(m, m, r == 1 ? t + 1 : t + 2);            // ... - v2 | ... - v1 | ... - v2
      if (r < 2) s_mul_small<FID, L, NR>(r == 0 ? m : m + 1, r == 0 ? m : t + 2);   // NR m12 | NR v2 -> t+6
      s_add<FID, L>(d + r, r == 0 ? t : m, r == 0 ? m : r == 1 ? m + 1 : t + 1);
    }
    L::sync();
  }
  // Chung-Hasan SQR2 (fp3.rs:165-185): a0^2, a0 a1, (a0 - a1 + a2)^2, a1 a2, a2^2 on lanes 0..4
  // (TP = 4: the fifth product in a second round on lane 0)
  static G753_NI void sqr(int d, int a, int t) {
    const int r = L::role();
    if (r == 2) {
      s_sub<FID, L>(t + 5, a, a + 1);
      s_add<FID, L>(t + 5, t + 5, a + 2);
    }
    L::sync();
    if (r < (L::TP == 8 ? 5 : 4)) {
      const int x = r == 0 ? a : r == 1 ? a : r == 2 ? t + 5 : r == 3 ? a + 1 : a + 2;
      const int y = r == 0 ? a : r == 1 ? a + 1 : r == 2 ? t + 5 : r == 3 ? a + 2 : a + 2;
      s_mul<FID, L>(t + r, x, y);
    }
    if (L::TP == 4 && r == 0) s_mul<FID, L>(t + 4, a + 2, a + 2);
    L::sync();
    if (r == 1 || r == 3) s_dbl<FID, L>(t + r, t + r);        // s1 = 2 a0 a1, s3 = 2 a1 a2
    L::sync();
    if (r < 2) {
      s_mul_small<FID, L, NR>(t + 6 + r, r == 0 ? t + 3 : t + 4);   // NR s3 | NR s4
      s_add<FID, L>(d + r, r == 0 ? t : t + 1, t + 6 + r);          // c0 = s0 + NR s3 | c1 = s1 + NR s4
    } else if (r == 2) {
      s_add<FID, L>(t + 5, t + 1, t + 2);
      s_add<FID, L>(t + 5, t + 5, t + 3);
      s_sub<FID, L>(t + 5, t + 5, t);
      s_sub<FID, L>(d + 2, t + 5, t + 4);                           // c2 = s1 + s2 + s3 - s0 - s4
    }
    L::sync();
  }
  static G753_NI void inv(int d, int a, int t) {
    if (L::role() == 0) Tw3<FID, L, NR>::inv(d, a, t);
    L::sync();
  }
  static G753_D void add(int d, int a, int b) {
    const int r = L::role();
    if (r < K) s_add<FID, L>(d + r, a + r, b + r);
    L::sync();
  }
  static G753_D void sub(int d, int a, int b) {
    const int r = L::role();
    if (r < K) s_sub<FID, L>(d + r, a + r, b + r);
    L::sync();
  }
  static G753_D void dbl(int d, int a) {
    const int r = L::role();
    if (r < K) s_dbl<FID, L>(d + r, a + r);
    L::sync();
  }
  static G753_D void neg(int d, int a) {
    const int r = L::role();
    if (r < K) s_neg<FID, L>(d + r, a + r);
    L::sync();
  }
  static G753_D void copy(int d, int a) {
    const int r = L::role();
    if (r < K) s_copy<L>(d + r, a + r);
    L::sync();
  }
  static G753_D bool is_zero(int a) { return s_is_zero<L>(a) && s_is_zero<L>(a + 1) && s_is_zero<L>(a + 2); }
  static G753_D void set_zero(int d) {
    const int r = L::role();
    if (r < K) s_set_zero<L>(d + r);
    L::sync();
  }
  static G753_D void set_one(int d) {
    const int r = L::role();
    if (r == 0) s_set_one<FID, L>(d);
    if (r == 1 || r == 2) s_set_zero<L>(d + r);
    L::sync();
  }
  static G753_D void ldg(int d, const Fq* g) {
    const int r = L::role();
    if (r < K) s_ldg<L>(d + r, g + r);
    L::sync();
  }
  static G753_D void stg(Fq* g, int a) {
    const int r = L::role();
    if (r < K) s_stg<L>(g + r, a + r);
  }
};

// Fq3 on THREE lanes, one coefficient per lane, schoolbook with lazy reduction (fp3.rs:433-478 gives the
// same values): c0 = a0 b0 + nr (a1 b2 + a2 b1), c1 = a0 b1 + a1 b0 + nr a2 b2, c2 = a0 b2 + a1 b1 + a2 b0 -
// every coefficient is ONE three-product body (s_mul3_stream: 3 x 576 + 600 limb-MACs), no lane idles during a
// product and no temporaries are needed, so a warp carries ten columns (lanes 0 .. 29) instead of the eight of
// the four-lane Karatsuba form whose second round leaves two (squaring: three) lanes idle.  The squaring goes
// through the same body (one multiplier body in the accumulation kernel's hot code).  The scratch slots
// (NTMP) serve inv only, which runs on lane 0 with the one-thread tower.
template <int FID, class L, unsigned NR>
struct Tw3L {
  typedef L T;
  static constexpr int K = 3, NTMP = 6, FIELD = FID;
  static_assert(L::TP == 3, "one lane per coefficient");
  static G753_NI void mul(int d, int a, int b, int) {
    const int r = L::role();
    // lane r:  x0 * b[r]  +  x1 * b[(r + 2) % 3]  +  x2 * b[(r + 1) % 3],  x_i = a_i times nr where the
    // product wraps around u^3 = nr: (r = 0: a1, a2;  r = 1: a2)
    const Fq x0 = s_ld<L>(a);
    Fq x1 = s_ld<L>(a + 1), x2 = s_ld<L>(a + 2);
    const Fq n1 = fq_mul_small<FID, NR>(x1), n2 = fq_mul_small<FID, NR>(x2);
#pragma unroll
    for (int i = 0; i < NL; i++) {
      x1.l[i] = r == 0 ? n1.l[i] : x1.l[i];
      x2.l[i] = r <= 1 ? n2.l[i] : x2.l[i];
    }
    const int y0 = b + r, y1 = b + (r == 0 ? 2 : r - 1), y2 = b + (r == 2 ? 0 : r + 1);
    const Fq res = s_mul3_stream<FID, L>(x0, y0, x1, y1, x2, y2);
    L::sync();               // every lane has read a and b: d may alias either
    s_st<L>(d + r, res);
    L::sync();
  }
  static G753_D void sqr(int d, int a, int t) { mul(d, a, a, t); }
  static G753_NI void inv(int d, int a, int t) {
    if (L::role() == 0) Tw3<FID, L, NR>::inv(d, a, t);
    L::sync();
  }
  static G753_D void add(int d, int a, int b) {
    const int r = L::role();
    s_add<FID, L>(d + r, a + r, b + r);
    L::sync();
  }
  static G753_D void sub(int d, int a, int b) {
    const int r = L::role();
    s_sub<FID, L>(d + r, a + r, b + r);
    L::sync();
  }
  static G753_D void dbl(int d, int a) {
    const int r = L::role();
    s_dbl<FID, L>(d + r, a + r);
    L::sync();
  }
  static G753_D void neg(int d, int a) {
    const int r = L::role();
    s_neg<FID, L>(d + r, a + r);
    L::sync();
  }
  static G753_D void copy(int d, int a) {
    const int r = L::role();
    s_copy<L>(d + r, a + r);
    L::sync();
  }
  static G753_D bool is_zero(int a) { return s_is_zero<L>(a) && s_is_zero<L>(a + 1) && s_is_zero<L>(a + 2); }
  static G753_D void set_zero(int d) {
    s_set_zero<L>(d + L::role());
    L::sync();
  }
  static G753_D void set_one(int d) {
    const int r = L::role();
    if (r == 0) s_set_one<FID, L>(d);
    else s_set_zero<L>(d + r);
    L::sync();
  }
  static G753_D void ldg(int d, const Fq* g) {
    const int r = L::role();
    s_ldg<L>(d + r, g + r);
    L::sync();
  }
  static G753_D void stg(Fq* g, int a) {
    const int r = L::role();
    s_stg<L>(g + r, a + r);
  }
};
#endif

}  // namespace g753
