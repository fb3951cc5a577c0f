// Warp-cooperative group law for the LATENCY-BOUND steps of the MSM: the Horner fold over the window
// sums (variable_base.rs:72-82), the fold of per-GPU partial points, the tail of the bucket reduction.
//
// Why: one thread needs ~2.4 us per 753-bit Montgomery product - its 2352 IMAD issue slots go through
// ONE warp scheduler whatever the other 31 lanes do - so the reference's serial chain of 753 doublings
// costs 31-37 ms on one thread while 147 SMs idle.  Here one field product is spread over the 8 lanes
// of an "octet" (3 x 32-bit limbs per lane):
//     digit-serial Montgomery multiplication in base 2^96, one digit of b per round, 8 rounds:
//       t += a_lane * B_j            (B_j broadcast from lane j)
//       M  = t_lane0 * (-p^-1) mod 2^96, broadcast
//       t += p_lane * M              (lane 0's low digit is now zero)
//       shift one digit = one lane:  lane l keeps its own high half and receives lane l+1's low half
//     carries between lanes stay in a small per-lane word and are resolved once at the end with a
//     generate / propagate pass over the octet's ballot.
// and the four octets of a warp run up to FOUR independent products of a curve formula at a time.
// Values are lazily reduced (R / p >= 2^15: products of inputs below x p, y p stay below
// (1 + x y / 2^15) p; a - b is a + 2^k p - b), so no conditional subtraction is on the hot path.
// The formulas (XYZZ doubling / addition / mixed addition over Fq, Fq2, Fq3, conversions) are compiled
// into straight-line micro-programs by tools/gen_coop.py - rows of four instructions on shared-memory
// slots - whose bit-exact model runs in the CPU test tier (tests/test_coop_model.py); this file
// interprets them.  One warp works on one point.
#pragma once
#include "device.cuh"
#include "fq.cuh"

#if !defined(G753_HOST_EMUL)
#include "coop_programs.inc"

namespace g753 {

constexpr unsigned COOP_FULL = 0xffffffffu;
constexpr unsigned COOP_HAS_MUL = 1u << 28, COOP_HAS_LIN = 1u << 29;
enum { COOP_NOP = 0, COOP_MUL = 1, COOP_LIN = 2 };
constexpr int COOP_KP_ROWS = 12;                     // 2^0 .. 2^11 times p
constexpr int COOP_CONST_WORDS = COOP_KP_ROWS * 24;  // shared-memory words in front of a block's slots

template <int FID> struct CoopField;
template <> struct CoopField<0> {
  static G753_D const uint32_t* kp() { return &COOP_KP_0[0][0]; }
  static G753_D const uint32_t* np() { return COOP_NP96_0; }
};
template <> struct CoopField<1> {
  static G753_D const uint32_t* kp() { return &COOP_KP_1[0][0]; }
  static G753_D const uint32_t* np() { return COOP_NP96_1; }
};

G753_D unsigned coop_lane() { return threadIdx.x & 7u; }
G753_D unsigned coop_octet() { return (threadIdx.x & 31u) >> 3; }

// Clean lanes of sum_l (r_l + cw_l 2^96) 2^(96 l) mod 2^768: every lane adds its lower neighbour's carry
// word, then the 0 / 1 carries that remain ripple through one generate / propagate evaluation on the
// octet's ballot.  Returns the carry out of the top lane (same value in all lanes of the octet).
G753_D uint32_t coop_resolve(uint32_t* r, uint32_t cw) {
  const unsigned l = coop_lane();
  uint32_t cin = __shfl_up_sync(COOP_FULL, cw, 1, 8);
  if (l == 0) cin = 0;
  r[0] = add_cc(r[0], cin);
  r[1] = addc_cc(r[1], 0);
  r[2] = addc_cc(r[2], 0);
  const uint32_t g = addc(0, 0);
  const bool p = (r[0] & r[1] & r[2]) == 0xffffffffu;
  const unsigned G = __ballot_sync(COOP_FULL, g != 0), P = __ballot_sync(COOP_FULL, p);
  const unsigned sh = threadIdx.x & 24u;
  const unsigned g8 = (G >> sh) & 0xffu, p8 = (P >> sh) & 0xffu;
  const unsigned x = p8 | g8, sum = x + g8;
  const unsigned mine = ((sum ^ x ^ g8) >> l) & 1u;
  r[0] = add_cc(r[0], mine);
  r[1] = addc_cc(r[1], 0);
  r[2] = addc(r[2], 0);
  const uint32_t cw7 = __shfl_sync(COOP_FULL, cw, 7, 8);
  return ((sum >> 8) & 1u) + cw7;
}

// E, O += a[0..2] * b[0..2] (columns 0 .. 5, carries into column 6).  Two accumulators as in fq.cuh: the
// partial products a_i b_j with i + j even land on the even-aligned column pairs of E, those with i + j odd
// on the odd-aligned pairs of O, so every lo / hi pair is one IMAD.WIDE on an aligned register pair and no
// partial product is ever re-aligned (with one accumulator half of the instructions ptxas emits are moves).
// The value is sum_c (E[c] + O[c]) 2^(32 c); O[0] is unused.
G753_D void coop_prod_acc(uint32_t* E, uint32_t* O, const uint32_t* a, const uint32_t* b) {
  E[0] = mad_lo_cc(a[0], b[0], E[0]);
  E[1] = madc_hi_cc(a[0], b[0], E[1]);
  E[2] = madc_lo_cc(a[0], b[2], E[2]);
  E[3] = madc_hi_cc(a[0], b[2], E[3]);
  E[4] = madc_lo_cc(a[2], b[2], E[4]);
  E[5] = madc_hi_cc(a[2], b[2], E[5]);
  E[6] = addc(E[6], 0);
  E[2] = mad_lo_cc(a[1], b[1], E[2]);
  E[3] = madc_hi_cc(a[1], b[1], E[3]);
  E[4] = addc_cc(E[4], 0);
  E[5] = addc_cc(E[5], 0);
  E[6] = addc(E[6], 0);
  E[2] = mad_lo_cc(a[2], b[0], E[2]);
  E[3] = madc_hi_cc(a[2], b[0], E[3]);
  E[4] = addc_cc(E[4], 0);
  E[5] = addc_cc(E[5], 0);
  E[6] = addc(E[6], 0);
  O[1] = mad_lo_cc(a[0], b[1], O[1]);
  O[2] = madc_hi_cc(a[0], b[1], O[2]);
  O[3] = madc_lo_cc(a[1], b[2], O[3]);
  O[4] = madc_hi_cc(a[1], b[2], O[4]);
  O[5] = addc_cc(O[5], 0);
  O[6] = addc(O[6], 0);
  O[1] = mad_lo_cc(a[1], b[0], O[1]);
  O[2] = madc_hi_cc(a[1], b[0], O[2]);
  O[3] = madc_lo_cc(a[2], b[1], O[3]);
  O[4] = madc_hi_cc(a[2], b[1], O[4]);
  O[5] = addc_cc(O[5], 0);
  O[6] = addc(O[6], 0);
}

// out = a b / 2^768 mod p (lazily reduced: below (1 + x y / 2^15) p for inputs below x p, y p).
// n: this lane's limbs of p; np: -p^-1 mod 2^96.  All 32 lanes call it (4 products per warp).
// The warp runs alone on its scheduler: what counts is the number of instructions.
G753_D void coop_mul(uint32_t* out, const uint32_t* a, const uint32_t* b, const uint32_t* n, const uint32_t* np) {
  const unsigned l = coop_lane();
  uint32_t E[7], O[7];
#pragma unroll
  for (int i = 0; i < 7; i++) E[i] = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
#pragma unroll
    for (int i = 0; i < 7; i++) O[i] = 0;
    uint32_t B[3];
    B[0] = __shfl_sync(COOP_FULL, b[0], j, 8);
    B[1] = __shfl_sync(COOP_FULL, b[1], j, 8);
    B[2] = __shfl_sync(COOP_FULL, b[2], j, 8);
    coop_prod_acc(E, O, a, B);
    // the low digit t[0..2] of E + O, then M = t * np mod 2^96 (lane 0's is the one that counts)
    const uint32_t t0 = E[0];
    const uint32_t t1 = add_cc(E[1], O[1]);
    const uint32_t t2 = addc(E[2], O[2]);
    uint32_t M[3];
    M[0] = t0 * np[0];
    const uint64_t s1 = (uint64_t)__umulhi(t0, np[0]) + (uint64_t)(t0 * np[1]) + (uint64_t)(t1 * np[0]);
    M[1] = (uint32_t)s1;
    M[2] = (uint32_t)(s1 >> 32) + __umulhi(t0, np[1]) + __umulhi(t1, np[0]) + t0 * np[2] + t1 * np[1] + t2 * np[0];
    M[0] = __shfl_sync(COOP_FULL, M[0], 0, 8);
    M[1] = __shfl_sync(COOP_FULL, M[1], 0, 8);
    M[2] = __shfl_sync(COOP_FULL, M[2], 0, 8);
    coop_prod_acc(E, O, n, M);
    // merge the accumulators: columns 0 .. 6 of this lane's total
    E[1] = add_cc(E[1], O[1]);
#pragma unroll
    for (int i = 2; i < 6; i++) E[i] = addc_cc(E[i], O[i]);
    E[6] = addc(E[6], O[6]);
    // divide by 2^96: lane l keeps its high half, receives the low half of lane l + 1
    uint32_t r0 = __shfl_down_sync(COOP_FULL, E[0], 1, 8);
    uint32_t r1 = __shfl_down_sync(COOP_FULL, E[1], 1, 8);
    uint32_t r2 = __shfl_down_sync(COOP_FULL, E[2], 1, 8);
    if (l == 7) r0 = r1 = r2 = 0;
    E[0] = add_cc(E[3], r0);
    E[1] = addc_cc(E[4], r1);
    E[2] = addc_cc(E[5], r2);
    E[3] = addc(E[6], 0);
    E[4] = E[5] = E[6] = 0;
  }
  out[0] = E[0];
  out[1] = E[1];
  out[2] = E[2];
  coop_resolve(out, E[3]);
}

// out = a + y + z + inc over the octet (three-operand lane sums, carries resolved); returns the carry
// out of 2^768
G753_D uint32_t coop_add3(uint32_t* out, const uint32_t* a, const uint32_t* y, const uint32_t* z, uint32_t inc) {
  out[0] = add_cc(a[0], y[0]);
  out[1] = addc_cc(a[1], y[1]);
  out[2] = addc_cc(a[2], y[2]);
  uint32_t cw = addc(0, 0);
  out[0] = add_cc(out[0], inc);
  out[1] = addc_cc(out[1], 0);
  out[2] = addc_cc(out[2], 0);
  cw = addc(cw, 0);
  out[0] = add_cc(out[0], z[0]);
  out[1] = addc_cc(out[1], z[1]);
  out[2] = addc_cc(out[2], z[2]);
  cw = addc(cw, 0);
  return coop_resolve(out, cw);
}

// One warp's view of its slots: slot s, lane l holds limbs 3 l .. 3 l + 2 at cs[24 s + 3 l].
template <int FID>
struct CoopWarp {
  uint32_t* cs;          // this warp's slots (shared memory)
  const uint32_t* kp;    // [8][24] multiples of p (shared memory copy)
  uint32_t n[3], np[3];  // this lane's limbs of p, -p^-1 mod 2^96

  // shared memory of a block: COOP_CONST_WORDS words of constants, then warps x slots x 24 words
  G753_D void init(uint32_t* smem, unsigned slots_per_warp) {
    for (unsigned i = threadIdx.x; i < (unsigned)COOP_CONST_WORDS; i += blockDim.x) smem[i] = CoopField<FID>::kp()[i];
    __syncthreads();
    kp = smem;
    cs = smem + COOP_CONST_WORDS + (threadIdx.x >> 5) * slots_per_warp * 24;
    const unsigned l = coop_lane();
#pragma unroll
    for (int i = 0; i < 3; i++) {
      n[i] = kp[3 * l + i];
      np[i] = CoopField<FID>::np()[i];
    }
  }
  G753_D void ld(uint32_t* v, unsigned slot) const {
    const uint32_t* p = cs + slot * 24 + 3 * coop_lane();
    v[0] = p[0];
    v[1] = p[1];
    v[2] = p[2];
  }
  G753_D void st(unsigned slot, const uint32_t* v) {
    uint32_t* p = cs + slot * 24 + 3 * coop_lane();
    p[0] = v[0];
    p[1] = v[1];
    p[2] = v[2];
  }
  // `count` consecutive field elements global <-> slots (whole warp; canonical 24-limb elements)
  G753_D void load(unsigned slot, const Fq* g, unsigned count) {
    const uint32_t* src = (const uint32_t*)g;
    for (unsigned i = threadIdx.x & 31u; i < count * 24; i += 32) cs[slot * 24 + i] = src[i];
    __syncwarp();
  }
  G753_D void store(Fq* g, unsigned slot, unsigned count) {
    __syncwarp();
    uint32_t* dst = (uint32_t*)g;
    for (unsigned i = threadIdx.x & 31u; i < count * 24; i += 32) dst[i] = cs[slot * 24 + i];
  }
  G753_D void set_zero(unsigned slot, unsigned count) {
    for (unsigned i = threadIdx.x & 31u; i < count * 24; i += 32) cs[slot * 24 + i] = 0;
    __syncwarp();
  }
  G753_D void set_one(unsigned slot) {
    for (unsigned i = threadIdx.x & 31u; i < 24; i += 32) cs[slot * 24 + i] = G753_FC(FID).one[i];
    __syncwarp();
  }
  G753_D void copy(unsigned dst, unsigned src, unsigned count) {
    __syncwarp();
    for (unsigned i = threadIdx.x & 31u; i < count * 24; i += 32) cs[dst * 24 + i] = cs[src * 24 + i];
    __syncwarp();
  }

  // acc (3 limbs + carry word) += c * x
  static G753_D void mac3(uint32_t* acc, uint32_t& cw, const uint32_t* x, uint32_t c) {
    acc[0] = mad_lo_cc(x[0], c, acc[0]);
    acc[1] = madc_lo_cc(x[1], c, acc[1]);
    acc[2] = madc_lo_cc(x[2], c, acc[2]);
    cw = addc(cw, 0);
    acc[1] = mad_hi_cc(x[0], c, acc[1]);
    acc[2] = madc_hi_cc(x[1], c, acc[2]);
    cw = madc_hi(x[2], c, cw);
  }

  // interpret a micro-program (rows of 4 instructions x 2 words, __constant__ memory); uniform control flow:
  // in a row of products every octet multiplies, in a linear row every octet evaluates
  // ca a + cb b + cc c (+ 2^k p), idle octets on harmless operands.
  template <class PROG>
  G753_D void run() {
    const unsigned o = coop_octet(), l = coop_lane();
    uint32_t nf = PROG::w(0), n0 = PROG::w(2 * o), n1 = PROG::w(2 * o + 1);
#pragma unroll 1
    for (unsigned r = 0; r < PROG::ROWS; r++) {
      const uint32_t flags = nf, w0 = n0, w1 = n1;
      if (r + 1 < PROG::ROWS) {        // the next row's words arrive while this row executes
        nf = PROG::w(8 * (r + 1));
        n0 = PROG::w(8 * (r + 1) + 2 * o);
        n1 = PROG::w(8 * (r + 1) + 2 * o + 1);
      }
      const unsigned op = w0 & 7u, d = (w0 >> 3) & 127u, a = (w0 >> 10) & 127u, b = (w0 >> 17) & 127u, k = (w0 >> 24) & 15u;
      uint32_t A[3], B[3], R[3];
      ld(A, a);
      ld(B, b);
      if (flags & COOP_HAS_MUL) {
        coop_mul(R, A, B, n, np);
      } else {
        uint32_t C[3];
        ld(C, w1 & 127u);
        const uint32_t ca = (w1 >> 7) & 15u, cb = (w1 >> 12) & 15u, cc = (w1 >> 17) & 15u;
        const bool na = (w1 >> 11) & 1u, nb = (w1 >> 16) & 1u, nc = (w1 >> 21) & 1u;
        const bool neg = na || nb || nc;
        const uint32_t* kpl = kp + k * 24 + 3 * l;
        uint32_t cw = 0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
          R[i] = neg ? kpl[i] : 0u;
          A[i] = na ? ~A[i] : A[i];
          B[i] = nb ? ~B[i] : B[i];
          C[i] = nc ? ~C[i] : C[i];
        }
        mac3(R, cw, A, ca);
        mac3(R, cw, B, cb);
        mac3(R, cw, C, cc);
        // the + 1 of every complemented term enters at the lowest lane
        const uint32_t inc = l == 0 ? (na ? ca : 0u) + (nb ? cb : 0u) + (nc ? cc : 0u) : 0u;
        R[0] = add_cc(R[0], inc);
        R[1] = addc_cc(R[1], 0);
        R[2] = addc_cc(R[2], 0);
        cw = addc(cw, 0);
        coop_resolve(R, cw);
      }
      __syncwarp();
      if (op != COOP_NOP) st(d, R);
      __syncwarp();
    }
  }

  // slot value (below 2 p) is 0 mod p: 0 or p
  G753_D bool is_zero(unsigned slot) const {
    uint32_t v[3];
    ld(v, slot);
    const bool z = (v[0] | v[1] | v[2]) == 0;
    const bool e = v[0] == n[0] && v[1] == n[1] && v[2] == n[2];
    const unsigned sh = threadIdx.x & 24u;
    const unsigned Z = (__ballot_sync(COOP_FULL, z) >> sh) & 0xffu, E = (__ballot_sync(COOP_FULL, e) >> sh) & 0xffu;
    return Z == 0xffu || E == 0xffu;
  }
  G753_D bool is_zero(unsigned slot, unsigned count) const {
    bool z = true;
    for (unsigned i = 0; i < count; i++) z = is_zero(slot + i) && z;
    return z;
  }
  // slots [slot, slot + count): values below 2 p -> canonical
  G753_D void canonicalize(unsigned slot, unsigned count) {
    const unsigned l = coop_lane();
    for (unsigned i = 0; i < count; i++) {
      uint32_t v[3], y[3], z[3], d[3];
      ld(v, slot + i);
#pragma unroll
      for (int q = 0; q < 3; q++) {
        y[q] = ~n[q];
        z[q] = 0;
      }
      const uint32_t top = coop_add3(d, v, y, z, l == 0 ? 1u : 0u);   // v - p; carry out <=> v >= p
      __syncwarp();
      if (top && coop_octet() == 0) st(slot + i, d);
      __syncwarp();
    }
  }
};

// per-group programs and slot map (tools/gen_coop.py: Layout)
template <int GID> struct CoopGroup;
#define G753_COOP_GROUP(GID, NAME, KK, FIELD)                                                                     \
  template <> struct CoopGroup<GID> {                                                                             \
    static constexpr int K = KK, FID = FIELD, P = 0, Q = 4 * KK, ONE = 8 * KK, SLOTS = COOP_##NAME##_SLOTS;       \
    typedef COOP_##NAME##_DBL_P Dbl;                                                                               \
    typedef COOP_##NAME##_ADD_HEAD_P AddHead;                                                                      \
    typedef COOP_##NAME##_ADD_TAIL_P AddTail;                                                                      \
    typedef COOP_##NAME##_REDUCE_P Reduce;                                                                         \
    typedef COOP_##NAME##_TO_PROJ_P ToProj;                                                                        \
    typedef COOP_##NAME##_FROM_PROJ_P FromProj;                                                                    \
    static G753_D const unsigned char* add_tp() { return COOP_##NAME##_ADD_TP; }                                   \
    static G753_D const unsigned char* add_tr() { return COOP_##NAME##_ADD_TR; }                                   \
  };
G753_COOP_GROUP(0, M4G1, 1, 0)
G753_COOP_GROUP(1, M4G2, 2, 0)
G753_COOP_GROUP(2, M6G1, 1, 1)
G753_COOP_GROUP(3, M6G2, 3, 1)
#undef G753_COOP_GROUP

// The group law of one warp on its accumulator P (XYZZ, lazily reduced) and operand Q (XYZZ, canonical).
template <int GID>
struct CoopEc {
  typedef CoopGroup<GID> Gp;
  static constexpr int K = Gp::K;
  CoopWarp<Gp::FID> w;

  G753_D void init(uint32_t* smem) {
    w.init(smem, Gp::SLOTS);
    w.set_one(Gp::ONE);
  }
  G753_D void set_inf() { w.set_zero(Gp::P, 4 * K); }
  G753_D bool is_inf() const { return w.is_zero(Gp::P + 2 * K, K); }
  G753_D void dbl() { w.template run<typename Gp::Dbl>(); }
  // P += Q, Q an XYZZ point in the Q slots (zero ZZ = infinity)
  G753_D void add_q() {
    if (w.is_zero(Gp::Q + 2 * K, K)) return;
    if (is_inf()) {
      w.copy(Gp::P, Gp::Q, 4 * K);
      return;
    }
    w.template run<typename Gp::AddHead>();
    bool pz = true, rz = true;
    for (int i = 0; i < K; i++) {
      pz = w.is_zero(Gp::add_tp()[i]) && pz;
      rz = w.is_zero(Gp::add_tr()[i]) && rz;
    }
    if (pz) {
      if (rz) dbl();          // P == Q
      else set_inf();         // P == -Q
      return;
    }
    w.template run<typename Gp::AddTail>();
  }
  G753_D void add_g(const Fq* q_xyzz) {
    w.load(Gp::Q, q_xyzz, 4 * K);
    add_q();
  }
  // homogeneous (X : Y : Z), canonical, in global memory -> added to P
  G753_D void add_projective_g(const Fq* q_xyz) {
    w.load(Gp::Q, q_xyz, 3 * K);
    if (w.is_zero(Gp::Q + 2 * K, K)) return;    // Z == 0: the point at infinity
    w.template run<typename Gp::FromProj>();
    add_q();
  }
  // P -> the reference's homogeneous projective, canonical, (0 : 1 : 0) for infinity; written to out_xyz
  G753_D void store_projective(Fq* out_xyz) {
    if (is_inf()) {
      w.set_zero(Gp::P, 3 * K);
      w.copy(Gp::P + K, Gp::ONE, 1);
    } else {
      w.template run<typename Gp::ToProj>();
      w.canonicalize(Gp::P, 3 * K);
    }
    w.store(out_xyz, Gp::P, 3 * K);
  }
  // P (lazily reduced XYZZ) -> canonical XYZZ in global memory
  G753_D void store_xyzz(Fq* out) {
    if (is_inf()) {
      w.set_zero(Gp::P, 4 * K);    // all-zero limbs, the encoding of infinity the slot kernels use
    } else {
      w.template run<typename Gp::Reduce>();   // X, Y below 2 p (ZZ, ZZZ are); then one conditional subtraction each
      w.canonicalize(Gp::P, 4 * K);
    }
    w.store(out, Gp::P, 4 * K);
  }
};

template <int GID>
constexpr size_t coop_smem_bytes(unsigned warps) {
  return sizeof(uint32_t) * ((size_t)COOP_CONST_WORDS + (size_t)warps * CoopGroup<GID>::SLOTS * 24);
}

}  // namespace g753
#endif
