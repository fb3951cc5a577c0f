// CPU restatement (C++17) of the reference's ALGORITHMS for the Groth16 prover hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (ginger-lib_b200/) includes, links or
// executes this file; it is used by tests/ as the checker at sizes the Python oracle cannot
// reach, and by bench.py's cpu_baseline / --impl reference legs as the timed "port" of the
// reference's rayon CPU path (the Rust reference itself cannot be built here: no rustc/cargo,
// SURVEY.md F3).  Parity pinning: tests/test_oracle_cpp.py checks every function below against
// oracle/g753.py, which in turn replays the reference's own known-answer tests
// (tests/golden/reference_kat.json), and replays the raw-limb field KATs here directly.
//
// What is restated, with the reference lines it follows (relative to /root/reference/algebra/src):
//   biginteger/mod.rs:108-141        adc / sbb / mac_with_carry (u64 limbs, u128 accumulate)
//   fields/models/fp_768.rs:1009-1185  mul_assign: 12x12 schoolbook product, then
//   fields/models/fp_768.rs:50-281     mont_reduce: 12 rounds k = r_i * INV, r += k * p
//   fields/models/fp_768.rs:929-949, 870-883, 303-309  add / sub / neg / double
//   fields/models/fp_768.rs:551-605    inverse: binary extended Euclid on the Montgomery form
//   fields/models/fp2.rs:128-144, 387-401   Fp2 square / Karatsuba mul (non-residue 13)
//   fields/models/fp3.rs:165-185, 451-478   Fp3 CH-SQR2 / Karatsuba mul (non-residue 11)
//   curves/models/short_weierstrass_projective.rs:444-479  double_in_place (dbl-2007-bl)
//   curves/models/short_weierstrass_projective.rs:481-519  add_assign_mixed (madd-1998-cmo)
//   curves/models/short_weierstrass_projective.rs:574-617  add_assign (add-1998-cmo-2), :298-317 eq
//   curves/models/short_weierstrass_projective.rs:402-442  batch_normalization, :663-678 into_affine
//   msm/variable_base.rs:10-83   msm_inner: unsigned c-bit windows, c from scalars.len(),
//                                one task per window (rayon into_par_iter -> thread pool here)
//   fft/domain.rs:305-416        best_fft / serial_fft / parallel_fft, :113-179 the four transforms
// Constants (R, R2, INV) are derived at start-up from the moduli alone and checked against the
// reference's literals by the tests.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

typedef unsigned __int128 u128;
typedef uint64_t u64;
static const int N = 12;

// ------------------------------------------------------------------------------------------
// big integers (biginteger/mod.rs:108-141, biginteger/macros.rs)
// ------------------------------------------------------------------------------------------
struct Big {
  u64 l[N];
};
static inline u64 adc(u64 a, u64 b, u64& carry) {
  u128 t = (u128)a + b + carry;
  carry = (u64)(t >> 64);
  return (u64)t;
}
static inline u64 sbb(u64 a, u64 b, u64& borrow) {
  u128 t = ((u128)1 << 64) + a - b - borrow;
  borrow = (t >> 64) == 0 ? 1 : 0;
  return (u64)t;
}
static inline u64 mac_with_carry(u64 a, u64 b, u64 c, u64& carry) {
  u128 t = (u128)a + (u128)b * c + carry;
  carry = (u64)(t >> 64);
  return (u64)t;
}
static inline bool big_lt(const Big& a, const Big& b) {
  for (int i = N - 1; i >= 0; i--) {
    if (a.l[i] < b.l[i]) return true;
    if (a.l[i] > b.l[i]) return false;
  }
  return false;
}
static inline bool big_eq(const Big& a, const Big& b) { return memcmp(a.l, b.l, sizeof(a.l)) == 0; }
static inline bool big_is_zero(const Big& a) {
  u64 t = 0;
  for (int i = 0; i < N; i++) t |= a.l[i];
  return t == 0;
}
static inline void big_add_nocarry(Big& a, const Big& b) {
  u64 c = 0;
  for (int i = 0; i < N; i++) a.l[i] = adc(a.l[i], b.l[i], c);
}
static inline void big_sub_noborrow(Big& a, const Big& b) {
  u64 br = 0;
  for (int i = 0; i < N; i++) a.l[i] = sbb(a.l[i], b.l[i], br);
}
static inline void big_div2(Big& a) {
  u64 t = 0;
  for (int i = N - 1; i >= 0; i--) {
    u64 t2 = a.l[i] << 63;
    a.l[i] = (a.l[i] >> 1) | t;
    t = t2;
  }
}
static inline void big_mul2(Big& a) {
  u64 last = 0;
  for (int i = 0; i < N; i++) {
    u64 t = a.l[i] >> 63;
    a.l[i] = (a.l[i] << 1) | last;
    last = t;
  }
}
static inline bool big_is_even(const Big& a) { return (a.l[0] & 1) == 0; }
// (x >> shift) mod 2^c   - divn + as_ref()[0] % (1 << c) of variable_base.rs:47-51
static inline u64 big_window(const Big& a, unsigned shift, unsigned c) {
  unsigned limb = shift >> 6, sh = shift & 63;
  u64 v = limb < (unsigned)N ? a.l[limb] >> sh : 0;
  if (sh && limb + 1 < (unsigned)N) v |= a.l[limb + 1] << (64 - sh);
  return v & (((u64)1 << c) - 1);
}

// ------------------------------------------------------------------------------------------
// field parameters: field 0 = mnt4753::Fq (= mnt6753::Fr), field 1 = mnt6753::Fq (= mnt4753::Fr)
// ------------------------------------------------------------------------------------------
struct FieldParams {
  Big p, r, r2, gen, gen_inv_placeholder, root;
  u64 inv;
  unsigned two_adicity;
};
static FieldParams FP[2];

static const u64 MODULUS[2][N] = {
    // fields/mnt4753/fq.rs:18-31
    {0x5E9063DE245E8001ull, 0xE39D54522CDD119Full, 0x638810719AC425F0ull, 0x685ACCE9767254A4ull,
     0xB80F0DA5CB537E38ull, 0xB117E776F218059Dull, 0x99D124D9A15AF79Dull, 0x07FDB925E8A0ED8Dull,
     0x5EB7E8F96C97D873ull, 0xB7F997505B8FAFEDull, 0x10229022EEE2CDADull, 0x01C4C62D92C411ull},
    // fields/mnt6753/fq.rs:17-30
    {0xD90776E240000001ull, 0x4EA099170FA13A4Full, 0xD6C381BC3F005797ull, 0xB9DFF97634993AA4ull,
     0x3EEBCA9429212636ull, 0xB26C5C28C859A99Bull, 0x99D124D9A15AF79Dull, 0x07FDB925E8A0ED8Dull,
     0x5EB7E8F96C97D873ull, 0xB7F997505B8FAFEDull, 0x10229022EEE2CDADull, 0x01C4C62D92C411ull}};
static const unsigned TWO_ADICITY[2] = {15, 30};

template <int F>
struct Fq {
  Big v;  // Montgomery representation, always < p
};

template <int F>
static inline void fq_reduce(Big& a) {  // fp_768.rs:44-48
  if (!big_lt(a, FP[F].p)) big_sub_noborrow(a, FP[F].p);
}
template <int F>
static inline Fq<F> fq_add(const Fq<F>& a, const Fq<F>& b) {  // :929-937
  Fq<F> r = a;
  big_add_nocarry(r.v, b.v);
  fq_reduce<F>(r.v);
  return r;
}
template <int F>
static inline Fq<F> fq_sub(const Fq<F>& a, const Fq<F>& b) {  // :939-949
  Fq<F> r = a;
  if (big_lt(r.v, b.v)) big_add_nocarry(r.v, FP[F].p);
  big_sub_noborrow(r.v, b.v);
  return r;
}
template <int F>
static inline Fq<F> fq_double(const Fq<F>& a) {  // :303-309
  Fq<F> r = a;
  big_mul2(r.v);
  fq_reduce<F>(r.v);
  return r;
}
template <int F>
static inline Fq<F> fq_neg(const Fq<F>& a) {  // :870-883
  if (big_is_zero(a.v)) return a;
  Fq<F> r;
  r.v = FP[F].p;
  big_sub_noborrow(r.v, a.v);
  return r;
}
// mont_reduce (fp_768.rs:50-281): 12 rounds over a 24-limb value, then one conditional subtract
template <int F>
static inline Fq<F> mont_reduce(u64* r) {
  const u64* p = FP[F].p.l;
  u64 carry2 = 0;
  for (int i = 0; i < N; i++) {
    u64 k = r[i] * FP[F].inv;
    u64 carry = 0;
    mac_with_carry(r[i], k, p[0], carry);
    for (int j = 1; j < N; j++) r[i + j] = mac_with_carry(r[i + j], k, p[j], carry);
    r[i + N] = adc(r[i + N], carry2, carry);
    carry2 = carry;
  }
  Fq<F> out;
  for (int i = 0; i < N; i++) out.v.l[i] = r[i + N];
  fq_reduce<F>(out.v);
  return out;
}
template <int F>
static inline Fq<F> fq_mul(const Fq<F>& a, const Fq<F>& b) {  // :1009-1185
  u64 r[2 * N];
  u64 carry = 0;
  for (int j = 0; j < N; j++) r[j] = mac_with_carry(0, a.v.l[0], b.v.l[j], carry);
  r[N] = carry;
  for (int i = 1; i < N; i++) {
    carry = 0;
    for (int j = 0; j < N; j++) r[i + j] = mac_with_carry(r[i + j], a.v.l[i], b.v.l[j], carry);
    r[i + N] = carry;
  }
  return mont_reduce<F>(r);
}
template <int F>
static inline Fq<F> fq_square(const Fq<F>& a) {  // :339-548 (same value as a * a)
  return fq_mul<F>(a, a);
}
template <int F>
static inline Fq<F> fq_zero() {
  Fq<F> r;
  memset(&r, 0, sizeof(r));
  return r;
}
template <int F>
static inline Fq<F> fq_one() {
  Fq<F> r;
  r.v = FP[F].r;
  return r;
}
template <int F>
static inline bool fq_is_zero(const Fq<F>& a) {
  return big_is_zero(a.v);
}
template <int F>
static inline bool fq_eq(const Fq<F>& a, const Fq<F>& b) {
  return big_eq(a.v, b.v);
}
template <int F>
static Fq<F> fq_from_repr(const Big& x) {  // :627-635
  Fq<F> r, r2;
  r.v = x;
  r2.v = FP[F].r2;
  if (!big_lt(x, FP[F].p)) return fq_zero<F>();
  return fq_mul<F>(r, r2);
}
template <int F>
static Big fq_into_repr(const Fq<F>& a) {  // :637-667
  u64 r[2 * N];
  memset(r, 0, sizeof(r));
  for (int i = 0; i < N; i++) r[i] = a.v.l[i];
  return mont_reduce<F>(r).v;
}
template <int F>
static Fq<F> fq_inverse(const Fq<F>& a) {  // :551-605 (binary extended Euclid), a != 0
  Big one;
  memset(&one, 0, sizeof(one));
  one.l[0] = 1;
  Big u = a.v, v = FP[F].p;
  Fq<F> b, c = fq_zero<F>();
  b.v = FP[F].r2;
  while (!big_eq(u, one) && !big_eq(v, one)) {
    while (big_is_even(u)) {
      big_div2(u);
      if (big_is_even(b.v)) big_div2(b.v);
      else {
        // b + p < 2^768 (p has 753 bits): no carry out
        big_add_nocarry(b.v, FP[F].p);
        big_div2(b.v);
      }
    }
    while (big_is_even(v)) {
      big_div2(v);
      if (big_is_even(c.v)) big_div2(c.v);
      else {
        big_add_nocarry(c.v, FP[F].p);
        big_div2(c.v);
      }
    }
    if (big_lt(v, u)) {
      big_sub_noborrow(u, v);
      b = fq_sub<F>(b, c);
    } else {
      big_sub_noborrow(v, u);
      c = fq_sub<F>(c, b);
    }
  }
  return big_eq(u, one) ? b : c;
}
template <int F>
static Fq<F> fq_pow(const Fq<F>& a, u64 e) {
  Fq<F> r = fq_one<F>();
  for (int i = 63; i >= 0; i--) {
    r = fq_square<F>(r);
    if ((e >> i) & 1) r = fq_mul<F>(r, a);
  }
  return r;
}

static void init_field(int f) {
  FieldParams& P = FP[f];
  memcpy(P.p.l, MODULUS[f], sizeof(P.p.l));
  P.two_adicity = TWO_ADICITY[f];
  // R = 2^768 mod p, R2 = 2^1536 mod p by repeated modular doubling of 1
  Big x;
  memset(&x, 0, sizeof(x));
  x.l[0] = 1;
  for (int i = 0; i < 1536; i++) {
    big_mul2(x);  // x < p < 2^753 so 2x < 2^768: no overflow
    if (!big_lt(x, P.p)) big_sub_noborrow(x, P.p);
    if (i == 767) P.r = x;
  }
  P.r2 = x;
  // INV = -p^-1 mod 2^64 (Newton)
  u64 inv = 1;
  for (int i = 0; i < 6; i++) inv *= 2 - P.p.l[0] * inv;
  P.inv = (u64)0 - inv;
}
static struct Init {
  Init() {
    init_field(0);
    init_field(1);
  }
} g_init;

template <int F>
static Fq<F> fq_from_u64(u64 x) {
  Big b;
  memset(&b, 0, sizeof(b));
  b.l[0] = x;
  return fq_from_repr<F>(b);
}
// 2^s-th root of unity: 17^T, T = (p - 1) >> s   (FpParameters::ROOT_OF_UNITY)
template <int F>
static Fq<F> fq_root_of_unity() {
  Big t = FP[F].p;
  t.l[0] -= 1;
  for (unsigned i = 0; i < FP[F].two_adicity; i++) big_div2(t);
  Fq<F> g = fq_from_u64<F>(17), r = fq_one<F>();
  for (int i = 64 * N - 1; i >= 0; i--) {
    r = fq_square<F>(r);
    if ((t.l[i >> 6] >> (i & 63)) & 1) r = fq_mul<F>(r, g);
  }
  return r;
}

// ------------------------------------------------------------------------------------------
// base fields of the four groups: Fp (k=1), Fp2 (non-residue 13), Fp3 (non-residue 11)
// ------------------------------------------------------------------------------------------
template <int F>
struct K1 {
  static const int K = 1;
  Fq<F> c0;
  static K1 zero() { return K1{fq_zero<F>()}; }
  static K1 one() { return K1{fq_one<F>()}; }
  bool is_zero() const { return fq_is_zero(c0); }
  bool is_one() const { return fq_eq(c0, fq_one<F>()); }
  bool operator==(const K1& o) const { return fq_eq(c0, o.c0); }
  K1 operator+(const K1& o) const { return K1{fq_add(c0, o.c0)}; }
  K1 operator-(const K1& o) const { return K1{fq_sub(c0, o.c0)}; }
  K1 operator*(const K1& o) const { return K1{fq_mul(c0, o.c0)}; }
  K1 square() const { return K1{fq_square(c0)}; }
  K1 dbl() const { return K1{fq_double(c0)}; }
  K1 neg() const { return K1{fq_neg(c0)}; }
  K1 inverse() const { return K1{fq_inverse(c0)}; }
};
template <int F, int NR>
struct K2 {
  static const int K = 2;
  Fq<F> c0, c1;
  static Fq<F> nr() {
    static const Fq<F> v = fq_from_u64<F>(NR);
    return v;
  }
  static K2 zero() { return K2{fq_zero<F>(), fq_zero<F>()}; }
  static K2 one() { return K2{fq_one<F>(), fq_zero<F>()}; }
  bool is_zero() const { return fq_is_zero(c0) && fq_is_zero(c1); }
  bool is_one() const { return fq_eq(c0, fq_one<F>()) && fq_is_zero(c1); }
  bool operator==(const K2& o) const { return fq_eq(c0, o.c0) && fq_eq(c1, o.c1); }
  K2 operator+(const K2& o) const { return K2{fq_add(c0, o.c0), fq_add(c1, o.c1)}; }
  K2 operator-(const K2& o) const { return K2{fq_sub(c0, o.c0), fq_sub(c1, o.c1)}; }
  K2 operator*(const K2& o) const {  // fp2.rs:387-401
    Fq<F> v0 = fq_mul(c0, o.c0), v1 = fq_mul(c1, o.c1);
    Fq<F> t = fq_mul(fq_add(c1, c0), fq_add(o.c0, o.c1));
    t = fq_sub(fq_sub(t, v0), v1);
    return K2{fq_add(v0, fq_mul(nr(), v1)), t};
  }
  K2 square() const {  // fp2.rs:128-144
    Fq<F> v0 = fq_sub(c0, c1);
    Fq<F> v3 = fq_sub(c0, fq_mul(nr(), c1));
    Fq<F> v2 = fq_mul(c0, c1);
    v0 = fq_add(fq_mul(v0, v3), v2);
    return K2{fq_add(v0, fq_mul(nr(), v2)), fq_double(v2)};
  }
  K2 dbl() const { return K2{fq_double(c0), fq_double(c1)}; }
  K2 neg() const { return K2{fq_neg(c0), fq_neg(c1)}; }
  K2 inverse() const {  // Guide to PBC alg 5.19: v0 = c0^2 - nr c1^2
    Fq<F> v0 = fq_sub(fq_square(c0), fq_mul(nr(), fq_square(c1)));
    Fq<F> vi = fq_inverse(v0);
    return K2{fq_mul(c0, vi), fq_neg(fq_mul(c1, vi))};
  }
};
template <int F, int NR>
struct K3 {
  static const int K = 3;
  Fq<F> c0, c1, c2;
  static Fq<F> nr() {
    static const Fq<F> v = fq_from_u64<F>(NR);
    return v;
  }
  static K3 zero() { return K3{fq_zero<F>(), fq_zero<F>(), fq_zero<F>()}; }
  static K3 one() { return K3{fq_one<F>(), fq_zero<F>(), fq_zero<F>()}; }
  bool is_zero() const { return fq_is_zero(c0) && fq_is_zero(c1) && fq_is_zero(c2); }
  bool is_one() const { return fq_eq(c0, fq_one<F>()) && fq_is_zero(c1) && fq_is_zero(c2); }
  bool operator==(const K3& o) const { return fq_eq(c0, o.c0) && fq_eq(c1, o.c1) && fq_eq(c2, o.c2); }
  K3 operator+(const K3& o) const { return K3{fq_add(c0, o.c0), fq_add(c1, o.c1), fq_add(c2, o.c2)}; }
  K3 operator-(const K3& o) const { return K3{fq_sub(c0, o.c0), fq_sub(c1, o.c1), fq_sub(c2, o.c2)}; }
  K3 operator*(const K3& o) const {  // fp3.rs:451-478
    const Fq<F>&a = o.c0, &b = o.c1, &c = o.c2, &d = c0, &e = c1, &f = c2;
    Fq<F> ad = fq_mul(d, a), be = fq_mul(e, b), cf = fq_mul(f, c);
    Fq<F> x = fq_sub(fq_sub(fq_mul(fq_add(e, f), fq_add(b, c)), be), cf);
    Fq<F> y = fq_sub(fq_sub(fq_mul(fq_add(d, e), fq_add(a, b)), ad), be);
    Fq<F> z = fq_sub(fq_add(fq_sub(fq_mul(fq_add(d, f), fq_add(a, c)), ad), be), cf);
    return K3{fq_add(ad, fq_mul(nr(), x)), fq_add(y, fq_mul(nr(), cf)), z};
  }
  K3 square() const {  // fp3.rs:165-185 (CH-SQR2)
    Fq<F> s0 = fq_square(c0);
    Fq<F> ab = fq_mul(c0, c1);
    Fq<F> s1 = fq_add(ab, ab);
    Fq<F> s2 = fq_square(fq_add(fq_sub(c0, c1), c2));
    Fq<F> bc = fq_mul(c1, c2);
    Fq<F> s3 = fq_add(bc, bc);
    Fq<F> s4 = fq_square(c2);
    return K3{fq_add(s0, fq_mul(nr(), s3)), fq_add(s1, fq_mul(nr(), s4)),
              fq_sub(fq_sub(fq_add(fq_add(s1, s2), s3), s0), s4)};
  }
  K3 dbl() const { return K3{fq_double(c0), fq_double(c1), fq_double(c2)}; }
  K3 neg() const { return K3{fq_neg(c0), fq_neg(c1), fq_neg(c2)}; }
  K3 inverse() const {  // fp3.rs:186-215 (Beuchat et al. alg 17)
    Fq<F> t0 = fq_square(c0), t1 = fq_square(c1), t2 = fq_square(c2);
    Fq<F> t3 = fq_mul(c0, c1), t4 = fq_mul(c0, c2), t5 = fq_mul(c1, c2);
    Fq<F> n5 = fq_mul(nr(), t5);
    Fq<F> s0 = fq_sub(t0, n5);
    Fq<F> s1 = fq_sub(fq_mul(nr(), t2), t3);
    Fq<F> s2 = fq_sub(t1, t4);
    Fq<F> a1 = fq_mul(c2, s1), a2 = fq_mul(c1, s2);
    Fq<F> a3 = fq_mul(nr(), fq_add(a1, a2));
    Fq<F> t6 = fq_inverse(fq_add(fq_mul(c0, s0), a3));
    return K3{fq_mul(t6, s0), fq_mul(t6, s1), fq_mul(t6, s2)};
  }
};

// ------------------------------------------------------------------------------------------
// curves (SWModelParameters): group ids 0 = MNT4 G1, 1 = MNT4 G2, 2 = MNT6 G1, 3 = MNT6 G2
// ------------------------------------------------------------------------------------------
struct M4G1 {
  typedef K1<0> BF;
  static BF mul_by_a(const BF& x) { return x.dbl(); }  // a = 2: same value as COEFF_A * x
};
struct M6G1 {
  typedef K1<1> BF;
  static BF mul_by_a(const BF& x) {
    static const Fq<1> a = fq_from_u64<1>(11);
    return BF{fq_mul(a, x.c0)};
  }
};
struct M4G2 {  // curves/mnt4753/g2.rs:112-118
  typedef K2<0, 13> BF;
  static BF mul_by_a(const BF& x) {
    static const Fq<0> a = fq_from_u64<0>(26);
    return BF{fq_mul(a, x.c0), fq_mul(a, x.c1)};
  }
};
struct M6G2 {  // curves/mnt6753/g2.rs:148-155
  typedef K3<1, 11> BF;
  static BF mul_by_a(const BF& x) {
    static const Fq<1> a0 = fq_from_u64<1>(121), a2 = fq_from_u64<1>(11);
    return BF{fq_mul(a0, x.c1), fq_mul(a0, x.c2), fq_mul(a2, x.c0)};
  }
};

template <class C>
struct AffineP {
  typename C::BF x, y;
  bool infinity;
};
template <class C>
struct ProjP {
  typedef typename C::BF BF;
  BF x, y, z;
  static ProjP zero() { return ProjP{BF::zero(), BF::one(), BF::zero()}; }  // :378-385
  bool is_zero() const { return z.is_zero(); }
  bool is_normalized() const { return is_zero() || z.is_one(); }

  bool equals(const ProjP& o) const {  // :298-317
    if (is_zero()) return o.is_zero();
    if (o.is_zero()) return false;
    if (!((x * o.z) == (o.x * z))) return false;
    return (y * o.z) == (o.y * z);
  }
  void double_in_place() {  // :444-479 dbl-2007-bl
    if (is_zero()) return;
    BF xx = x.square();
    BF zz = z.square();
    BF w = C::mul_by_a(zz) + (xx + xx.dbl());
    BF s = (y * z).dbl();
    BF sss = s.square() * s;
    BF r = y * s;
    BF rr = r.square();
    BF b = (x + r).square() - xx - rr;
    BF h = w.square() - (b + b);
    x = h * s;
    y = w * (b - h) - (rr + rr);
    z = sss;
  }
  void add_assign_mixed(const AffineP<C>& o) {  // :481-519 madd-1998-cmo
    if (o.infinity) return;
    if (is_zero()) {
      x = o.x;
      y = o.y;
      z = BF::one();
      return;
    }
    BF v = o.x * z;
    BF u = o.y * z;
    if (u == y && v == x) {
      double_in_place();
      return;
    }
    u = u - y;
    BF uu = u.square();
    v = v - x;
    BF vv = v.square();
    BF vvv = v * vv;
    BF r = vv * x;
    BF a = uu * z - vvv - r.dbl();
    x = v * a;
    y = u * (r - a) - vvv * y;
    z = vvv * z;
  }
  void add_assign(const ProjP& o) {  // :574-617 add-1998-cmo-2
    if (is_zero()) {
      *this = o;
      return;
    }
    if (o.is_zero()) return;
    if (equals(o)) {
      double_in_place();
      return;
    }
    BF y1z2 = y * o.z;
    BF x1z2 = x * o.z;
    BF z1z2 = z * o.z;
    BF u = (z * o.y) - y1z2;
    BF uu = u.square();
    BF v = (z * o.x) - x1z2;
    BF vv = v.square();
    BF vvv = v * vv;
    BF r = vv * x1z2;
    BF a = (uu * z1z2) - (vvv + r + r);
    x = v * a;
    y = ((r - a) * u) - (vvv * y1z2);
    z = vvv * z1z2;
  }
  AffineP<C> into_affine() const {  // :663-678
    if (is_zero()) return AffineP<C>{BF::zero(), BF::one(), true};
    if (z.is_one()) return AffineP<C>{x, y, false};
    BF zi = z.inverse();
    return AffineP<C>{x * zi, y * zi, false};
  }
};

template <class C>
static void batch_normalization(std::vector<ProjP<C>>& v) {  // :402-442
  typedef typename C::BF BF;
  std::vector<BF> prod;
  prod.reserve(v.size());
  BF tmp = BF::one();
  for (auto& g : v)
    if (!g.is_normalized()) {
      tmp = tmp * g.z;
      prod.push_back(tmp);
    }
  if (prod.empty()) return;
  tmp = tmp.inverse();
  size_t k = prod.size();
  for (size_t i = v.size(); i-- > 0;) {
    ProjP<C>& g = v[i];
    if (g.is_normalized()) continue;
    k--;
    BF s = k > 0 ? prod[k - 1] : BF::one();
    BF newtmp = tmp * g.z;
    g.z = tmp * s;
    tmp = newtmp;
  }
  for (auto& g : v)
    if (!g.is_normalized()) {
      g.x = g.x * g.z;
      g.y = g.y * g.z;
      g.z = BF::one();
    }
}

// simple thread pool: run task(i) for i in [0, count) on nthreads threads
template <class Fn>
static void parallel_for(size_t count, unsigned nthreads, Fn fn) {
  if (nthreads <= 1 || count <= 1) {
    for (size_t i = 0; i < count; i++) fn(i);
    return;
  }
  std::atomic<size_t> next(0);
  std::vector<std::thread> th;
  unsigned t = (unsigned)(count < nthreads ? count : nthreads);
  for (unsigned k = 0; k < t; k++)
    th.emplace_back([&]() {
      for (;;) {
        size_t i = next.fetch_add(1);
        if (i >= count) break;
        fn(i);
      }
    });
  for (auto& x : th) x.join();
}

// ------------------------------------------------------------------------------------------
// VariableBaseMSM::msm_inner (msm/variable_base.rs:10-83)
// ------------------------------------------------------------------------------------------
static unsigned ln_window(size_t n_scalars) {  // :14-18
  if (n_scalars < 32) return 3;
  double l = __builtin_log2((double)(uint32_t)n_scalars);
  return (unsigned)__builtin_ceil(2.0 / 3.0 * l + 2.0);
}

// the per-window task of msm_inner (:30-70): windows [first, first + count) of the scalars, one task each
template <class C>
static void msm_window_sums(const AffineP<C>* bases, size_t n_bases, const Big* scalars, size_t n_scalars, unsigned c,
                            size_t first, size_t count, ProjP<C>* window_sums, unsigned nthreads) {
  Big fr_one;
  memset(&fr_one, 0, sizeof(fr_one));
  fr_one.l[0] = 1;
  const size_t n = n_bases < n_scalars ? n_bases : n_scalars;  // zip
  parallel_for(count, nthreads, [&](size_t k) {
    const unsigned w_start = (unsigned)((first + k) * c);
    ProjP<C> res = ProjP<C>::zero();
    std::vector<ProjP<C>> buckets(((size_t)1 << c) - 1, ProjP<C>::zero());
    for (size_t i = 0; i < n; i++) {
      const Big& s = scalars[i];
      if (big_is_zero(s)) continue;
      if (big_eq(s, fr_one)) {
        if (w_start == 0) res.add_assign_mixed(bases[i]);
      } else {
        u64 d = big_window(s, w_start, c);
        if (d != 0) buckets[d - 1].add_assign_mixed(bases[i]);
      }
    }
    batch_normalization(buckets);
    ProjP<C> running = ProjP<C>::zero();
    for (size_t b = buckets.size(); b-- > 0;) {
      running.add_assign_mixed(buckets[b].into_affine());
      res.add_assign(running);
    }
    window_sums[k] = res;
  });
}

template <class C>
static ProjP<C> msm_inner(const AffineP<C>* bases, size_t n_bases, const Big* scalars, size_t n_scalars,
                          unsigned nthreads) {
  const unsigned c = ln_window(n_scalars);
  const unsigned num_bits = 753;
  const size_t windows = (num_bits + c - 1) / c;   // (0..num_bits).step_by(c)
  std::vector<ProjP<C>> window_sums(windows);
  msm_window_sums<C>(bases, n_bases, scalars, n_scalars, c, 0, windows, window_sums.data(), nthreads);
  ProjP<C> total = ProjP<C>::zero();
  for (size_t wi = window_sums.size(); wi-- > 1;) {
    total.add_assign(window_sums[wi]);
    for (unsigned k = 0; k < c; k++) total.double_in_place();
  }
  total.add_assign(window_sums[0]);
  return total;
}

// ------------------------------------------------------------------------------------------
// FFT (fft/domain.rs)
// ------------------------------------------------------------------------------------------
template <int F>
static Fq<F> fq_pow_big(const Fq<F>& a, u64 e) {
  return fq_pow<F>(a, e);
}
template <int F>
static void serial_fft(Fq<F>* a, size_t n, const Fq<F>& omega, unsigned log_n) {  // :315-358
  for (size_t k = 0; k < n; k++) {
    size_t rk = 0;
    for (unsigned b = 0; b < log_n; b++) rk |= ((k >> b) & 1) << (log_n - 1 - b);
    if (k < rk) {
      Fq<F> t = a[k];
      a[k] = a[rk];
      a[rk] = t;
    }
  }
  size_t m = 1;
  for (unsigned s = 0; s < log_n; s++) {
    Fq<F> w_m = fq_pow<F>(omega, (u64)(n / (2 * m)));
    for (size_t k = 0; k < n; k += 2 * m) {
      Fq<F> w = fq_one<F>();
      for (size_t j = 0; j < m; j++) {
        Fq<F> t = fq_mul(a[k + j + m], w);
        Fq<F> tmp = fq_sub(a[k + j], t);
        a[k + j + m] = tmp;
        a[k + j] = fq_add(a[k + j], t);
        w = fq_mul(w, w_m);
      }
    }
    m *= 2;
  }
}
template <int F>
static void parallel_fft(Fq<F>* a, size_t n, const Fq<F>& omega, unsigned log_n, unsigned log_cpus,
                         unsigned nthreads) {  // :360-416
  const size_t num_cpus = (size_t)1 << log_cpus;
  const unsigned log_new_n = log_n - log_cpus;
  const size_t new_n = (size_t)1 << log_new_n;
  std::vector<std::vector<Fq<F>>> tmp(num_cpus, std::vector<Fq<F>>(new_n, fq_zero<F>()));
  Fq<F> new_omega = fq_pow<F>(omega, (u64)num_cpus);
  parallel_for(num_cpus, nthreads, [&](size_t j) {
    Fq<F> omega_j = fq_pow<F>(omega, (u64)j);
    Fq<F> omega_step = fq_pow<F>(omega, (u64)j << log_new_n);
    Fq<F> elt = fq_one<F>();
    for (size_t i = 0; i < new_n; i++) {
      for (size_t s = 0; s < num_cpus; s++) {
        size_t idx = (i + (s << log_new_n)) % n;
        tmp[j][i] = fq_add(tmp[j][i], fq_mul(a[idx], elt));
        elt = fq_mul(elt, omega_step);
      }
      elt = fq_mul(elt, omega_j);
    }
    serial_fft<F>(tmp[j].data(), new_n, new_omega, log_new_n);
  });
  const size_t mask = num_cpus - 1;
  const size_t chunk = (n + nthreads - 1) / (nthreads ? nthreads : 1);
  parallel_for((n + chunk - 1) / chunk, nthreads, [&](size_t ci) {
    size_t lo = ci * chunk, hi = lo + chunk < n ? lo + chunk : n;
    for (size_t idx = lo; idx < hi; idx++) a[idx] = tmp[idx & mask][idx >> log_cpus];
  });
}
template <int F>
static void best_fft(Fq<F>* a, size_t n, const Fq<F>& omega, unsigned log_n, unsigned nthreads) {  // :305-313
  unsigned log_cpus = 0;
  while (((size_t)2 << log_cpus) <= nthreads) log_cpus++;  // Worker::log_num_cpus = floor(log2)
  if (log_n <= log_cpus) serial_fft<F>(a, n, omega, log_n);
  else parallel_fft<F>(a, n, omega, log_n, log_cpus, nthreads);
}
template <int F>
static void distribute_powers(Fq<F>* a, size_t n, const Fq<F>& g, unsigned nthreads) {  // :140-152
  const size_t chunk = (n + nthreads - 1) / (nthreads ? nthreads : 1);
  if (chunk == 0) return;
  parallel_for((n + chunk - 1) / chunk, nthreads, [&](size_t ci) {
    size_t lo = ci * chunk, hi = lo + chunk < n ? lo + chunk : n;
    Fq<F> u = fq_pow<F>(g, (u64)lo);
    for (size_t i = lo; i < hi; i++) {
      a[i] = fq_mul(a[i], u);
      u = fq_mul(u, g);
    }
  });
}
template <int F>
static int domain_transform(Fq<F>* a, unsigned log_n, int mode, unsigned nthreads) {
  if (log_n >= FP[F].two_adicity) return 4;  // EvaluationDomain::new -> None (:70-72)
  const size_t n = (size_t)1 << log_n;
  Fq<F> group_gen = fq_root_of_unity<F>();
  for (unsigned k = log_n; k < FP[F].two_adicity; k++) group_gen = fq_square(group_gen);  // :76-79
  Fq<F> gen = fq_from_u64<F>(17);
  if (mode == 2) distribute_powers<F>(a, n, gen, nthreads);  // coset_fft :163-166
  if (mode == 0 || mode == 2) {
    best_fft<F>(a, n, group_gen, log_n, nthreads);
  } else {
    Fq<F> gi = fq_inverse(group_gen);
    best_fft<F>(a, n, gi, log_n, nthreads);
    Fq<F> size_inv = fq_inverse(fq_from_u64<F>((u64)n));
    const size_t chunk = (n + nthreads - 1) / (nthreads ? nthreads : 1);
    parallel_for((n + chunk - 1) / chunk, nthreads, [&](size_t ci) {  // :137
      size_t lo = ci * chunk, hi = lo + chunk < n ? lo + chunk : n;
      for (size_t i = lo; i < hi; i++) a[i] = fq_mul(a[i], size_inv);
    });
    if (mode == 3) distribute_powers<F>(a, n, fq_inverse(gen), nthreads);  // :176-179
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// marshalling + extern "C"
// ------------------------------------------------------------------------------------------
template <class C>
static void load_affine(const u64* coords, const uint8_t* inf, size_t n, std::vector<AffineP<C>>& out) {
  typedef typename C::BF BF;
  out.resize(n);
  for (size_t i = 0; i < n; i++) {
    memcpy(&out[i].x, coords + i * 2 * BF::K * N, sizeof(BF));
    memcpy(&out[i].y, coords + (i * 2 + 1) * BF::K * N, sizeof(BF));
    out[i].infinity = inf ? inf[i] != 0 : false;
  }
}
template <class C>
static void store_proj(const ProjP<C>& p, u64* out) {
  typedef typename C::BF BF;
  memcpy(out, &p.x, sizeof(BF));
  memcpy(out + BF::K * N, &p.y, sizeof(BF));
  memcpy(out + 2 * BF::K * N, &p.z, sizeof(BF));
}
template <class C>
static ProjP<C> load_proj(const u64* in) {
  typedef typename C::BF BF;
  ProjP<C> p;
  memcpy(&p.x, in, sizeof(BF));
  memcpy(&p.y, in + BF::K * N, sizeof(BF));
  memcpy(&p.z, in + 2 * BF::K * N, sizeof(BF));
  return p;
}

template <class C>
static int msm_entry(const u64* coords, const uint8_t* inf, size_t n_bases, const u64* scalars, size_t n_scalars,
                     u64* out_xyz, unsigned nthreads) {
  std::vector<AffineP<C>> bases;
  size_t n = n_bases < n_scalars ? n_bases : n_scalars;
  load_affine<C>(coords, inf, n, bases);
  ProjP<C> r = msm_inner<C>(bases.data(), n, (const Big*)scalars, n_scalars, nthreads);
  store_proj<C>(r, out_xyz);
  return 0;
}

// A bounded SAMPLE of one MSM for the CPU baseline (bench.py): `count` of the reference's per-window
// tasks, windows [first, first + count), with the window size c of the FULL input; out receives the
// `count` window sums (projective).  Every window task costs the same (n mixed additions into 2^c - 1
// buckets + the running-sum reduction), so count / windows of the MSM's time is measured exactly.
template <class C>
static int msm_windows_entry(const u64* coords, const uint8_t* inf, size_t n_bases, const u64* scalars,
                             size_t n_scalars, unsigned first, unsigned count, u64* out_xyz, unsigned nthreads) {
  typedef typename C::BF BF;
  std::vector<AffineP<C>> bases;
  size_t n = n_bases < n_scalars ? n_bases : n_scalars;
  load_affine<C>(coords, inf, n, bases);
  const unsigned c = ln_window(n_scalars);
  const unsigned windows = (753 + c - 1) / c;
  if (first + count > windows) return 1;
  std::vector<ProjP<C>> sums(count);
  msm_window_sums<C>(bases.data(), n, (const Big*)scalars, n_scalars, c, first, count, sums.data(), nthreads);
  for (unsigned k = 0; k < count; k++) store_proj<C>(sums[k], out_xyz + (size_t)k * 3 * BF::K * N);
  return 0;
}

// P_i = P_0 + i * D for i < n (affine, normalised): the bench / large-test base generator.
// Each thread starts its segment at P_0 + lo * D by double-and-add, walks with mixed additions
// and batch-normalises.
template <class C>
static int walk_entry(const u64* p0_xy, const u64* d_xy, size_t n, u64* out_coords, unsigned nthreads) {
  typedef typename C::BF BF;
  std::vector<AffineP<C>> in;
  load_affine<C>(p0_xy, nullptr, 1, in);
  AffineP<C> P0 = in[0];
  load_affine<C>(d_xy, nullptr, 1, in);
  AffineP<C> D = in[0];
  const size_t seg = 4096;
  parallel_for((n + seg - 1) / seg, nthreads, [&](size_t si) {
    size_t lo = si * seg, hi = lo + seg < n ? lo + seg : n;
    ProjP<C> acc = ProjP<C>::zero();  // lo * D
    for (int b = 63; b >= 0; b--) {
      acc.double_in_place();
      if ((lo >> b) & 1) acc.add_assign_mixed(D);
    }
    acc.add_assign_mixed(P0);
    std::vector<ProjP<C>> pts(hi - lo);
    for (size_t i = lo; i < hi; i++) {
      pts[i - lo] = acc;
      acc.add_assign_mixed(D);
    }
    batch_normalization(pts);
    for (size_t i = lo; i < hi; i++) {
      memcpy(out_coords + i * 2 * BF::K * N, &pts[i - lo].x, sizeof(BF));
      memcpy(out_coords + (i * 2 + 1) * BF::K * N, &pts[i - lo].y, sizeof(BF));
    }
  });
  return 0;
}

template <class C>
static int point_entry(int op, const u64* a, const u64* b, u64* out) {
  // op 0: proj(a) + proj(b) ; 1: 2 * proj(a) ; 2: scalar b * affine a ; 3: into_affine(proj a) -> x,y,(inf in out[2k*12])
  // op 4: proj(a) += affine(b) (add_assign_mixed)
  typedef typename C::BF BF;
  if (op == 0) {
    ProjP<C> p = load_proj<C>(a), q = load_proj<C>(b);
    p.add_assign(q);
    store_proj<C>(p, out);
  } else if (op == 1) {
    ProjP<C> p = load_proj<C>(a);
    p.double_in_place();
    store_proj<C>(p, out);
  } else if (op == 2) {
    std::vector<AffineP<C>> in;
    load_affine<C>(a, nullptr, 1, in);
    const Big* s = (const Big*)b;
    ProjP<C> base{in[0].x, in[0].y, BF::one()}, res = ProjP<C>::zero();
    bool found = false;  // short_weierstrass_projective.rs:520-538
    for (int i = 64 * N - 1; i >= 0; i--) {
      bool bit = (s->l[i >> 6] >> (i & 63)) & 1;
      if (found) res.double_in_place();
      else found = bit;
      if (bit) res.add_assign(base);
    }
    store_proj<C>(res, out);
  } else if (op == 3) {
    AffineP<C> q = load_proj<C>(a).into_affine();
    memcpy(out, &q.x, sizeof(BF));
    memcpy(out + BF::K * N, &q.y, sizeof(BF));
    out[2 * BF::K * N] = q.infinity ? 1 : 0;
  } else if (op == 4) {
    ProjP<C> p = load_proj<C>(a);
    std::vector<AffineP<C>> in;
    load_affine<C>(b, nullptr, 1, in);
    p.add_assign_mixed(in[0]);
    store_proj<C>(p, out);
  } else {
    return 1;
  }
  return 0;
}

template <int F>
static int field_entry(int op, const u64* a, const u64* b, u64* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    Fq<F> x, y, r;
    memcpy(&x, a + i * N, sizeof(x));
    if (b) memcpy(&y, b + i * N, sizeof(y));
    else y = x;
    switch (op) {
      case 0: r = fq_mul(x, y); break;
      case 1: r = fq_add(x, y); break;
      case 2: r = fq_sub(x, y); break;
      case 3: r = fq_square(x); break;
      case 4: r = fq_neg(x); break;
      case 5: r = fq_is_zero(x) ? x : fq_inverse(x); break;
      case 6: r = fq_from_repr<F>(x.v); break;
      case 7: r.v = fq_into_repr(x); break;
      default: return 1;
    }
    memcpy(out + i * N, &r, sizeof(r));
  }
  return 0;
}

template <class BF>
static int ext_entry(int op, const u64* a, const u64* b, u64* out) {
  BF x, y, r;
  memcpy(&x, a, sizeof(BF));
  memcpy(&y, b ? b : a, sizeof(BF));
  switch (op) {
    case 0: r = x * y; break;
    case 1: r = x + y; break;
    case 2: r = x - y; break;
    case 3: r = x.square(); break;
    case 4: r = x.neg(); break;
    case 5: r = x.inverse(); break;
    default: return 1;
  }
  memcpy(out, &r, sizeof(BF));
  return 0;
}

extern "C" {
unsigned ref753_hardware_threads() {
  unsigned t = std::thread::hardware_concurrency();
  return t ? t : 1;
}
// constants for the tests: which = 0 R, 1 R2, 2 ROOT_OF_UNITY (Montgomery), 3 INV (limb 0)
int ref753_constant(int field, int which, u64* out) {
  if (field < 0 || field > 1) return 1;
  memset(out, 0, 96);
  if (which == 0) memcpy(out, FP[field].r.l, 96);
  else if (which == 1) memcpy(out, FP[field].r2.l, 96);
  else if (which == 2) {
    if (field == 0) {
      Fq<0> r = fq_root_of_unity<0>();
      memcpy(out, &r, 96);
    } else {
      Fq<1> r = fq_root_of_unity<1>();
      memcpy(out, &r, 96);
    }
  } else if (which == 3) out[0] = FP[field].inv;
  else return 1;
  return 0;
}
int ref753_field_op(int field, int op, const u64* a, const u64* b, u64* out, size_t n) {
  return field == 0 ? field_entry<0>(op, a, b, out, n) : field_entry<1>(op, a, b, out, n);
}
int ref753_ext_op(int ext, int op, const u64* a, const u64* b, u64* out) {
  return ext == 2 ? ext_entry<K2<0, 13>>(op, a, b, out) : ext_entry<K3<1, 11>>(op, a, b, out);
}
int ref753_point_op(int group, int op, const u64* a, const u64* b, u64* out) {
  switch (group) {
    case 0: return point_entry<M4G1>(op, a, b, out);
    case 1: return point_entry<M4G2>(op, a, b, out);
    case 2: return point_entry<M6G1>(op, a, b, out);
    case 3: return point_entry<M6G2>(op, a, b, out);
  }
  return 1;
}
// VariableBaseMSM::multi_scalar_mul; out = GroupProjective {x, y, z}
int ref753_msm(int group, const u64* coords, const uint8_t* inf, size_t n_bases, const u64* scalars,
               size_t n_scalars, u64* out_xyz, unsigned nthreads) {
  switch (group) {
    case 0: return msm_entry<M4G1>(coords, inf, n_bases, scalars, n_scalars, out_xyz, nthreads);
    case 1: return msm_entry<M4G2>(coords, inf, n_bases, scalars, n_scalars, out_xyz, nthreads);
    case 2: return msm_entry<M6G1>(coords, inf, n_bases, scalars, n_scalars, out_xyz, nthreads);
    case 3: return msm_entry<M6G2>(coords, inf, n_bases, scalars, n_scalars, out_xyz, nthreads);
  }
  return 1;
}
// windows [first, first + count) of the same MSM: out = count GroupProjective window sums
int ref753_msm_windows(int group, const u64* coords, const uint8_t* inf, size_t n_bases, const u64* scalars,
                       size_t n_scalars, unsigned first, unsigned count, u64* out_xyz, unsigned nthreads) {
  switch (group) {
    case 0: return msm_windows_entry<M4G1>(coords, inf, n_bases, scalars, n_scalars, first, count, out_xyz, nthreads);
    case 1: return msm_windows_entry<M4G2>(coords, inf, n_bases, scalars, n_scalars, first, count, out_xyz, nthreads);
    case 2: return msm_windows_entry<M6G1>(coords, inf, n_bases, scalars, n_scalars, first, count, out_xyz, nthreads);
    case 3: return msm_windows_entry<M6G2>(coords, inf, n_bases, scalars, n_scalars, first, count, out_xyz, nthreads);
  }
  return 1;
}
unsigned ref753_msm_window_bits(size_t n_scalars) { return ln_window(n_scalars); }
int ref753_walk(int group, const u64* p0_xy, const u64* d_xy, size_t n, u64* out_coords, unsigned nthreads) {
  switch (group) {
    case 0: return walk_entry<M4G1>(p0_xy, d_xy, n, out_coords, nthreads);
    case 1: return walk_entry<M4G2>(p0_xy, d_xy, n, out_coords, nthreads);
    case 2: return walk_entry<M6G1>(p0_xy, d_xy, n, out_coords, nthreads);
    case 3: return walk_entry<M6G2>(p0_xy, d_xy, n, out_coords, nthreads);
  }
  return 1;
}
// EvaluationDomain::{fft, ifft, coset_fft, coset_ifft}_in_place on n = 2^log_n elements
int ref753_fft(int field, u64* data, unsigned log_n, int mode, unsigned nthreads) {
  if (mode < 0 || mode > 3) return 1;
  if (nthreads == 0) nthreads = 1;
  return field == 0 ? domain_transform<0>((Fq<0>*)data, log_n, mode, nthreads)
                    : domain_transform<1>((Fq<1>*)data, log_n, mode, nthreads);
}
}
