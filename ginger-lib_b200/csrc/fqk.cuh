// Base-field wrappers with a uniform static interface: Fp (k = 1), Fp2 = Fq[u]/(u^2 - NR),
// Fp3 = Fq[u]/(u^3 - NR).  The curve code (ec.cuh) is written once against this interface.
//
// Replaces (reference, relative to /root/reference/algebra/src):
//   fields/models/fp2.rs:387-401 (mul), :128-144 (square), :83-127 (inverse)
//   fields/models/fp3.rs:451-478 (mul), :165-185 (square), :107-164 (inverse)
//   fields/mnt4753/fq2.rs:17-32  NONRESIDUE = 13 ; fields/mnt6753/fq3.rs:17-31 NONRESIDUE = 11
// The reference multiplies by the non-residue with a full-width Montgomery product
// (fp2.rs:36-38, fp3.rs:45-47); 13 and 11 are small, so here it is an addition chain with
// the same canonical value.
#pragma once
#include "fq.cuh"

namespace g753 {

template <int FID>
struct Fp {
  static constexpr int K = 1;
  static constexpr int FIELD = FID;
  Fq c0;
  static G753_HD Fp zero() { return Fp{fq_zero<FID>()}; }
  static G753_HD Fp one() { return Fp{fq_one<FID>()}; }
  static G753_HD bool is_zero(const Fp& a) { return fq_is_zero(a.c0); }
  static G753_HD bool eq(const Fp& a, const Fp& b) { return fq_eq(a.c0, b.c0); }
  static G753_HD Fp add(const Fp& a, const Fp& b) { return Fp{fq_add<FID>(a.c0, b.c0)}; }
  static G753_HD Fp sub(const Fp& a, const Fp& b) { return Fp{fq_sub<FID>(a.c0, b.c0)}; }
  static G753_HD Fp neg(const Fp& a) { return Fp{fq_neg<FID>(a.c0)}; }
  static G753_HD Fp dbl(const Fp& a) { return Fp{fq_dbl<FID>(a.c0)}; }
  static G753_HD Fp mul(const Fp& a, const Fp& b) { return Fp{fq_mulc<FID>(a.c0, b.c0)}; }
  static G753_HD Fp sqr(const Fp& a) { return Fp{fq_sqrc<FID>(a.c0)}; }
  static G753_HD Fp inv(const Fp& a) { return Fp{fq_inv<FID>(a.c0)}; }
  template <unsigned KK>
  static G753_HD Fp mul_small(const Fp& a) { return Fp{fq_mul_small<FID, KK>(a.c0)}; }
};

template <int FID, unsigned NR>
struct Fp2 {
  static constexpr int K = 2;
  static constexpr int FIELD = FID;
  Fq c0, c1;
  static G753_HD Fp2 zero() { return Fp2{fq_zero<FID>(), fq_zero<FID>()}; }
  static G753_HD Fp2 one() { return Fp2{fq_one<FID>(), fq_zero<FID>()}; }
  static G753_HD bool is_zero(const Fp2& a) { return fq_is_zero(a.c0) && fq_is_zero(a.c1); }
  static G753_HD bool eq(const Fp2& a, const Fp2& b) { return fq_eq(a.c0, b.c0) && fq_eq(a.c1, b.c1); }
  static G753_HD Fp2 add(const Fp2& a, const Fp2& b) {
    return Fp2{fq_add<FID>(a.c0, b.c0), fq_add<FID>(a.c1, b.c1)};
  }
  static G753_HD Fp2 sub(const Fp2& a, const Fp2& b) {
    return Fp2{fq_sub<FID>(a.c0, b.c0), fq_sub<FID>(a.c1, b.c1)};
  }
  static G753_HD Fp2 neg(const Fp2& a) { return Fp2{fq_neg<FID>(a.c0), fq_neg<FID>(a.c1)}; }
  static G753_HD Fp2 dbl(const Fp2& a) { return Fp2{fq_dbl<FID>(a.c0), fq_dbl<FID>(a.c1)}; }
  // Karatsuba: 3 base multiplications
  static G753_NI Fp2 mul(const Fp2& a, const Fp2& b) {
    Fq v0 = fq_mulc<FID>(a.c0, b.c0);
    Fq v1 = fq_mulc<FID>(a.c1, b.c1);
    Fq s = fq_mulc<FID>(fq_add<FID>(a.c0, a.c1), fq_add<FID>(b.c0, b.c1));
    Fp2 r;
    r.c1 = fq_sub<FID>(fq_sub<FID>(s, v0), v1);
    r.c0 = fq_add<FID>(v0, fq_mul_small<FID, NR>(v1));
    return r;
  }
  // complex squaring: 2 base multiplications
  static G753_NI Fp2 sqr(const Fp2& a) {
    Fq v = fq_mulc<FID>(a.c0, a.c1);
    Fq t = fq_mulc<FID>(fq_add<FID>(a.c0, a.c1), fq_add<FID>(a.c0, fq_mul_small<FID, NR>(a.c1)));
    Fp2 r;
    r.c0 = fq_sub<FID>(fq_sub<FID>(t, v), fq_mul_small<FID, NR>(v));
    r.c1 = fq_dbl<FID>(v);
    return r;
  }
  static G753_NI Fp2 inv(const Fp2& a) {
    Fq d = fq_sub<FID>(fq_sqrc<FID>(a.c0), fq_mul_small<FID, NR>(fq_sqrc<FID>(a.c1)));
    d = fq_inv<FID>(d);
    return Fp2{fq_mulc<FID>(a.c0, d), fq_neg<FID>(fq_mulc<FID>(a.c1, d))};
  }
};

template <int FID, unsigned NR>
struct Fp3 {
  static constexpr int K = 3;
  static constexpr int FIELD = FID;
  Fq c0, c1, c2;
  static G753_HD Fp3 zero() { return Fp3{fq_zero<FID>(), fq_zero<FID>(), fq_zero<FID>()}; }
  static G753_HD Fp3 one() { return Fp3{fq_one<FID>(), fq_zero<FID>(), fq_zero<FID>()}; }
  static G753_HD bool is_zero(const Fp3& a) {
    return fq_is_zero(a.c0) && fq_is_zero(a.c1) && fq_is_zero(a.c2);
  }
  static G753_HD bool eq(const Fp3& a, const Fp3& b) {
    return fq_eq(a.c0, b.c0) && fq_eq(a.c1, b.c1) && fq_eq(a.c2, b.c2);
  }
  static G753_HD Fp3 add(const Fp3& a, const Fp3& b) {
    return Fp3{fq_add<FID>(a.c0, b.c0), fq_add<FID>(a.c1, b.c1), fq_add<FID>(a.c2, b.c2)};
  }
  static G753_HD Fp3 sub(const Fp3& a, const Fp3& b) {
    return Fp3{fq_sub<FID>(a.c0, b.c0), fq_sub<FID>(a.c1, b.c1), fq_sub<FID>(a.c2, b.c2)};
  }
  static G753_HD Fp3 neg(const Fp3& a) {
    return Fp3{fq_neg<FID>(a.c0), fq_neg<FID>(a.c1), fq_neg<FID>(a.c2)};
  }
  static G753_HD Fp3 dbl(const Fp3& a) {
    return Fp3{fq_dbl<FID>(a.c0), fq_dbl<FID>(a.c1), fq_dbl<FID>(a.c2)};
  }
  // Karatsuba (Devegili et al.): 6 base multiplications
  static G753_NI Fp3 mul(const Fp3& a, const Fp3& b) {
    Fq v0 = fq_mulc<FID>(a.c0, b.c0);
    Fq v1 = fq_mulc<FID>(a.c1, b.c1);
    Fq v2 = fq_mulc<FID>(a.c2, b.c2);
    Fq t12 = fq_mulc<FID>(fq_add<FID>(a.c1, a.c2), fq_add<FID>(b.c1, b.c2));
    Fq t01 = fq_mulc<FID>(fq_add<FID>(a.c0, a.c1), fq_add<FID>(b.c0, b.c1));
    Fq t02 = fq_mulc<FID>(fq_add<FID>(a.c0, a.c2), fq_add<FID>(b.c0, b.c2));
    Fp3 r;
    r.c0 = fq_add<FID>(v0, fq_mul_small<FID, NR>(fq_sub<FID>(fq_sub<FID>(t12, v1), v2)));
    r.c1 = fq_add<FID>(fq_sub<FID>(fq_sub<FID>(t01, v0), v1), fq_mul_small<FID, NR>(v2));
    r.c2 = fq_add<FID>(fq_sub<FID>(fq_sub<FID>(t02, v0), v2), v1);
    return r;
  }
  // Chung-Hasan SQR2: 2 multiplications + 3 squarings
  static G753_NI Fp3 sqr(const Fp3& a) {
    Fq s0 = fq_sqrc<FID>(a.c0);
    Fq s1 = fq_dbl<FID>(fq_mulc<FID>(a.c0, a.c1));
    Fq s2 = fq_sqrc<FID>(fq_add<FID>(fq_sub<FID>(a.c0, a.c1), a.c2));
    Fq s3 = fq_dbl<FID>(fq_mulc<FID>(a.c1, a.c2));
    Fq s4 = fq_sqrc<FID>(a.c2);
    Fp3 r;
    r.c0 = fq_add<FID>(s0, fq_mul_small<FID, NR>(s3));
    r.c1 = fq_add<FID>(s1, fq_mul_small<FID, NR>(s4));
    r.c2 = fq_sub<FID>(fq_sub<FID>(fq_add<FID>(fq_add<FID>(s1, s2), s3), s0), s4);
    return r;
  }
  static G753_NI Fp3 inv(const Fp3& a) {
    Fq t0 = fq_sub<FID>(fq_sqrc<FID>(a.c0), fq_mul_small<FID, NR>(fq_mulc<FID>(a.c1, a.c2)));
    Fq t1 = fq_sub<FID>(fq_mul_small<FID, NR>(fq_sqrc<FID>(a.c2)), fq_mulc<FID>(a.c0, a.c1));
    Fq t2 = fq_sub<FID>(fq_sqrc<FID>(a.c1), fq_mulc<FID>(a.c0, a.c2));
    Fq d = fq_add<FID>(fq_mulc<FID>(a.c0, t0),
                       fq_mul_small<FID, NR>(fq_add<FID>(fq_mulc<FID>(a.c2, t1), fq_mulc<FID>(a.c1, t2))));
    d = fq_inv<FID>(d);
    return Fp3{fq_mulc<FID>(t0, d), fq_mulc<FID>(t1, d), fq_mulc<FID>(t2, d)};
  }
};

typedef Fp<0> FqM4;        // mnt4753::Fq  (= mnt6753::Fr)
typedef Fp<1> FqM6;        // mnt6753::Fq  (= mnt4753::Fr)
typedef Fp2<0, 13> Fq2M4;  // mnt4753::Fq2
typedef Fp3<1, 11> Fq3M6;  // mnt6753::Fq3

}  // namespace g753
