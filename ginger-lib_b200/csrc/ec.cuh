// Short-Weierstrass group law for the four prover groups, in XYZZ coordinates
// (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; infinity <=> ZZ == 0).
//
// Replaces (reference, relative to /root/reference/algebra/src):
//   curves/models/short_weierstrass_projective.rs:481-519  add_assign_mixed -> xyzz_madd
//   curves/models/short_weierstrass_projective.rs:444-479  double_in_place  -> xyzz_dbl
//   curves/models/short_weierstrass_projective.rs:574-617  AddAssign        -> xyzz_add
//   curves/models/short_weierstrass_projective.rs:520-538  mul_assign       -> xyzz_scalar_mul
//   curves/mnt4753/g1.rs:18-50 (a = 2), curves/mnt6753/g1.rs:18-52 (a = 11),
//   curves/mnt4753/g2.rs:112-118 and curves/mnt6753/g2.rs:148-155 (mul_by_a on the twists)
// The reference works in homogeneous projective (X:Y:Z) coordinates; the group element
// computed is the same, and xyzz_to_projective hands back a homogeneous representative,
// which is what VariableBaseMSM::multi_scalar_mul returns (any representative is accepted,
// SURVEY.md 8b "Result contract").  Unlike the reference's madd formula, P + (-P) is
// handled explicitly.
#pragma once
#include "fqk.cuh"

namespace g753 {

struct CurveM4G1 {
  typedef FqM4 F;
  static constexpr int ID = 0;
  static G753_HD F mul_by_a(const F& x) { return F::dbl(x); }  // a = 2
};
struct CurveM4G2 {
  typedef Fq2M4 F;
  static constexpr int ID = 1;
  // twist a' = (26, 0): coefficient-wise (curves/mnt4753/g2.rs:112-118)
  static G753_HD F mul_by_a(const F& x) {
    return F{fq_mul_small<0, 26>(x.c0), fq_mul_small<0, 26>(x.c1)};
  }
};
struct CurveM6G1 {
  typedef FqM6 F;
  static constexpr int ID = 2;
  static G753_HD F mul_by_a(const F& x) { return F::template mul_small<11>(x); }  // a = 11
};
struct CurveM6G2 {
  typedef Fq3M6 F;
  static constexpr int ID = 3;
  // twist a' = 11 u^2: (c0, c1, c2) -> (121 c1, 121 c2, 11 c0) (curves/mnt6753/g2.rs:148-155)
  static G753_HD F mul_by_a(const F& x) {
    return F{fq_mul_small<1, 121>(x.c1), fq_mul_small<1, 121>(x.c2), fq_mul_small<1, 11>(x.c0)};
  }
};

template <class C>
struct Affine {
  typename C::F x, y;  // (0, 0) encodes the point at infinity (not on any of the four curves)
};
template <class C>
struct Xyzz {
  typename C::F x, y, zz, zzz;
};

template <class C>
G753_HD bool affine_is_inf(const Affine<C>& q) {
  typedef typename C::F F;
  return F::is_zero(q.x) && F::is_zero(q.y);
}
template <class C>
G753_HD bool xyzz_is_inf(const Xyzz<C>& p) {
  return C::F::is_zero(p.zz);
}
template <class C>
G753_HD Xyzz<C> xyzz_inf() {
  typedef typename C::F F;
  return Xyzz<C>{F::zero(), F::one(), F::zero(), F::zero()};
}
template <class C>
G753_HD Xyzz<C> xyzz_from_affine(const Affine<C>& q) {
  typedef typename C::F F;
  if (affine_is_inf<C>(q)) return xyzz_inf<C>();
  return Xyzz<C>{q.x, q.y, F::one(), F::one()};
}

// 2 * (affine q), q finite: mdbl-2008-s-1 with general a
template <class C>
G753_NI Xyzz<C> xyzz_mdbl(const Affine<C>& q) {
  typedef typename C::F F;
  F u = F::dbl(q.y);
  if (F::is_zero(u)) return xyzz_inf<C>();
  F v = F::sqr(u);
  F w = F::mul(u, v);
  F s = F::mul(q.x, v);
  F xx = F::sqr(q.x);
  F m = F::add(F::add(F::dbl(xx), xx), C::mul_by_a(F::one()));
  Xyzz<C> r;
  r.x = F::sub(F::sqr(m), F::dbl(s));
  r.y = F::sub(F::mul(m, F::sub(s, r.x)), F::mul(w, q.y));
  r.zz = v;
  r.zzz = w;
  return r;
}

// p = 2p: dbl-2008-s-1
template <class C>
G753_NI void xyzz_dbl(Xyzz<C>& p) {
  typedef typename C::F F;
  if (xyzz_is_inf<C>(p)) return;
  F u = F::dbl(p.y);
  if (F::is_zero(u)) {
    p = xyzz_inf<C>();
    return;
  }
  F v = F::sqr(u);
  F w = F::mul(u, v);
  F s = F::mul(p.x, v);
  F xx = F::sqr(p.x);
  F m = F::add(F::add(F::dbl(xx), xx), C::mul_by_a(F::sqr(p.zz)));
  F x3 = F::sub(F::sqr(m), F::dbl(s));
  F y3 = F::sub(F::mul(m, F::sub(s, x3)), F::mul(w, p.y));
  p.x = x3;
  p.y = y3;
  p.zz = F::mul(v, p.zz);
  p.zzz = F::mul(w, p.zzz);
}

// p += q (q affine): madd-2008-s, 8M + 2S
template <class C>
G753_NI void xyzz_madd(Xyzz<C>& p, const Affine<C>& q) {
  typedef typename C::F F;
  if (affine_is_inf<C>(q)) return;
  if (xyzz_is_inf<C>(p)) {
    p = Xyzz<C>{q.x, q.y, F::one(), F::one()};
    return;
  }
  F pp = F::sub(F::mul(q.x, p.zz), p.x);   // P = U2 - X1
  F r = F::sub(F::mul(q.y, p.zzz), p.y);   // R = S2 - Y1
  if (F::is_zero(pp)) {
    if (F::is_zero(r)) p = xyzz_mdbl<C>(q);
    else p = xyzz_inf<C>();
    return;
  }
  F p2 = F::sqr(pp);
  F p3 = F::mul(pp, p2);
  F qq = F::mul(p.x, p2);
  F x3 = F::sub(F::sub(F::sqr(r), p3), F::dbl(qq));
  F y3 = F::sub(F::mul(r, F::sub(qq, x3)), F::mul(p.y, p3));
  p.x = x3;
  p.y = y3;
  p.zz = F::mul(p.zz, p2);
  p.zzz = F::mul(p.zzz, p3);
}

// p += q (both XYZZ): add-2008-s, 12M + 2S
template <class C>
G753_NI void xyzz_add(Xyzz<C>& p, const Xyzz<C>& q) {
  typedef typename C::F F;
  if (xyzz_is_inf<C>(q)) return;
  if (xyzz_is_inf<C>(p)) {
    p = q;
    return;
  }
  F u1 = F::mul(p.x, q.zz);
  F s1 = F::mul(p.y, q.zzz);
  F pp = F::sub(F::mul(q.x, p.zz), u1);
  F r = F::sub(F::mul(q.y, p.zzz), s1);
  if (F::is_zero(pp)) {
    if (F::is_zero(r)) xyzz_dbl<C>(p);
    else p = xyzz_inf<C>();
    return;
  }
  F p2 = F::sqr(pp);
  F p3 = F::mul(pp, p2);
  F qq = F::mul(u1, p2);
  F x3 = F::sub(F::sub(F::sqr(r), p3), F::dbl(qq));
  F y3 = F::sub(F::mul(r, F::sub(qq, x3)), F::mul(s1, p3));
  p.x = x3;
  p.y = y3;
  p.zz = F::mul(F::mul(p.zz, q.zz), p2);
  p.zzz = F::mul(F::mul(p.zzz, q.zzz), p3);
}

template <class C>
G753_HD Affine<C> affine_neg(const Affine<C>& q) {
  return Affine<C>{q.x, C::F::neg(q.y)};
}

// homogeneous projective representative (X : Y : Z), x = X/Z, y = Y/Z - the reference's
// GroupProjective layout; infinity is (0 : 1 : 0) as in GroupProjective::zero()
template <class C>
G753_HD void xyzz_to_projective(const Xyzz<C>& p, typename C::F& X, typename C::F& Y, typename C::F& Z) {
  typedef typename C::F F;
  if (xyzz_is_inf<C>(p)) {
    X = F::zero();
    Y = F::one();
    Z = F::zero();
    return;
  }
  X = F::mul(p.x, p.zzz);
  Y = F::mul(p.y, p.zz);
  Z = F::mul(p.zz, p.zzz);
}

// k * q for a 768-bit little-endian scalar (MSB-first double-and-add, as the reference's
// mul_assign); used for tiny inputs and tests, not on the bucket hot loop
template <class C>
G753_NI Xyzz<C> xyzz_scalar_mul(const Affine<C>& q, const uint32_t* k) {
  Xyzz<C> r = xyzz_inf<C>();
  bool started = false;
  for (int i = NL * 32 - 1; i >= 0; i--) {
    if (started) xyzz_dbl<C>(r);
    if ((k[i >> 5] >> (i & 31)) & 1) {
      xyzz_madd<C>(r, q);
      started = true;
    }
  }
  return r;
}

}  // namespace g753
