// Short-Weierstrass group law in XYZZ coordinates on shared-memory slots (see slots.cuh).
//
// Replaces (reference, relative to /root/reference/algebra/src):
//   curves/models/short_weierstrass_projective.rs:481-519  add_assign_mixed -> EcS::madd_g
//   curves/models/short_weierstrass_projective.rs:444-479  double_in_place  -> EcS::dbl
//   curves/models/short_weierstrass_projective.rs:574-617  AddAssign        -> EcS::add / add_g
// The reference's formulas are homogeneous-projective (madd-1998-cmo, dbl-2007-bl,
// add-1998-cmo-2); XYZZ (x = X/ZZ, y = Y/ZZZ) computes the same group element in 8M+2S /
// 12M+2S and converts to the reference's (X:Y:Z) layout once at the end.  Unlike the
// reference's mixed addition, P + (-P) and P + P are both handled explicitly.
//
// A point occupies 4 tower elements = 4K slots in the order X, Y, ZZ, ZZZ; infinity <=> ZZ == 0.
// `W` is the first scratch slot: madd_g / mdbl_g need 3K + NTMP scratch slots (4 on the prime-field
// curves), add / add_g / dbl 4K + NTMP.
#pragma once
#include "slots.cuh"

// Mixed addition of the prime-field accumulation kernel: 0 = madd_g on eight slots (two 128-thread
// blocks per SM), 1 = madd6_g on six (three blocks).  Measured on a B200 at 2^22 points: 224 ms
// against 266 ms - with 12 warps per SM walking a ~95 KB multiplier body the SM waits for
// instructions - and against 229 ms with the rolled multiplier (G753_ROLLED); the eight-slot form stays
// the default.
#ifndef G753_ACC6
#define G753_ACC6 0
#endif

namespace g753 {

// curve descriptors on slots: tower type + multiplication by the curve coefficient a.
// L is the block layout (slots.cuh): Lay<columns, 1> for the prime-field curves and the test-only
// host build, Lay<columns, 4> / Lay<columns, 8> (cooperative towers) for G2 on the device.
template <bool COND, class A, class B>
struct SelectT { typedef A type; };
template <class A, class B>
struct SelectT<false, A, B> { typedef B type; };

template <class L>
struct SCurveM4G1 {
  typedef Tw1<0, L> M;
  static constexpr int ID = 0;
  static G753_D void mul_by_a(int d, int a) { s_dbl<0, L>(d, a); }  // a = 2 (curves/mnt4753/g1.rs:18-33)
};
template <class L>
struct SCurveM6G1 {
  typedef Tw1<1, L> M;
  static constexpr int ID = 2;
  static G753_D void mul_by_a(int d, int a) { s_mul_small<1, L, 11>(d, a); }  // a = 11 (curves/mnt6753/g1.rs:18-33)
};
#if defined(G753_HOST_EMUL)
template <class L>
struct SCurveM4G2 {
  typedef Tw2<0, L, 13> M;
  static constexpr int ID = 1;
  // twist a' = (26, 0), coefficient-wise (curves/mnt4753/g2.rs:112-118)
  static G753_D void mul_by_a(int d, int a) {
    s_mul_small<0, L, 26>(d, a);
    s_mul_small<0, L, 26>(d + 1, a + 1);
  }
};
template <class L>
struct SCurveM6G2 {
  typedef Tw3<1, L, 11> M;
  static constexpr int ID = 3;
  // twist a' = 11 u^2: (c0, c1, c2) -> (121 c1, 121 c2, 11 c0) (curves/mnt6753/g2.rs:148-155); d != a
  static G753_D void mul_by_a(int d, int a) {
    s_mul_small<1, L, 121>(d, a + 1);
    s_mul_small<1, L, 121>(d + 1, a + 2);
    s_mul_small<1, L, 11>(d + 2, a);
  }
};
#else
template <class L>
struct SCurveM4G2 {
  typedef Tw2C<0, L, 13> M;
  static constexpr int ID = 1;
  // twist a' = (26, 0), coefficient-wise (curves/mnt4753/g2.rs:112-118)
  static G753_D void mul_by_a(int d, int a) { M::template mul_small2<26, 26>(d, a); }
};
template <class L>
struct SCurveM6G2 {
  // three lanes: the lazy one-coefficient-per-lane tower (accumulation kernels); four / eight: Karatsuba
  typedef typename SelectT<L::TP == 3, Tw3L<1, L, 11>, Tw3C<1, L, 11>>::type M;
  static constexpr int ID = 3;
  // twist a' = 11 u^2: (c0, c1, c2) -> (121 c1, 121 c2, 11 c0) (curves/mnt6753/g2.rs:148-155); d != a
  static G753_D void mul_by_a(int d, int a) {
    const int r = L::role();
    if (r < 2) s_mul_small<1, L, 121>(d + r, a + 1 + r);
    if (r == 2) s_mul_small<1, L, 11>(d + 2, a);
    L::sync();
  }
};
#endif

template <class SC>
struct EcS {
  typedef typename SC::M M;
  static constexpr int K = M::K;
  static constexpr int PT = 4 * K;                      // slots per XYZZ point
  static constexpr int MADD_SCRATCH = (K == 1 ? 4 : 3) * K + M::NTMP;  // scratch slots of madd_g / mdbl_g
  static constexpr int ADD_SCRATCH = 4 * K + M::NTMP;   // scratch slots of add / add_g / dbl
  // scratch slots of madd_acc_g, the mixed addition of the accumulation kernel: two on the prime-field curves
  static constexpr int ACC_SCRATCH = (K == 1 && G753_ACC6) ? 2 : MADD_SCRATCH;

  static G753_D bool is_inf(int P) { return M::is_zero(P + 2 * K); }
  static G753_D void set_inf(int P) {
    M::set_zero(P);
    M::set_one(P + K);
    M::set_zero(P + 2 * K);
    M::set_zero(P + 3 * K);
  }
  static G753_D void copy(int D, int P) {
    for (int i = 0; i < 4; i++) M::copy(D + i * K, P + i * K);
  }
  static G753_D void ldg(int P, const Fq* g) {
    for (int i = 0; i < 4; i++) M::ldg(P + i * K, g + i * K);
  }
  static G753_D void stg(Fq* g, int P) {
    for (int i = 0; i < 4; i++) M::stg(g + i * K, P + i * K);
  }

  // P = 2 * (+-q), q affine (x, y) in global memory, finite: mdbl-2008-s-1 with general a.
  // Three temporaries (P's own slots hold V, W and, briefly, the coefficient a).
  static G753_NI void mdbl_g(int P, const Fq* q, bool negq, int W) {
    const int X = P, Y = P + K, ZZ = P + 2 * K, ZZZ = P + 3 * K;
    const int t0 = W, t1 = W + K, t2 = W + 2 * K, tt = W + 3 * K;
    M::ldg(t0, q + K);
    if (negq) M::neg(t0, t0);  // y
    M::dbl(t1, t0);            // U = 2y
    if (M::is_zero(t1)) {
      set_inf(P);
      return;
    }
    M::sqr(ZZ, t1, tt);        // V
    M::mul(ZZZ, t1, ZZ, tt);   // W
    M::ldg(t1, q);             // x
    M::mul(t2, t1, ZZ, tt);    // S = x V
    M::sqr(t1, t1, tt);        // x^2
    M::dbl(Y, t1);
    M::add(t1, Y, t1);         // 3 x^2
    M::set_one(X);
    SC::mul_by_a(Y, X);        // a
    M::add(t1, t1, Y);         // M = 3 x^2 + a
    M::sqr(X, t1, tt);
    M::dbl(Y, t2);
    M::sub(X, X, Y);           // X3 = M^2 - 2S
    M::sub(t2, t2, X);
    M::mul(t2, t1, t2, tt);    // M (S - X3)
    M::mul(t1, ZZZ, t0, tt);   // W y
    M::sub(Y, t2, t1);
  }

  // P += (+-q), q affine in global memory, finite: madd-2008-s, 8M + 2S.
  // Extension-field curves, three temporaries: Q = X1 PP replaces X1 in place and Q - X3 is formed
  // as 3Q - (R^2 - PPP), so their accumulation kernels need 7K + NTMP slots per column.
  static G753_NI void madd_g(int P, const Fq* q, bool negq, int W) {
    const int X = P, Y = P + K, ZZ = P + 2 * K, ZZZ = P + 3 * K;
    const int t0 = W, t1 = W + K, t2 = W + 2 * K, tt = W + 3 * K;
    if (M::is_zero(ZZ)) {
      M::ldg(X, q);
      M::ldg(Y, q + K);
      if (negq) M::neg(Y, Y);
      M::set_one(ZZ);
      M::set_one(ZZZ);
      return;
    }
    M::ldg(t0, q);
    M::mul(t0, t0, ZZ, tt);
    M::sub(t0, t0, X);         // P = x2 ZZ1 - X1
    M::ldg(t1, q + K);
    M::mul(t1, t1, ZZZ, tt);
    if (negq) {
      M::add(t1, t1, Y);
      M::neg(t1, t1);          // R = -(y2 ZZZ1) - Y1
    } else {
      M::sub(t1, t1, Y);       // R = y2 ZZZ1 - Y1
    }
    if (M::is_zero(t0)) {
      if (M::is_zero(t1)) mdbl_g(P, q, negq, W);
      else set_inf(P);
      return;
    }
    if (K == 1) {
      // prime-field curves: one more temporary (the slot budget of two 128-thread blocks per SM
      // has room for it) saves the two additions of the 3Q form - measured 226 vs 230 ms at 2^22
      const int t3 = W + 3 * K;
      M::sqr(t2, t0, tt);      // PP
      M::mul(t0, t0, t2, tt);  // PPP
      M::mul(t3, X, t2, tt);   // Q = X1 PP
      M::mul(ZZ, ZZ, t2, tt);
      M::mul(ZZZ, ZZZ, t0, tt);
      M::sqr(X, t1, tt);
      M::sub(X, X, t0);
      M::dbl(t2, t3);
      M::sub(X, X, t2);        // X3 = R^2 - PPP - 2Q
      M::mul(t2, Y, t0, tt);   // Y1 PPP
      M::sub(t3, t3, X);
      M::mul(Y, t1, t3, tt);
      M::sub(Y, Y, t2);        // Y3 = R (Q - X3) - Y1 PPP
      return;
    }
    M::sqr(t2, t0, tt);        // PP
    M::mul(t0, t0, t2, tt);    // PPP
    M::mul(X, X, t2, tt);      // Q = X1 PP   (X1 is dead from here on)
    M::mul(ZZ, ZZ, t2, tt);    // ZZ3
    M::mul(ZZZ, ZZZ, t0, tt);  // ZZZ3
    M::sqr(t2, t1, tt);
    M::sub(t2, t2, t0);        // R^2 - PPP
    M::mul(Y, Y, t0, tt);      // Y1 PPP      (PPP is dead from here on)
    M::dbl(t0, X);
    M::add(t0, t0, X);
    M::sub(t0, t0, t2);        // 3Q - (R^2 - PPP) = Q - X3
    M::mul(t0, t1, t0, tt);    // R (Q - X3)
    M::sub(Y, t0, Y);          // Y3
    M::dbl(X, X);
    M::sub(X, t2, X);          // X3 = R^2 - PPP - 2Q
  }

  // ---- prime-field curves, six slots per column (P's four + two temporaries) --------------------
  // Same madd-2008-s / mdbl-2008-s-1 values as madd_g / mdbl_g; see s_madd6 (slots.cuh).
  static G753_NI void mdbl6_g(int P, const Fq* q, bool negq, int W) {
    const int X = P, Y = P + 1, ZZ = P + 2, ZZZ = P + 3, t0 = W, t1 = W + 1;
    M::ldg(t0, q + 1);
    if (negq) M::neg(t0, t0);  // y
    M::dbl(t1, t0);            // U = 2y
    if (M::is_zero(t1)) {
      set_inf(P);
      return;
    }
    M::sqr(ZZ, t1, 0);         // V
    M::mul(ZZZ, t1, ZZ, 0);    // W
    M::ldg(t1, q);             // x
    M::mul(Y, t1, ZZ, 0);      // S = x V
    M::sqr(t1, t1, 0);         // x^2
    M::dbl(X, t1);
    M::add(t1, X, t1);         // 3 x^2
    M::set_one(X);
    SC::mul_by_a(X, X);        // a (the prime-field forms read their operand before writing)
    M::add(t1, t1, X);         // M = 3 x^2 + a
    M::sqr(X, t1, 0);
    M::sub(X, X, Y);
    M::sub(X, X, Y);           // X3 = M^2 - 2S
    M::sub(Y, Y, X);
    M::mul(Y, t1, Y, 0);       // M (S - X3)
    M::mul(t1, ZZZ, t0, 0);    // W y
    M::sub(Y, Y, t1);
  }
  static G753_NI void madd6_g(int P, const Fq* q, bool negq, int W) {
    typedef typename M::T L;
    constexpr int F = M::FIELD;
    const int X = P, Y = P + 1, ZZ = P + 2, ZZZ = P + 3;
    if (M::is_zero(ZZ)) {
      M::ldg(X, q);
      M::ldg(Y, q + 1);
      if (negq) M::neg(Y, Y);
      M::set_one(ZZ);
      M::set_one(ZZZ);
      return;
    }
    const unsigned z = s_madd6<F, L>(P, W, q, negq, 0);   // P = x2 ZZ1 - X1, R = +-(y2 ZZZ1) - Y1
    if (z & 1u) {
      if (z & 2u) mdbl6_g(P, q, negq, W);
      else set_inf(P);
      return;
    }
    s_madd6<F, L>(P, W, q, negq, 1);
  }
  // the accumulation kernel's mixed addition: needs ACC_SCRATCH slots at W
  static G753_D void madd_acc_g(int P, const Fq* q, bool negq, int W) {
    if constexpr (K == 1 && G753_ACC6) madd6_g(P, q, negq, W);
    else madd_g(P, q, negq, W);
  }

  // P = 2P: dbl-2008-s-1.  Four temporaries: V's slot is recycled once ZZ3 = V ZZ1 is formed.
  static G753_NI void dbl(int P, int W) {
    const int X = P, Y = P + K, ZZ = P + 2 * K, ZZZ = P + 3 * K;
    const int t0 = W, t1 = W + K, t2 = W + 2 * K, t3 = W + 3 * K, tt = W + 4 * K;
    if (M::is_zero(ZZ)) return;
    M::dbl(t0, Y);             // U = 2 Y1
    if (M::is_zero(t0)) {
      set_inf(P);
      return;
    }
    M::sqr(t1, t0, tt);        // V
    M::mul(t0, t0, t1, tt);    // W
    M::mul(t2, X, t1, tt);     // S = X1 V
    M::sqr(t3, ZZ, tt);        // ZZ1^2
    M::mul(ZZ, ZZ, t1, tt);    // ZZ3 = V ZZ1
    SC::mul_by_a(t1, t3);      // a ZZ1^2   (t1 != t3: the Fq3 twist permutes coordinates)
    M::sqr(t3, X, tt);
    M::add(t1, t1, t3);
    M::dbl(t3, t3);
    M::add(t3, t3, t1);        // M = 3 X1^2 + a ZZ1^2
    M::mul(ZZZ, ZZZ, t0, tt);  // ZZZ3 = W ZZZ1
    M::sqr(X, t3, tt);
    M::dbl(t1, t2);
    M::sub(X, X, t1);          // X3 = M^2 - 2S
    M::sub(t2, t2, X);
    M::mul(t2, t3, t2, tt);    // M (S - X3)
    M::mul(t1, t0, Y, tt);     // W Y1
    M::sub(Y, t2, t1);
  }

  // P += Q (both on slots, Q preserved): add-2008-s, 12M + 2S.  Four temporaries: once U1, S1, P, R
  // exist, X1 and Y1 are dead and their slots carry PP and the later intermediates.
  static G753_NI void add(int P, int Q, int W) {
    const int X = P, Y = P + K, ZZ = P + 2 * K, ZZZ = P + 3 * K;
    const int t0 = W, t1 = W + K, t2 = W + 2 * K, t3 = W + 3 * K, tt = W + 4 * K;
    if (M::is_zero(Q + 2 * K)) return;
    if (M::is_zero(ZZ)) {
      copy(P, Q);
      return;
    }
    M::mul(t0, X, Q + 2 * K, tt);   // U1 = X1 ZZ2
    M::mul(t1, Y, Q + 3 * K, tt);   // S1 = Y1 ZZZ2
    M::mul(t2, Q, ZZ, tt);
    M::sub(t2, t2, t0);             // P = X2 ZZ1 - U1
    M::mul(t3, Q + K, ZZZ, tt);
    M::sub(t3, t3, t1);             // R = Y2 ZZZ1 - S1
    if (M::is_zero(t2)) {
      if (M::is_zero(t3)) dbl(P, W);
      else set_inf(P);
      return;
    }
    M::mul(ZZ, ZZ, Q + 2 * K, tt);
    M::mul(ZZZ, ZZZ, Q + 3 * K, tt);
    M::sqr(X, t2, tt);              // PP
    M::mul(t2, t2, X, tt);          // PPP
    M::mul(t0, t0, X, tt);          // Q = U1 PP
    M::mul(ZZ, ZZ, X, tt);
    M::mul(ZZZ, ZZZ, t2, tt);
    M::sqr(X, t3, tt);
    M::sub(X, X, t2);
    M::dbl(Y, t0);
    M::sub(X, X, Y);                // X3 = R^2 - PPP - 2Q
    M::sub(t0, t0, X);
    M::mul(t0, t3, t0, tt);         // R (Q - X3)
    M::mul(Y, t1, t2, tt);          // S1 PPP
    M::sub(Y, t0, Y);
  }

  // P += q, q an XYZZ point in global memory (4K Fq, order X, Y, ZZ, ZZZ); q's coordinates are
  // (re)loaded where they are used
  static G753_NI void add_g(int P, const Fq* q, int W) {
    const int X = P, Y = P + K, ZZ = P + 2 * K, ZZZ = P + 3 * K;
    const int t0 = W, t1 = W + K, t2 = W + 2 * K, t3 = W + 3 * K, tt = W + 4 * K;
    M::ldg(t0, q + 2 * K);          // ZZ2
    if (M::is_zero(t0)) return;
    if (M::is_zero(ZZ)) {
      ldg(P, q);
      return;
    }
    M::mul(t0, X, t0, tt);          // U1 = X1 ZZ2
    M::ldg(t1, q + 3 * K);
    M::mul(t1, Y, t1, tt);          // S1 = Y1 ZZZ2
    M::ldg(t2, q);
    M::mul(t2, t2, ZZ, tt);
    M::sub(t2, t2, t0);             // P = X2 ZZ1 - U1
    M::ldg(t3, q + K);
    M::mul(t3, t3, ZZZ, tt);
    M::sub(t3, t3, t1);             // R = Y2 ZZZ1 - S1
    if (M::is_zero(t2)) {
      if (M::is_zero(t3)) dbl(P, W);
      else set_inf(P);
      return;
    }
    M::ldg(X, q + 2 * K);
    M::mul(ZZ, ZZ, X, tt);          // ZZ1 ZZ2
    M::ldg(X, q + 3 * K);
    M::mul(ZZZ, ZZZ, X, tt);        // ZZZ1 ZZZ2
    M::sqr(X, t2, tt);              // PP
    M::mul(t2, t2, X, tt);          // PPP
    M::mul(t0, t0, X, tt);          // Q = U1 PP
    M::mul(ZZ, ZZ, X, tt);
    M::mul(ZZZ, ZZZ, t2, tt);
    M::sqr(X, t3, tt);
    M::sub(X, X, t2);
    M::dbl(Y, t0);
    M::sub(X, X, Y);
    M::sub(t0, t0, X);
    M::mul(t0, t3, t0, tt);
    M::mul(Y, t1, t2, tt);
    M::sub(Y, t0, Y);
  }

  // XYZZ -> the reference's homogeneous projective (X:Y:Z) = (X ZZZ : Y ZZ : ZZ ZZZ), in place
  // over the first three elements of P (infinity -> (0 : 1 : 0), GroupProjective::zero())
  static G753_D void to_projective(int P, int W) {
    const int X = P, Y = P + K, ZZ = P + 2 * K, ZZZ = P + 3 * K;
    if (M::is_zero(ZZ)) {
      M::set_zero(X);
      M::set_one(Y);
      return;
    }
    M::mul(X, X, ZZZ, W);
    M::mul(Y, Y, ZZ, W);
    M::mul(ZZ, ZZ, ZZZ, W);
  }
  // XYZZ -> affine (x, y) in the X, Y elements (ZZ, ZZZ become 1); P finite.  2K + NTMP scratch.
  static G753_D void to_affine(int P, int W) {
    const int X = P, Y = P + K, ZZ = P + 2 * K, ZZZ = P + 3 * K;
    const int t0 = W, t1 = W + K, tt = W + 2 * K;
    M::mul(t0, ZZ, ZZZ, tt);
    M::inv(t0, t0, tt);
    M::mul(t1, t0, ZZZ, tt);  // 1 / ZZ
    M::mul(X, X, t1, tt);
    M::mul(t1, t0, ZZ, tt);   // 1 / ZZZ
    M::mul(Y, Y, t1, tt);
    M::set_one(ZZ);
    M::set_one(ZZZ);
  }
  // homogeneous projective (X:Y:Z) in global memory -> XYZZ on slots
  static G753_D void from_projective_g(int P, const Fq* g, int W) {
    const int X = P, Y = P + K, ZZ = P + 2 * K, ZZZ = P + 3 * K;
    M::ldg(ZZZ, g + 2 * K);  // Z
    if (M::is_zero(ZZZ)) {
      set_inf(P);
      return;
    }
    M::sqr(ZZ, ZZZ, W);       // Z^2
    M::ldg(X, g);
    M::mul(X, X, ZZZ, W);     // X Z
    M::ldg(Y, g + K);
    M::mul(Y, Y, ZZ, W);      // Y Z^2
    M::mul(ZZZ, ZZZ, ZZ, W);  // Z^3
  }
};

}  // namespace g753
