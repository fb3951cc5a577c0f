// Host-emulation build of the device limb arithmetic (TEST ONLY - never linked into the
// product library).  The PTX carry-chain primitives of csrc/fq.cuh are replaced by their
// thread-local-carry emulation and the shared-memory slots of csrc/slots.cuh by a static
// buffer, so the limb logic, the tower formulas and the group law can be checked against the
// oracle without a GPU.
#define G753_HOST_EMUL 1
#include "../../ginger-lib_b200/csrc/ec_slots.cuh"
#include <cstring>

using namespace g753;
typedef Lay<32, 1> T;  // emulated block: 32 columns, one lane each (thread 0 is the one that runs)

template <int FID>
static void field_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  Fq x, y, r;
  memcpy(x.l, a, 96);
  memcpy(y.l, b, 96);
  switch (op) {
    case 0: r = fq_mul<FID>(x, y); break;
    case 1: r = fq_add<FID>(x, y); break;
    case 2: r = fq_sub<FID>(x, y); break;
    case 3: r = fq_sqr<FID>(x); break;
    case 4: r = fq_neg<FID>(x); break;
    case 5: r = fq_inv<FID>(x); break;
    case 6: r = fq_to_mont<FID>(x); break;
    case 7: r = fq_from_mont<FID>(x); break;
    case 8: r = fq_mul_small<FID, 11>(x); break;
    case 9: r = fq_mul_small<FID, 13>(x); break;
    case 10: r = fq_mul_small<FID, 26>(x); break;
    case 11: r = fq_mul_small<FID, 121>(x); break;
    case 12: r = fq_dbl<FID>(x); break;
    case 13:  // the product / squaring as the slot kernels run them (rolled multiplier under -DG753_ROLLED=1)
    case 14:
      threadIdx.x = 0;
      s_st<T>(0, x);
      s_st<T>(1, y);
      if (op == 13) s_mul<FID, T>(1, 0, 1);  // d aliases b
      else s_sqr<FID, T>(1, 0);
      r = s_ld<T>(1);
      break;
    default: r = fq_zero<FID>();
  }
  memcpy(out, r.l, 96);
}

static void put(int slot, const uint32_t* src, int count) {
  for (int i = 0; i < count; i++) {
    Fq v;
    memcpy(v.l, src + i * NL, 96);
    s_st<T>(slot + i, v);
  }
}
static void get(uint32_t* dst, int slot, int count) {
  for (int i = 0; i < count; i++) {
    Fq v = s_ld<T>(slot + i);
    memcpy(dst + i * NL, v.l, 96);
  }
}

// tower ops on slots; `alias` makes the destination alias the first operand (the in-place use
// the curve formulas rely on)
template <class M>
static void ext_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  const int K = M::K, A = 0, B = K, D = 2 * K, TMP = 3 * K;
  threadIdx.x = 0;
  put(A, a, K);
  put(B, b, K);
  int res = D;
  switch (op) {
    case 0: M::mul(D, A, B, TMP); break;
    case 1: M::add(D, A, B); break;
    case 2: M::sub(D, A, B); break;
    case 3: M::sqr(D, A, TMP); break;
    case 4: M::neg(D, A); break;
    case 12: M::dbl(D, A); break;
    case 20: M::mul(A, A, B, TMP); res = A; break;  // d aliases a
    case 21: M::mul(B, A, B, TMP); res = B; break;  // d aliases b
    case 23: M::sqr(A, A, TMP); res = A; break;
    default: M::set_zero(D);
  }
  get(out, res, K);
}

// curve ops on XYZZ accumulators: op 0 = madd(acc, affine), 1 = add(acc, acc2), 2 = dbl(acc),
// 3 = to homogeneous projective (X, Y, Z), 5 = add_g(acc, acc2 in "global" memory),
// 6 = madd(acc, -affine)
template <class SC>
static void curve_op(int op, const uint32_t* acc_in, const uint32_t* other, uint32_t* out) {
  typedef EcS<SC> E;
  const int K = E::K, P = 0, Q = E::PT, S = 2 * E::PT;
  threadIdx.x = 0;
  put(P, acc_in, 4 * K);
  static Fq gbuf[16];
  if (op == 0 || op == 6) {
    memcpy(gbuf, other, 96 * 2 * K);
    E::madd_g(P, gbuf, op == 6, S);
    get(out, P, 4 * K);
  } else if (op == 7 || op == 8) {  // the accumulation kernel's form (six slots on the prime-field curves)
    memcpy(gbuf, other, 96 * 2 * K);
    if constexpr (E::K == 1) E::madd6_g(P, gbuf, op == 8, S);  // six-slot form (opt-in on the device: G753_ACC6)
    else E::madd_acc_g(P, gbuf, op == 8, S);
    get(out, P, 4 * K);
  } else if (op == 1) {
    put(Q, other, 4 * K);
    E::add(P, Q, S);
    get(out, P, 4 * K);
  } else if (op == 5) {
    memcpy(gbuf, other, 96 * 4 * K);
    E::add_g(P, gbuf, S);
    get(out, P, 4 * K);
  } else if (op == 2) {
    E::dbl(P, S);
    get(out, P, 4 * K);
  } else if (op == 3) {
    E::to_projective(P, S);
    get(out, P, 3 * K);
  }
}

extern "C" {
void emul_field_op(int fid, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  if (fid == 0) field_op<0>(op, a, b, out);
  else field_op<1>(op, a, b, out);
}
// (a b + c d) / R under one reduction: fq_mul2 (registers) and s_mul2 (multipliers streamed from slots,
// c optionally negated / times 13)
void emul_field_mul2(int fid, int mode, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d,
                     uint32_t* out) {
  threadIdx.x = 0;
  Fq x[4], r;
  memcpy(x[0].l, a, 96);
  memcpy(x[1].l, b, 96);
  memcpy(x[2].l, c, 96);
  memcpy(x[3].l, d, 96);
  if (mode < 0) {
    r = fid == 0 ? fq_mul2<0>(x[0], x[1], x[2], x[3]) : fq_mul2<1>(x[0], x[1], x[2], x[3]);
  } else {
    for (int i = 0; i < 4; i++) s_st<T>(i, x[i]);
    if (fid == 0) s_mul2<0, T, 13>(0, 0, 1, 2, 3, mode);  // d aliases a
    else s_mul2<1, T, 13>(0, 0, 1, 2, 3, mode);
    r = s_ld<T>(0);
  }
  memcpy(out, r.l, 96);
}
// ext 2 = Fq2 over field 0 (MNT4 G2 base field); ext 3 = Fq3 over field 1 (MNT6 G2 base field)
void emul_ext_op(int ext, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  if (ext == 2) ext_op<Tw2<0, T, 13>>(op, a, b, out);
  else ext_op<Tw3<1, T, 11>>(op, a, b, out);
}
// curve ids as in include/g753.h: 0 = MNT4 G1, 1 = MNT4 G2, 2 = MNT6 G1, 3 = MNT6 G2
void emul_curve_op(int curve, int op, const uint32_t* acc, const uint32_t* other, uint32_t* out) {
  switch (curve) {
    case 0: curve_op<SCurveM4G1<T>>(op, acc, other, out); break;
    case 1: curve_op<SCurveM4G2<T>>(op, acc, other, out); break;
    case 2: curve_op<SCurveM6G1<T>>(op, acc, other, out); break;
    case 3: curve_op<SCurveM6G2<T>>(op, acc, other, out); break;
  }
}
}
