// Host-emulation build of the device limb arithmetic (TEST ONLY - never linked into the
// product library).  The PTX carry-chain primitives of csrc/fq.cuh are replaced by their
// thread-local-carry emulation so the limb logic can be checked against the oracle
// without a GPU.
#include "../../ginger-lib_b200/csrc/fq.cuh"
#include "../../ginger-lib_b200/csrc/fqk.cuh"
#include "../../ginger-lib_b200/csrc/ec.cuh"
#include <cstring>

using namespace g753;

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
    default: r = fq_zero<FID>();
  }
  memcpy(out, r.l, 96);
}

template <class F>
static void ext_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  F x, y, r;
  memcpy(&x, a, sizeof(F));
  memcpy(&y, b, sizeof(F));
  switch (op) {
    case 0: r = F::mul(x, y); break;
    case 1: r = F::add(x, y); break;
    case 2: r = F::sub(x, y); break;
    case 3: r = F::sqr(x); break;
    case 4: r = F::neg(x); break;
    case 5: r = F::inv(x); break;
    case 12: r = F::dbl(x); break;
    default: r = F::zero();
  }
  memcpy(out, &r, sizeof(F));
}

// curve ops on XYZZ accumulators: op 0 = madd(acc, affine), 1 = add(acc, acc2), 2 = dbl(acc),
// 3 = to homogeneous projective (X, Y, Z)
template <class C>
static void curve_op(int op, const uint32_t* acc_in, const uint32_t* other, uint32_t* out) {
  typedef typename C::F F;
  Xyzz<C> p;
  memcpy(&p, acc_in, sizeof(p));
  if (op == 0) {
    Affine<C> q;
    memcpy(&q, other, sizeof(q));
    xyzz_madd<C>(p, q);
    memcpy(out, &p, sizeof(p));
  } else if (op == 1) {
    Xyzz<C> q;
    memcpy(&q, other, sizeof(q));
    xyzz_add<C>(p, q);
    memcpy(out, &p, sizeof(p));
  } else if (op == 2) {
    xyzz_dbl<C>(p);
    memcpy(out, &p, sizeof(p));
  } else if (op == 3) {
    F X, Y, Z;
    xyzz_to_projective<C>(p, X, Y, Z);
    memcpy(out, &X, sizeof(F));
    memcpy((char*)out + sizeof(F), &Y, sizeof(F));
    memcpy((char*)out + 2 * sizeof(F), &Z, sizeof(F));
  } else if (op == 4) {  // scalar multiplication by a 768-bit scalar in `other` of an affine point in acc_in.x/.y
    Affine<C> q;
    memcpy(&q, acc_in, sizeof(q));
    Xyzz<C> r = xyzz_scalar_mul<C>(q, other);
    memcpy(out, &r, sizeof(r));
  }
}

extern "C" {
void emul_field_op(int fid, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  if (fid == 0) field_op<0>(op, a, b, out);
  else field_op<1>(op, a, b, out);
}
// ext 2 = Fq2 over field 0 (MNT4 G2 base field); ext 3 = Fq3 over field 1 (MNT6 G2 base field)
void emul_ext_op(int ext, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  if (ext == 2) ext_op<Fq2M4>(op, a, b, out);
  else ext_op<Fq3M6>(op, a, b, out);
}
// curve ids as in include/g753.h: 0 = MNT4 G1, 1 = MNT4 G2, 2 = MNT6 G1, 3 = MNT6 G2
void emul_curve_op(int curve, int op, const uint32_t* acc, const uint32_t* other, uint32_t* out) {
  switch (curve) {
    case 0: curve_op<CurveM4G1>(op, acc, other, out); break;
    case 1: curve_op<CurveM4G2>(op, acc, other, out); break;
    case 2: curve_op<CurveM6G1>(op, acc, other, out); break;
    case 3: curve_op<CurveM6G2>(op, acc, other, out); break;
  }
}
}
