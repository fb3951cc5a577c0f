// Four-step NTT sharded over the GPUs of one box: n = n1 * n2, rank g owns n2 / G
// columns of the n1 x n2 view of the input.
//
//   X[k1 + n1 k2] = sum_{i2} omega^(i2 k1) (omega^n1)^(i2 k2) sum_{i1} x[i1 n2 + i2] (omega^n2)^(i1 k1)
//
//   step 1 (local):   for every owned column i2, the length-n1 transform over i1, outputs scaled by
//                     the twiddle omega^(i2 k1) (fused into the last butterfly pass), then packed
//                     per destination rank;
//   exchange:         all-to-all of (n / G^2)-element blocks - the only collective, done by the
//                     caller (torch.distributed all_to_all_single over NCCL / NVLink);
//   step 2 (local):   unpack to rows, for every owned row k1 the length-n2 transform over i2.
// Layouts (rank g, cols = n2 / G, rows = n1 / G):
//   input  local[i2l][i1]  = x[i1 n2 + g cols + i2l]
//   output local[k1l][k2]  = X[(g rows + k1l) + n1 k2]
// so for n1 == n2 an output shard is a valid input shard of the next transform: the seven
// transforms of the witness map chain without any data movement besides the exchange itself.
// The reference has no counterpart (its parallel_fft, domain.rs:360-416, is the same
// decomposition over CPU threads); the contract is "sharded result == single-GPU result".
#pragma once
#include "ntt.cuh"

namespace g753 {

// out[r][j] = S_r * B_r^j,  S_r = c0 * cs^(row0 + r),  B_r = x0 * xs^(row0 + r);  k = {x0, xs, c0, cs}
template <int FID>
G753_D Fq fq_pow_u64(const Fq& a, uint64_t e) {
  Fq r = fq_one<FID>();
  bool started = false;
  for (int i = 63; i >= 0; i--) {
    if (started) r = fq_sqr<FID>(r);
    if ((e >> i) & 1) {
      r = started ? fq_mul<FID>(r, a) : a;
      started = true;
    }
  }
  return r;
}
template <int FID>
__global__ void k_pow_table(Fq* __restrict__ out, unsigned rows, unsigned len, uint64_t row0, const Fq* __restrict__ k0,
                            const Fq* __restrict__ k1, const Fq* __restrict__ k2, const Fq* __restrict__ k3) {
  unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const Fq one = fq_one<FID>();
  Fq x0 = k0 ? *k0 : one, xs = k1 ? *k1 : one, c0 = k2 ? *k2 : one, cs = k3 ? *k3 : one;
  Fq base = fq_mul<FID>(x0, fq_pow_u64<FID>(xs, row0 + r));
  Fq v = fq_mul<FID>(c0, fq_pow_u64<FID>(cs, row0 + r));
  for (unsigned j = 0; j < len; j++) {
    out[(size_t)r * len + j] = v;
    v = fq_mul<FID>(v, base);
  }
}

// out[a sA + b sB + c sC] = in[(a B + b) C + c] over an A x B x C array of field elements
static __global__ void k_permute3(const Fq* __restrict__ in, Fq* __restrict__ out, unsigned A, unsigned B, unsigned C,
                                  size_t sA, size_t sB, size_t sC) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)A * B * C) return;
  unsigned c = (unsigned)(t % C), b = (unsigned)((t / C) % B), a = (unsigned)(t / ((size_t)B * C));
  out[a * sA + b * sB + c * sC] = in[t];
}

}  // namespace g753

using namespace g753;

struct g753_ntt_shard {
  int field = 0;
  unsigned log_n = 0, log_n1 = 0, log_n2 = 0, world = 1, rank = 0;
  size_t n1 = 1, n2 = 1, cols = 1, rows = 1;  // cols = n2 / world, rows = n1 / world
  Fq* t1f = nullptr;   // [cols][n1]  omega^(i2 k1)
  Fq* t1i = nullptr;   // [cols][n1]  omega^(-i2 k1) / n
  Fq* cp = nullptr;    // [cols][n1]  g^(i1 n2 + i2)
  Fq* cq = nullptr;    // [rows][n2]  g^-(k1 + n1 k2)
  Fq* consts = nullptr;
  void release() {
    dev_free(t1f);
    dev_free(t1i);
    dev_free(cp);
    dev_free(cq);
    dev_free(consts);
    t1f = t1i = cp = cq = consts = nullptr;
  }
};
