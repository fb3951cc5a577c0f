// Radix-2 number-theoretic transforms behind EvaluationDomain.
//
// Replaces (reference, relative to /root/reference/algebra/src/fft/domain.rs):
//   :65-94    EvaluationDomain::new          -> ntt_tables_build (omega, omega^-1, n^-1, 17^+-i)
//   :120-123  fft_in_place   -> best_fft :305-313 -> serial_fft :315-358 / parallel_fft :360-416
//   :134-138  ifft_in_place  (omega^-1, then * size_inv)
//   :140-152  distribute_powers, :163-166 coset_fft_in_place, :176-179 coset_ifft_in_place
//   :245-256, :289-302  element-wise helpers used by R1CStoQAP::witness_map
// Output convention is the reference's: natural order, out[i] = sum_j in[j] * omega^(i j),
// omega = ROOT_OF_UNITY^(2^(TWO_ADICITY - log_n)).
//
// Algorithm: Stockham autosort (natural order in and out, no bit-reversal pass), twiddles
// omega^j (j < n/2) and coset powers 17^i / 17^-i * n^-1 from tables that are built once per
// (field, log_n) and stay resident in HBM.  The coset scaling and the 1/n factor are fused
// into the first / last butterfly pass.
#pragma once
#include "device.cuh"
#include "fq.cuh"

namespace g753 {

constexpr unsigned NTT_MAX_LOG = 29;

struct NttTables {
  unsigned log_n = 0;
  Fq* tw_fwd = nullptr;     // omega^j,        j < max(n/2, 1)
  Fq* tw_inv = nullptr;     // omega^-j
  Fq* coset = nullptr;      // g^i,            i < n       (g = 17)
  Fq* coset_inv = nullptr;  // g^-i * n^-1
  Fq* consts = nullptr;     // [0] = n^-1, the four squaring chains, then (g^n - 1)^-1
  void release() {
    dev_free(tw_fwd);
    dev_free(tw_inv);
    dev_free(coset);
    dev_free(coset_inv);
    dev_free(consts);
    tw_fwd = tw_inv = coset = coset_inv = consts = nullptr;
  }
};

// consts layout: [0] n^-1 ; [1 + l] omega^(2^l) ; [1 + L + l] omega^-(2^l) ; [1 + 2L + l] g^(2^l) ;
// [1 + 3L + l] g^-(2^l), L = max(log_n, 1), l < L ; [1 + 4L] (g^n - 1)^-1, the inverse of the
// vanishing polynomial on the coset (divide_by_vanishing_poly_on_coset_in_place, domain.rs:245-256)
template <int FID>
__global__ void k_ntt_setup(unsigned log_n, Fq* consts) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const unsigned L = log_n ? log_n : 1;
  Fq w;
#pragma unroll
  for (int i = 0; i < NL; i++) w.l[i] = G753_FC(FID).root[i];
  for (unsigned k = log_n; k < G753_FC(FID).two_adicity; k++) w = fq_sqr<FID>(w);  // domain.rs:76-79
  Fq wi = fq_inv<FID>(w);
  Fq g, gi;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    g.l[i] = G753_FC(FID).gen[i];
    gi.l[i] = G753_FC(FID).gen_inv[i];
  }
  // n^-1 = (2^-1)^log_n, with 2^-1 = (p + 1) / 2
  Fq two = fq_dbl<FID>(fq_one<FID>());
  Fq half = fq_inv<FID>(two);
  Fq ninv = fq_one<FID>();
  for (unsigned k = 0; k < log_n; k++) ninv = fq_mul<FID>(ninv, half);
  consts[0] = ninv;
  for (unsigned l = 0; l < L; l++) {
    consts[1 + l] = w;
    consts[1 + L + l] = wi;
    consts[1 + 2 * L + l] = g;
    consts[1 + 3 * L + l] = gi;
    w = fq_sqr<FID>(w);
    wi = fq_sqr<FID>(wi);
    g = fq_sqr<FID>(g);
    gi = fq_sqr<FID>(gi);
  }
  // after the loop g = 17^(2^L); for log_n = 0 the domain has one element and g^n = 17
  Fq gn = log_n ? g : consts[1 + 2 * L];
  consts[1 + 4 * L] = fq_inv<FID>(fq_sub<FID>(gn, fq_one<FID>()));
}

// tab[0] = *first (or one)
template <int FID>
__global__ void k_table_init(Fq* tab, const Fq* first) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  tab[0] = first ? *first : fq_one<FID>();
}
// out[l] = x^(2^l), l < count
template <int FID>
__global__ void k_square_chain(const Fq* x, Fq* out, unsigned count) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Fq v = *x;
  for (unsigned l = 0; l < count; l++) {
    out[l] = v;
    v = fq_sqr<FID>(v);
  }
}
// tab[len + t] = tab[t] * (*step) for t < len  (doubles the table: step = base^len)
template <int FID>
__global__ void k_table_double(Fq* tab, const Fq* step, unsigned len, unsigned cap) {
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= len || len + t >= cap) return;
  tab[len + t] = fq_mul<FID>(tab[t], *step);
}

// the constants of a size-2^log_n domain alone (no O(n) tables): consts layout as k_ntt_setup's
template <int FID>
static int ntt_consts_build(Fq** consts, unsigned log_n, cudaStream_t stream, uint64_t* launches) {
  const unsigned L = log_n ? log_n : 1;
  G753_TRY(dev_alloc((void**)consts, sizeof(Fq) * (2 + 4 * L)));
  G753_LAUNCH(k_ntt_setup<FID>, 1, 1, stream, log_n, *consts);
  if (launches) ++*launches;
  return launch_check("k_ntt_setup");
}

template <int FID>
static int ntt_tables_build(NttTables& T, unsigned log_n, cudaStream_t stream, uint64_t* launches) {
  T.log_n = log_n;
  const size_t n = (size_t)1 << log_n;
  const size_t half = n > 1 ? n / 2 : 1;
  const unsigned L = log_n ? log_n : 1;
  G753_TRY(dev_alloc((void**)&T.consts, sizeof(Fq) * (2 + 4 * L)));
  G753_TRY(dev_alloc((void**)&T.tw_fwd, sizeof(Fq) * half));
  G753_TRY(dev_alloc((void**)&T.tw_inv, sizeof(Fq) * half));
  G753_TRY(dev_alloc((void**)&T.coset, sizeof(Fq) * n));
  G753_TRY(dev_alloc((void**)&T.coset_inv, sizeof(Fq) * n));
  G753_LAUNCH(k_ntt_setup<FID>, 1, 1, stream, log_n, T.consts);
  G753_LAUNCH(k_table_init<FID>, 1, 1, stream, T.tw_fwd, (const Fq*)nullptr);
  G753_LAUNCH(k_table_init<FID>, 1, 1, stream, T.tw_inv, (const Fq*)nullptr);
  G753_LAUNCH(k_table_init<FID>, 1, 1, stream, T.coset, (const Fq*)nullptr);
  G753_LAUNCH(k_table_init<FID>, 1, 1, stream, T.coset_inv, (const Fq*)T.consts);
  if (launches) *launches += 5;
  for (unsigned l = 0; l < log_n; l++) {
    unsigned len = 1u << l;
    if (len < half) {
      G753_LAUNCH(k_table_double<FID>, div_up(len, 128), 128, stream, T.tw_fwd, T.consts + 1 + l, len,
                  (unsigned)half);
      G753_LAUNCH(k_table_double<FID>, div_up(len, 128), 128, stream, T.tw_inv, T.consts + 1 + L + l, len,
                  (unsigned)half);
      if (launches) *launches += 2;
    }
    G753_LAUNCH(k_table_double<FID>, div_up(len, 128), 128, stream, T.coset, T.consts + 1 + 2 * L + l, len,
                (unsigned)n);
    G753_LAUNCH(k_table_double<FID>, div_up(len, 128), 128, stream, T.coset_inv, T.consts + 1 + 3 * L + l, len,
                (unsigned)n);
    if (launches) *launches += 2;
  }
  return launch_check("ntt_tables_build");
}

// Stockham radix-2 stage s (sub-transform length Ns = 2^s -> 2^(s+1)) on the whole vector:
//   out[j0], out[j0 + Ns] = in[j] +- in[j + n/2] * omega^(k n / (2 Ns)),  k = j mod Ns,
//   j0 = (j div Ns) * 2 Ns + k.
// In index bits (n = 2^L): the input index is [h | u | k] (h = top bit, k = low s bits) and the
// output index [u | b | k]: every stage consumes the top bit and re-inserts the butterflied bit at
// position s.  q consecutive stages s .. s+q-1 therefore close over the 2^q elements
//   in:  (t << (L - q)) | m,                   t < 2^q,   column m = (u' << s) | k  fixed
//   out: (u' << (s + q)) | (P << s) | k,       P < 2^q,
// which is what one pass of k_ntt_pass keeps in shared memory: q stages for one trip through HBM.
// Local stage d of a pass is a Stockham stage of the 2^q-point column with the GLOBAL twiddle
//   omega^(((kl << s) | k) * 2^(L-1-s-d)),  kl = low d bits of the local butterfly index.
// pre  (first pass): inputs are multiplied by pre[index]            (coset_fft: g^i)
// post (last pass):  outputs are multiplied by post[index] or *post_const (ifft: n^-1,
//                    coset_ifft: g^-i n^-1)
// Fused exchange of the sharded four-step transform: the LAST pass of the column transforms stores
// every output straight into its final place on the destination GPU (peer memory mapped over
// NVLink), so the transpose + all-to-all costs no kernel and no collective call of its own.
// Output k1 of column bi goes to rank h = k1 / rows, element (k1 % rows) * n2 + rank * cols + bi.
struct NttScatter {
  Fq* peer[8];
  unsigned rows, cols, rank, enabled;
  size_t n2;
};

constexpr int NTT_T = 128;            // threads per block = butterflies per stage per block
constexpr int NTT_E = 2 * NTT_T;      // elements per block (columns x 2^q)
constexpr unsigned NTT_MAX_Q = 8;     // 2^q <= NTT_E
constexpr size_t NTT_SMEM = 2 * (size_t)NTT_E * sizeof(Fq);  // ping-pong buffers, 48 KB

#if !defined(G753_HOST_EMUL)
// shared-memory layout: buffer b, 16-byte chunk c, element e -> sm[(b * 6 + c) * NTT_E + e]
G753_D Fq ntt_lds(const uint4* sm, unsigned buf, unsigned e) {
  const uint4* p = sm + buf * (6 * NTT_E) + e;
  Fq r;
#pragma unroll
  for (int c = 0; c < 6; c++) {
    uint4 v = p[c * NTT_E];
    r.l[4 * c] = v.x;
    r.l[4 * c + 1] = v.y;
    r.l[4 * c + 2] = v.z;
    r.l[4 * c + 3] = v.w;
  }
  return r;
}
G753_D void ntt_sts(uint4* sm, unsigned buf, unsigned e, const Fq& a) {
  uint4* p = sm + buf * (6 * NTT_E) + e;
#pragma unroll
  for (int c = 0; c < 6; c++) {
    uint4 v;
    v.x = a.l[4 * c];
    v.y = a.l[4 * c + 1];
    v.z = a.l[4 * c + 2];
    v.w = a.l[4 * c + 3];
    p[c * NTT_E] = v;
  }
}
G753_D Fq ntt_ldg(const Fq* g) {
  const uint4* p = (const uint4*)g;
  Fq r;
#pragma unroll
  for (int c = 0; c < 6; c++) {
    uint4 v = p[c];
    r.l[4 * c] = v.x;
    r.l[4 * c + 1] = v.y;
    r.l[4 * c + 2] = v.z;
    r.l[4 * c + 3] = v.w;
  }
  return r;
}
G753_D void ntt_stg(Fq* g, const Fq& a) {
  uint4* p = (uint4*)g;
#pragma unroll
  for (int c = 0; c < 6; c++) {
    uint4 v;
    v.x = a.l[4 * c];
    v.y = a.l[4 * c + 1];
    v.z = a.l[4 * c + 2];
    v.w = a.l[4 * c + 3];
    p[c] = v;
  }
}

template <int FID>
__global__ void __launch_bounds__(NTT_T, 4)
k_ntt_pass(const Fq* __restrict__ in, Fq* __restrict__ out, const Fq* __restrict__ tw, unsigned L, unsigned s,
           unsigned q, const Fq* __restrict__ pre, size_t pre_stride, const Fq* __restrict__ post,
           size_t post_stride, const Fq* __restrict__ post_const, const NttScatter sc) {
  extern __shared__ uint4 ntt_sm[];
  const unsigned tid = threadIdx.x;
  {  // blockIdx.y = index of the vector within a batch of equal-length transforms
    const size_t bi = blockIdx.y;
    in += bi << L;
    out += bi << L;
    if (pre != nullptr) pre += bi * pre_stride;
    if (post != nullptr) post += bi * post_stride;
  }
  const unsigned half_q = 1u << (q - 1);
  const unsigned cidx = tid >> (q - 1);           // column within the block
  const unsigned jl = tid & (half_q - 1);         // local butterfly index
  const unsigned cols_per_block = NTT_E >> q;
  const size_t n_cols = (size_t)1 << (L - q);
  const size_t m = (size_t)blockIdx.x * cols_per_block + cidx;
  const bool active = m < n_cols;
  const size_t k = m & (((size_t)1 << s) - 1), u = m >> s;
  const unsigned ebase = cidx << q;
  const size_t i0 = ((size_t)jl << (L - q)) | m, i1 = i0 + ((size_t)1 << (L - 1));
  for (unsigned d = 0; d < q; d++) {
    const unsigned kl = jl & ((1u << d) - 1);
    const unsigned jl0 = ((jl >> d) << (d + 1)) | kl;
    if (active) {
      // b (and its twiddle product) first, a afterwards: keeps the multiplier's live set small
      Fq b;
      if (d == 0) {
        b = ntt_ldg(in + i1);
        if (pre != nullptr) b = fq_mul<FID>(b, ntt_ldg(pre + i1));
      } else {
        b = ntt_lds(ntt_sm, (d - 1) & 1, ebase + jl + half_q);
      }
      if (s + d != 0) {  // global stage 0: every twiddle is omega^0 = 1
        const size_t ti = (((size_t)kl << s) | k) << (L - 1 - s - d);
        b = fq_mul<FID>(b, ntt_ldg(tw + ti));
      }
      Fq a;
      if (d == 0) {
        a = ntt_ldg(in + i0);
        if (pre != nullptr) a = fq_mul<FID>(a, ntt_ldg(pre + i0));
      } else {
        a = ntt_lds(ntt_sm, (d - 1) & 1, ebase + jl);
      }
      Fq x = fq_add<FID>(a, b);
      Fq y = fq_sub<FID>(a, b);
      if (d + 1 < q) {
        ntt_sts(ntt_sm, d & 1, ebase + jl0, x);
        ntt_sts(ntt_sm, d & 1, ebase + jl0 + (1u << d), y);
      } else {
        const size_t o0 = (u << (s + q)) | ((size_t)jl0 << s) | k, o1 = o0 + ((size_t)1 << (d + s));
        if (post != nullptr) {
          x = fq_mul<FID>(x, ntt_ldg(post + o0));
          y = fq_mul<FID>(y, ntt_ldg(post + o1));
        } else if (post_const != nullptr) {
          Fq cst = ntt_ldg(post_const);
          x = fq_mul<FID>(x, cst);
          y = fq_mul<FID>(y, cst);
        }
        if (sc.enabled) {
          const size_t col = (size_t)sc.rank * sc.cols + blockIdx.y;
          ntt_stg(sc.peer[o0 / sc.rows] + (o0 % sc.rows) * sc.n2 + col, x);
          ntt_stg(sc.peer[o1 / sc.rows] + (o1 % sc.rows) * sc.n2 + col, y);
        } else {
          ntt_stg(out + o0, x);
          ntt_stg(out + o1, y);
        }
      }
    }
    if (d + 1 < q) __syncthreads();
  }
}
#endif

// length-1 "transforms" of a batch: out[b] = in[b] * pre * post (each optional)
#if defined(G753_HOST_EMUL)
template <int FID>
__global__ void k_ntt_scale1(const Fq* in, Fq* out, const Fq* pre, const Fq* post, const Fq* post_const) {
  Fq v = *in;
  if (pre) v = fq_mul<FID>(v, *pre);
  if (post) v = fq_mul<FID>(v, *post);
  else if (post_const) v = fq_mul<FID>(v, *post_const);
  *out = v;
}
#else
template <int FID>
__global__ void k_ntt_scale1(const Fq* in, Fq* out, const Fq* pre, size_t pre_stride, const Fq* post,
                             size_t post_stride, const Fq* post_const) {
  const size_t b = blockIdx.x;
  Fq v = in[b];
  if (pre) v = fq_mul<FID>(v, pre[b * pre_stride]);
  if (post) v = fq_mul<FID>(v, post[b * post_stride]);
  else if (post_const) v = fq_mul<FID>(v, *post_const);
  out[b] = v;
}
#endif

// single-stage form of the same recurrence (q = 1 per launch): the TEST-ONLY host-emulation
// build runs kernels thread by thread and cannot execute block barriers, so it checks the
// stage indexing through this kernel; the product library launches k_ntt_pass.
template <int FID>
__global__ void __launch_bounds__(256, 2)
k_ntt_stage(const Fq* __restrict__ in, Fq* __restrict__ out, const Fq* __restrict__ tw, unsigned log_n,
            unsigned s, const Fq* __restrict__ pre, const Fq* __restrict__ post,
            const Fq* __restrict__ post_const) {
  const unsigned half = 1u << (log_n - 1);
  unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= half) return;
  const unsigned Ns = 1u << s;
  const unsigned k = j & (Ns - 1);
  Fq a = in[j];
  Fq b = in[j + half];
  if (pre != nullptr) {
    a = fq_mul<FID>(a, pre[j]);
    b = fq_mul<FID>(b, pre[j + half]);
  }
  if (s != 0) b = fq_mul<FID>(b, tw[(size_t)k << (log_n - 1 - s)]);  // stage 0: every twiddle is omega^0 = 1
  const unsigned j0 = ((j >> s) << (s + 1)) | k;
  Fq x = fq_add<FID>(a, b);
  Fq y = fq_sub<FID>(a, b);
  if (post != nullptr) {
    x = fq_mul<FID>(x, post[j0]);
    y = fq_mul<FID>(y, post[j0 + Ns]);
  } else if (post_const != nullptr) {
    Fq cst = *post_const;
    x = fq_mul<FID>(x, cst);
    y = fq_mul<FID>(y, cst);
  }
  out[j0] = x;
  out[j0 + Ns] = y;
}

template <int FID>
__global__ void __launch_bounds__(256)
k_vec_op(Fq* __restrict__ a, const Fq* __restrict__ b, int op, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq x = a[i];
  Fq y = b ? b[i] : x;
  Fq r;
  switch (op) {
    case G753_OP_MUL: r = fq_mul<FID>(x, y); break;
    case G753_OP_ADD: r = fq_add<FID>(x, y); break;
    case G753_OP_SUB: r = fq_sub<FID>(x, y); break;
    case G753_OP_SQR: r = fq_sqr<FID>(x); break;
    case G753_OP_INV: r = fq_is_zero(x) ? x : fq_inv<FID>(x); break;
    case G753_OP_TO_MONT: r = fq_to_mont<FID>(x); break;
    case G753_OP_FROM_MONT: r = fq_from_mont<FID>(x); break;
    default: r = x;
  }
  a[i] = r;
}

template <int FID>
__global__ void __launch_bounds__(256) k_vec_scale(Fq* __restrict__ a, const Fq* __restrict__ k, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  a[i] = fq_mul<FID>(a[i], *k);
}
// the same with the factor passed BY VALUE as a kernel argument (96 bytes): a host constant reaches the
// kernel without a device allocation, a copy or a synchronisation on the caller's critical path
template <int FID>
__global__ void __launch_bounds__(256) k_vec_scale_val(Fq* __restrict__ a, const Fq k, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  a[i] = fq_mul<FID>(a[i], k);
}
struct Fq3 {
  Fq v[3];
};

// R1CStoQAP::witness_map element-wise steps (proof-systems/src/groth16/r1cs_to_qap.rs:137-166):
//   ab[i] = (a[i] * b[i] - c[i]) * (g^n - 1)^-1        mul_polynomials_in_evaluation_domain,
//                                                       ab -= c, divide_by_vanishing_poly_on_coset
template <int FID>
__global__ void __launch_bounds__(256)
k_witness_combine(Fq* __restrict__ a, const Fq* __restrict__ b, const Fq* __restrict__ c,
                  const Fq* __restrict__ zinv, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq t = fq_mul<FID>(a[i], b[i]);
  t = fq_sub<FID>(t, c[i]);
  a[i] = fq_mul<FID>(t, *zinv);
}
// h has n + 1 entries (r1cs_to_qap.rs:125-132, 163-166).  The reference builds h as
// `vec![zero; n]` and then MULTIPLIES each h_i by (d2 a_i + d1 b_i) (:127-130), which leaves the
// zeros in place; so h = [-d3 - d1 d2, 0, ..., 0, d1 d2] before the quotient is added to
// h[..n-1].  Reproduced as is: d = d1, d2, d3 (Montgomery form, passed by value).
template <int FID>
__global__ void __launch_bounds__(256)
k_witness_finish(const Fq* __restrict__ ab, const Fq3 d, Fq* __restrict__ h, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  Fq v = fq_zero<FID>();
  if (i + 1 < n) v = ab[i];
  if (i == 0 || i == n) {
    Fq d1d2 = fq_mul<FID>(d.v[0], d.v[1]);
    if (i == n) {
      v = d1d2;
    } else {
      Fq base = fq_sub<FID>(fq_sub<FID>(fq_zero<FID>(), d.v[2]), d1d2);  // h[0] -= d3; h[0] -= d1 d2
      v = fq_add<FID>(base, v);
    }
  }
  h[i] = v;
}

// What one transform (or a batch of equal-length ones) does around the butterflies.
struct NttCall {
  bool inverse = false;             // twiddles omega^-j instead of omega^j
  const Fq* pre = nullptr;          // inputs  *= pre[batch * pre_stride + index]   (first pass)
  size_t pre_stride = 0;
  const Fq* post = nullptr;         // outputs *= post[batch * post_stride + index] (last pass)
  size_t post_stride = 0;
  const Fq* post_const = nullptr;   // outputs *= *post_const (when post == nullptr)
  unsigned batch = 1;               // contiguous vectors of n elements each
  const NttScatter* scatter = nullptr;  // last pass stores to peer memory instead (device build, log_n >= 1)
};

static inline NttCall ntt_call_for_mode(const NttTables& T, int mode) {
  NttCall c;
  c.inverse = (mode == G753_IFFT || mode == G753_COSET_IFFT);
  if (mode == G753_COSET_FFT) c.pre = T.coset;
  if (mode == G753_COSET_IFFT) c.post = T.coset_inv;
  if (mode == G753_IFFT) c.post_const = T.consts;
  return c;
}

// In-place transform of d_data (batch x n elements, n = 2^log_n) using d_tmp (same size) as the
// ping-pong buffer.  The log_n stages are split into ceil(log_n / 8) passes of near-equal depth.
template <int FID>
static int ntt_run(const NttTables& T, cudaStream_t stream, Fq* d_data, Fq* d_tmp, const NttCall& call,
                   uint64_t* launches) {
  const unsigned log_n = T.log_n;
  const size_t n = (size_t)1 << log_n;
  if (call.batch == 0) return G753_OK;
  if (call.batch > 65535) return fail(G753_ERR_BAD_ARG, "ntt: more than 65535 vectors in one batch");
  const Fq* tw = call.inverse ? T.tw_inv : T.tw_fwd;
  Fq* src = d_data;
  Fq* dst = d_tmp;
  if (log_n == 0) {
    // size-1 transforms are the identity up to the pre / post factors
    if (call.pre == nullptr && call.post == nullptr && call.post_const == nullptr) return G753_OK;
  }
#if defined(G753_HOST_EMUL)
  const unsigned stages = log_n ? log_n : 1;
  for (unsigned s = 0; s < stages; s++) {
    for (unsigned bi = 0; bi < call.batch; bi++) {
      const Fq* pre = (s == 0 && call.pre) ? call.pre + bi * call.pre_stride : nullptr;
      const Fq* post = (s == stages - 1 && call.post) ? call.post + bi * call.post_stride : nullptr;
      const Fq* post_c = (s == stages - 1) ? call.post_const : nullptr;
      if (log_n == 0)
        G753_LAUNCH(k_ntt_scale1<FID>, 1, 1, stream, src + bi, dst + bi, pre, post, post_c);
      else
        G753_LAUNCH(k_ntt_stage<FID>, div_up(n / 2, 256), 256, stream, src + bi * n, dst + bi * n, tw, log_n, s, pre,
                    post, post_c);
    }
    if (launches) ++*launches;
    Fq* t = src;
    src = dst;
    dst = t;
  }
#else
  if (log_n == 0) {
    k_ntt_scale1<FID><<<call.batch, 1, 0, stream>>>(src, dst, call.pre, call.pre_stride, call.post, call.post_stride,
                                                    call.post_const);
    if (launches) ++*launches;
    Fq* t = src;
    src = dst;
    dst = t;
  }
  const unsigned passes = (log_n + NTT_MAX_Q - 1) / NTT_MAX_Q;
  unsigned s = 0;
  cudaFuncSetAttribute(k_ntt_pass<FID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT_SMEM);
  for (unsigned pi = 0; pi < passes; pi++) {
    const unsigned q = (log_n - s + (passes - pi) - 1) / (passes - pi);
    const bool first = (pi == 0), last = (pi + 1 == passes);
    const size_t n_cols = n >> q;
    const unsigned cols_per_block = NTT_E >> q;
    dim3 grid(div_up(n_cols, cols_per_block), call.batch);
    NttScatter sc;
    sc.enabled = 0;
    if (last && call.scatter) sc = *call.scatter;
    k_ntt_pass<FID><<<grid, NTT_T, NTT_SMEM, stream>>>(src, dst, tw, log_n, s, q, first ? call.pre : nullptr,
                                                       call.pre_stride, last ? call.post : nullptr, call.post_stride,
                                                       last ? call.post_const : nullptr, sc);
    if (launches) ++*launches;
    s += q;
    Fq* t = src;
    src = dst;
    dst = t;
  }
  if (call.scatter) return launch_check("ntt_run (scattered)");   // the results live in peer memory
#endif
  if (src != d_data) G753_TRY(d2d(d_data, src, sizeof(Fq) * n * call.batch, stream));
  return launch_check("ntt_run");
}

template <int FID>
static int ntt_run(const NttTables& T, cudaStream_t stream, Fq* d_data, Fq* d_tmp, int mode, uint64_t* launches) {
  return ntt_run<FID>(T, stream, d_data, d_tmp, ntt_call_for_mode(T, mode), launches);
}

}  // namespace g753
