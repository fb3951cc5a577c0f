// C ABI of libg753.so (see include/g753.h for the contract and the reference call sites).
//
// Built two ways: nvcc for sm_100a (the product) and, with G753_HOST_EMUL, by g++ into a
// test-only library that runs the barrier-free kernels sequentially on the host so that
// indexing and orchestration can be checked without a GPU (tests/host_emul).
#include "coop.cuh"
#include "ctx.cuh"
#include "ntt_dist.cuh"
#include "ntt_mixed.cuh"
#if defined(G753_HOST_EMUL)
#include "msm_impl.cuh"  // the test-only host build is a single translation unit
G753_INSTANTIATE_GROUP(0)
G753_INSTANTIATE_GROUP(1)
G753_INSTANTIATE_GROUP(2)
G753_INSTANTIATE_GROUP(3)
#else
namespace g753 {
thread_local char g_last_error[512] = "";
}
#endif

using namespace g753;

// zero the coordinates of bases flagged infinite so that (0, 0) is the only encoding the
// accumulation kernels ever see (their digits are dropped in k_msm_digits as well)
__global__ void k_bases_sanitize(Fq* __restrict__ bases, const uint8_t* __restrict__ inf, unsigned n,
                                 unsigned fq_per_point) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (inf[i]) {
    Fq z;
    for (int k = 0; k < NL; k++) z.l[k] = 0;
    for (unsigned k = 0; k < fq_per_point; k++) bases[(size_t)i * fq_per_point + k] = z;
  }
}

// Proving-key loader: the reference's wire format of an affine point is
// x || y || infinity (GroupAffine::write, curves/models/short_weierstrass_projective.rs:185-192), each
// coordinate element as its CANONICAL integer, 12 little-endian u64 = 96 bytes (Fp768::write,
// fields/models/fp_768.rs:784-789; bytes.rs:70-78), the flag one byte (bytes.rs:220-225).  One
// thread per base-field element: gather the 96 unaligned bytes, convert to Montgomery form
// (from_repr, fp_768.rs:627-635) and store into the resident layout.
// Validation as the reference's readers do it: a coordinate >= the modulus is an error (Fp768::read,
// fp_768.rs:791-803: "Attempt to deserialize a field element over the modulus") and so is a flag byte
// other than 0 or 1 (bool::read, bytes.rs:227-237); either sets a bit of *err and the load is rejected.
template <int FID>
__global__ void __launch_bounds__(128)
k_bases_from_wire(const uint8_t* __restrict__ wire, unsigned n, unsigned fq_per_point, Fq* __restrict__ out,
                  uint8_t* __restrict__ inf, unsigned* __restrict__ err) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * fq_per_point) return;
  const unsigned i = (unsigned)(t / fq_per_point), e = (unsigned)(t % fq_per_point);
  const size_t rec = (size_t)fq_per_point * 96 + 1;
  const uint8_t* src = wire + (size_t)i * rec + (size_t)e * 96;
  Fq v;
#pragma unroll
  for (int w = 0; w < NL; w++)
    v.l[w] = (uint32_t)src[4 * w] | ((uint32_t)src[4 * w + 1] << 8) | ((uint32_t)src[4 * w + 2] << 16) |
             ((uint32_t)src[4 * w + 3] << 24);
  const uint8_t flag = wire[(size_t)i * rec + rec - 1];
  const bool is_inf = flag != 0;
  {
    // v < p ?  (borrow out of v - p)
    uint32_t t = sub_cc(v.l[0], G753_FC(FID).p[0]);
#pragma unroll
    for (int w = 1; w < NL; w++) t = subc_cc(v.l[w], G753_FC(FID).p[w]);
    const uint32_t borrow = subc(0, 0);
    (void)t;
    if (!borrow) atomicOr(err, 1u);
    if (flag > 1) atomicOr(err, 2u);
  }
  if (is_inf) v = fq_zero<FID>();   // infinite bases are stored as (0, 0), see k_bases_sanitize
  else v = fq_to_mont<FID>(v);
  out[t] = v;
  if (e == 0) inf[i] = is_inf ? 1 : 0;
}

// ------------------------------------------------------------------------------------
// dispatch helpers
// ------------------------------------------------------------------------------------
static int msm_any(g753_ctx* ctx, const g753_bases* b, size_t first, size_t count,
                   const uint32_t* d_scalars, void* d_out) {
  switch (b->group) {
    case G753_MNT4_G1: return msm_dispatch<0>(ctx, b, first, count, d_scalars, d_out);
    case G753_MNT4_G2: return msm_dispatch<1>(ctx, b, first, count, d_scalars, d_out);
    case G753_MNT6_G1: return msm_dispatch<2>(ctx, b, first, count, d_scalars, d_out);
    case G753_MNT6_G2: return msm_dispatch<3>(ctx, b, first, count, d_scalars, d_out);
  }
  return fail(G753_ERR_BAD_ARG, "unknown group");
}

template <int FID>
static int ntt_tables_get(g753_ctx* ctx, unsigned log_n, const NttTables** out);
template <int FID>
static int ntt_field(g753_ctx* ctx, Fq* d_data, Fq* d_tmp, unsigned log_n, int mode) {
  const NttTables* T = nullptr;
  G753_TRY(ntt_tables_get<FID>(ctx, log_n, &T));
  return ntt_run<FID>(*T, ctx->stream, d_data, d_tmp, mode, &ctx->launches);
}

// R1CStoQAP::witness_map after the constraint evaluation (r1cs_to_qap.rs:121-166), chained on
// the device: 7 transforms + the element-wise steps, no host round trips.
// second half of the witness map: a, b, c hold the coset evaluations (ifft then coset_fft of the
// constraint evaluations); h = coset_ifft((a b - c) / Z) with the d1, d2, d3 terms (r1cs_to_qap.rs:137-166)
template <int FID>
static int witness_map_tail_field(g753_ctx* ctx, Fq* a, const Fq* b, const Fq* c, unsigned log_n, const Fq3& d_d, Fq* h) {
  const size_t n = (size_t)1 << log_n;
  G753_TRY(ctx->scratch_io.reserve(sizeof(Fq) * n + 1024));
  Fq* tmp = (Fq*)ctx->scratch_io.ptr;
  const NttTables* T = nullptr;
  G753_TRY(ntt_tables_get<FID>(ctx, log_n, &T));
  const unsigned L = log_n ? log_n : 1;
  G753_LAUNCH(k_witness_combine<FID>, div_up(n, 256), 256, ctx->stream, a, b, c, T->consts + 1 + 4 * L, n);
  ctx->launches++;
  G753_TRY(ntt_field<FID>(ctx, a, tmp, log_n, G753_COSET_IFFT));
  G753_LAUNCH(k_witness_finish<FID>, div_up(n + 1, 256), 256, ctx->stream, a, d_d, h, n);
  ctx->launches++;
  return launch_check("witness_map");
}
template <int FID>
static int witness_map_field(g753_ctx* ctx, Fq* a, Fq* b, Fq* c, unsigned log_n, const Fq3& d_d, Fq* h) {
  const size_t n = (size_t)1 << log_n;
  G753_TRY(ctx->scratch_io.reserve(sizeof(Fq) * n + 1024));
  Fq* tmp = (Fq*)ctx->scratch_io.ptr;
  G753_TRY(ntt_field<FID>(ctx, a, tmp, log_n, G753_IFFT));
  G753_TRY(ntt_field<FID>(ctx, b, tmp, log_n, G753_IFFT));
  G753_TRY(ntt_field<FID>(ctx, a, tmp, log_n, G753_COSET_FFT));
  G753_TRY(ntt_field<FID>(ctx, b, tmp, log_n, G753_COSET_FFT));
  G753_TRY(ntt_field<FID>(ctx, c, tmp, log_n, G753_IFFT));
  G753_TRY(ntt_field<FID>(ctx, c, tmp, log_n, G753_COSET_FFT));
  return witness_map_tail_field<FID>(ctx, a, b, c, log_n, d_d, h);
}

// ---- sharded four-step NTT (ntt_dist.cuh) ----------------------------------------------------
template <int FID>
static int ntt_tables_get(g753_ctx* ctx, unsigned log_n, const NttTables** out) {
  std::map<unsigned, NttTables>& m = ctx->tables[FID];
  auto it = m.find(log_n);
  if (it == m.end()) {
    NttTables T;
    int rc = ntt_tables_build<FID>(T, log_n, ctx->stream, &ctx->launches);
    if (rc != G753_OK) {
      T.release();
      return rc;
    }
    it = m.emplace(log_n, T).first;
  }
  *out = &it->second;
  return G753_OK;
}

template <int FID>
static int shard_build(g753_ctx* ctx, g753_ntt_shard* p) {
  const unsigned L = p->log_n ? p->log_n : 1;
  G753_TRY(ntt_consts_build<FID>(&p->consts, p->log_n, ctx->stream, &ctx->launches));
  const Fq* K = p->consts;
  const Fq *ninv = K, *w = K + 1, *wi = K + 1 + L, *g = K + 1 + 2 * L, *gi = K + 1 + 3 * L;
  // g^n2 = g^(2^log_n2), g^-n1 = g^-(2^log_n1); for log_n == 0 both exponents are 1
  const Fq* g_n2 = p->log_n ? K + 1 + 2 * L + p->log_n2 : g;
  const Fq* gi_n1 = p->log_n ? K + 1 + 3 * L + p->log_n1 : gi;
  // when log_n2 == log_n (n1 == 1) the chain stops one short of g^n: not needed, the table has len 1
  const size_t bytes1 = sizeof(Fq) * p->cols * p->n1, bytes2 = sizeof(Fq) * p->rows * p->n2;
  G753_TRY(dev_alloc((void**)&p->t1f, bytes1));
  G753_TRY(dev_alloc((void**)&p->t1i, bytes1));
  G753_TRY(dev_alloc((void**)&p->cp, bytes1));
  G753_TRY(dev_alloc((void**)&p->cq, bytes2));
  const Fq* none = nullptr;
  const uint64_t c0 = (uint64_t)p->rank * p->cols, r0 = (uint64_t)p->rank * p->rows;
  G753_LAUNCH(k_pow_table<FID>, div_up(p->cols, 64), 64, ctx->stream, p->t1f, (unsigned)p->cols, (unsigned)p->n1, c0,
              none, w, none, none);
  G753_LAUNCH(k_pow_table<FID>, div_up(p->cols, 64), 64, ctx->stream, p->t1i, (unsigned)p->cols, (unsigned)p->n1, c0,
              none, wi, ninv, none);
  G753_LAUNCH(k_pow_table<FID>, div_up(p->cols, 64), 64, ctx->stream, p->cp, (unsigned)p->cols, (unsigned)p->n1, c0,
              p->log_n1 ? g_n2 : none, none, none, g);
  G753_LAUNCH(k_pow_table<FID>, div_up(p->rows, 64), 64, ctx->stream, p->cq, (unsigned)p->rows, (unsigned)p->n2, r0,
              p->log_n2 ? gi_n1 : none, none, none, gi);
  ctx->launches += 4;
  G753_TRY(launch_check("k_pow_table"));
  return stream_sync(ctx->stream);
}

template <int FID>
static int shard_step1(g753_ctx* ctx, const g753_ntt_shard* p, Fq* data, Fq* send, int mode) {
  const NttTables* T = nullptr;
  G753_TRY(ntt_tables_get<FID>(ctx, p->log_n1, &T));
  NttCall c;
  c.inverse = (mode == G753_IFFT || mode == G753_COSET_IFFT);
  c.batch = (unsigned)p->cols;
  if (mode == G753_COSET_FFT) {
    c.pre = p->cp;
    c.pre_stride = p->n1;
  }
  c.post = c.inverse ? p->t1i : p->t1f;
  c.post_stride = p->n1;
  G753_TRY(ctx->scratch_io.reserve(sizeof(Fq) * p->cols * p->n1 + 1024));
  G753_TRY(ntt_run<FID>(*T, ctx->stream, data, (Fq*)ctx->scratch_io.ptr, c, &ctx->launches));
  // Y[i2l][h][k1l] -> send[h][k1l][i2l]
  const size_t total = p->cols * p->n1;
  G753_LAUNCH(k_permute3, div_up(total, 256), 256, ctx->stream, data, send, (unsigned)p->cols, p->world,
              (unsigned)p->rows, (size_t)1, p->rows * p->cols, p->cols);
  ctx->launches++;
  return launch_check("shard_step1");
}

#if !defined(G753_HOST_EMUL)
// step 1 with the exchange fused into the last butterfly pass (NttScatter): peer_z[h] is rank h's
// row buffer Z (rows x n2 elements), mapped into this process
template <int FID>
static int shard_step1_fused(g753_ctx* ctx, const g753_ntt_shard* p, Fq* data, void* const* peer_z, int mode) {
  const NttTables* T = nullptr;
  G753_TRY(ntt_tables_get<FID>(ctx, p->log_n1, &T));
  NttCall c;
  c.inverse = (mode == G753_IFFT || mode == G753_COSET_IFFT);
  c.batch = (unsigned)p->cols;
  if (mode == G753_COSET_FFT) {
    c.pre = p->cp;
    c.pre_stride = p->n1;
  }
  c.post = c.inverse ? p->t1i : p->t1f;
  c.post_stride = p->n1;
  NttScatter sc;
  for (unsigned h = 0; h < 8; h++) sc.peer[h] = h < p->world ? (Fq*)peer_z[h] : nullptr;
  sc.rows = (unsigned)p->rows;
  sc.cols = (unsigned)p->cols;
  sc.rank = p->rank;
  sc.n2 = p->n2;
  sc.enabled = 1;
  c.scatter = &sc;
  G753_TRY(ctx->scratch_io.reserve(sizeof(Fq) * p->cols * p->n1 + 1024));
  return ntt_run<FID>(*T, ctx->stream, data, (Fq*)ctx->scratch_io.ptr, c, &ctx->launches);
}
#endif

// step 2 on a row buffer that already holds Z[k1l][i2] (written by the peers' fused step 1)
template <int FID>
static int shard_step2_local(g753_ctx* ctx, const g753_ntt_shard* p, Fq* z, int mode) {
  const NttTables* T = nullptr;
  G753_TRY(ntt_tables_get<FID>(ctx, p->log_n2, &T));
  NttCall c;
  c.inverse = (mode == G753_IFFT || mode == G753_COSET_IFFT);
  c.batch = (unsigned)p->rows;
  if (mode == G753_COSET_IFFT) {
    c.post = p->cq;
    c.post_stride = p->n2;
  }
  G753_TRY(ctx->scratch_io.reserve(sizeof(Fq) * p->rows * p->n2 + 1024));
  return ntt_run<FID>(*T, ctx->stream, z, (Fq*)ctx->scratch_io.ptr, c, &ctx->launches);
}

template <int FID>
static int shard_step2(g753_ctx* ctx, const g753_ntt_shard* p, const Fq* recv, Fq* data, int mode) {
  const NttTables* T = nullptr;
  G753_TRY(ntt_tables_get<FID>(ctx, p->log_n2, &T));
  // recv[g][k1l][i2l] -> Z[k1l][g cols + i2l]
  const size_t total = p->rows * p->n2;
  G753_LAUNCH(k_permute3, div_up(total, 256), 256, ctx->stream, recv, data, p->world, (unsigned)p->rows,
              (unsigned)p->cols, p->cols, p->n2, (size_t)1);
  ctx->launches++;
  NttCall c;
  c.inverse = (mode == G753_IFFT || mode == G753_COSET_IFFT);
  c.batch = (unsigned)p->rows;
  if (mode == G753_COSET_IFFT) {
    c.post = p->cq;
    c.post_stride = p->n2;
  }
  G753_TRY(ctx->scratch_io.reserve(sizeof(Fq) * total + 1024));
  G753_TRY(ntt_run<FID>(*T, ctx->stream, data, (Fq*)ctx->scratch_io.ptr, c, &ctx->launches));
  return launch_check("shard_step2");
}

// ---- mixed-radix transform (ntt_mixed.cuh) -----------------------------------------------------
template <int FID>
static int mixed_tables_get(g753_ctx* ctx, uint64_t N, const MixedTables** out) {
  std::map<uint64_t, MixedTables>& mm = ctx->mixed[FID];
  auto it = mm.find(N);
  if (it != mm.end()) {
    *out = &it->second;
    return G753_OK;
  }
  MixedTables T;
  uint32_t exp[NL];
  if (!mixed_split(FID, N, &T.a, &T.m, exp)) return fail(G753_ERR_DOMAIN, "no mixed-radix domain of this size");
  T.N = N;
  const size_t n2 = (size_t)1 << T.a;
  uint32_t* d_exp = nullptr;
  int rc = dev_alloc((void**)&d_exp, sizeof(exp));
  if (rc == G753_OK) rc = dev_alloc((void**)&T.consts, sizeof(Fq) * 8);
  if (rc == G753_OK) rc = dev_alloc((void**)&T.zeta, sizeof(Fq) * 2 * T.m);
  if (rc == G753_OK) rc = dev_alloc((void**)&T.tw, sizeof(Fq) * 2 * N);
  if (rc == G753_OK) rc = dev_alloc((void**)&T.coset, sizeof(Fq) * N);
  if (rc == G753_OK) rc = dev_alloc((void**)&T.coset_inv, sizeof(Fq) * N);
  if (rc == G753_OK) rc = h2d(d_exp, exp, sizeof(exp), ctx->stream);
  if (rc == G753_OK) {
    const Fq* K = T.consts;
    const Fq* none = nullptr;
    G753_LAUNCH(k_mixed_setup<FID>, 1, 1, ctx->stream, d_exp, (uint64_t)N, (uint64_t)n2, T.consts);
    // zeta^j and zeta^-j: one row of m powers each
    G753_LAUNCH(k_pow_table<FID>, 1, 1, ctx->stream, T.zeta, 1u, T.m, (uint64_t)0, K + 3, none, none, none);
    G753_LAUNCH(k_pow_table<FID>, 1, 1, ctx->stream, T.zeta + T.m, 1u, T.m, (uint64_t)0, K + 4, none, none, none);
    // omega^(+-i2 k1): row k1, base omega^(+-k1)
    G753_LAUNCH(k_pow_table<FID>, div_up(T.m, 64), 64, ctx->stream, T.tw, T.m, (unsigned)n2, (uint64_t)0, none, K,
                none, none);
    G753_LAUNCH(k_pow_table<FID>, div_up(T.m, 64), 64, ctx->stream, T.tw + N, T.m, (unsigned)n2, (uint64_t)0, none,
                K + 1, none, none);
    // g^i and g^-i / N, by doubling tables
    G753_LAUNCH(k_table_init<FID>, 1, 1, ctx->stream, T.coset, (const Fq*)nullptr);
    G753_LAUNCH(k_table_init<FID>, 1, 1, ctx->stream, T.coset_inv, K + 2);
    ctx->launches += 7;
    rc = launch_check("mixed tables");
  }
  Fq* steps = nullptr;  // g^(2^l), g^-(2^l)
  if (rc == G753_OK) rc = dev_alloc((void**)&steps, sizeof(Fq) * 2 * 32);
  if (rc == G753_OK) {
    G753_LAUNCH(k_square_chain<FID>, 1, 1, ctx->stream, T.consts + 5, steps, 32u);
    G753_LAUNCH(k_square_chain<FID>, 1, 1, ctx->stream, T.consts + 6, steps + 32, 32u);
    for (unsigned l = 0; ((size_t)1 << l) < N; l++) {
      unsigned len = 1u << l;
      G753_LAUNCH(k_table_double<FID>, div_up(len, 128), 128, ctx->stream, T.coset, steps + l, len, (unsigned)N);
      G753_LAUNCH(k_table_double<FID>, div_up(len, 128), 128, ctx->stream, T.coset_inv, steps + 32 + l, len,
                  (unsigned)N);
      ctx->launches += 2;
    }
    rc = launch_check("mixed coset tables");
  }
  if (rc == G753_OK) rc = stream_sync(ctx->stream);
  dev_free(d_exp);
  dev_free(steps);
  if (rc != G753_OK) {
    T.release();
    return rc;
  }
  *out = &mm.emplace(N, T).first->second;
  return G753_OK;
}

// d_data: N elements, in place; d_tmp: N elements
template <int FID>
static int mixed_run(g753_ctx* ctx, Fq* d_data, Fq* d_tmp, uint64_t N, int mode) {
  const MixedTables* T = nullptr;
  G753_TRY(mixed_tables_get<FID>(ctx, N, &T));
  const NttTables* T2 = nullptr;
  G753_TRY(ntt_tables_get<FID>(ctx, T->a, &T2));
  const bool inverse = (mode == G753_IFFT || mode == G753_COSET_IFFT);
  const size_t n2 = (size_t)1 << T->a;
  // step 1: column DFTs of length m (coset_fft: inputs scaled by g^i on the fly), in two stages when m = q1 q2
  const Fq* zt = T->zeta + (inverse ? T->m : 0);
  const Fq* pre = mode == G753_COSET_FFT ? T->coset : (const Fq*)nullptr;
  const unsigned q1 = mixed_q1(T->m), q2 = T->m / q1;
  Fq *cols = d_tmp, *other = d_data;       // where the column DFTs end up / the free buffer
  if (q1 > 1) {
    G753_LAUNCH(k_small_dft_a<FID>, div_up(N, 128), 128, ctx->stream, d_data, d_tmp, zt, T->m, q1, q2, n2, pre);
    G753_LAUNCH(k_small_dft_c<FID>, div_up(N, 128), 128, ctx->stream, d_tmp, d_data, zt, T->m, q1, q2, n2);
    ctx->launches += 2;
    cols = d_data;
    other = d_tmp;
  } else {
    G753_LAUNCH(k_small_dft<FID>, div_up(N, 128), 128, ctx->stream, d_data, d_tmp, zt, T->m, n2, pre);
    ctx->launches++;
  }
  // steps 2 + 3: twiddles as the pre table of the m batched radix-2 transforms
  NttCall c;
  c.inverse = inverse;
  c.batch = T->m;
  c.pre = T->tw + (inverse ? N : 0);
  c.pre_stride = n2;
  G753_TRY(ntt_run<FID>(*T2, ctx->stream, cols, other, c, &ctx->launches));
  // step 4: Z[k1][k2] -> out[k1 + m k2]
  G753_LAUNCH(k_permute3, div_up(N, 256), 256, ctx->stream, cols, other, T->m, (unsigned)n2, 1u, (size_t)1,
              (size_t)T->m, (size_t)0);
  ctx->launches++;
  if (mode == G753_IFFT)
    G753_LAUNCH(k_vec_scale<FID>, div_up(N, 256), 256, ctx->stream, other, T->consts + 2, (size_t)N);
  else if (mode == G753_COSET_IFFT)
    G753_LAUNCH(k_vec_op<FID>, div_up(N, 256), 256, ctx->stream, other, T->coset_inv, G753_OP_MUL, (size_t)N);
  if (inverse) ctx->launches++;
  if (other != d_data) G753_TRY(d2d(d_data, other, sizeof(Fq) * N, ctx->stream));
  return launch_check("mixed_run");
}

// ------------------------------------------------------------------------------------
// test kernels: group law / field ops through the real device code
// ------------------------------------------------------------------------------------
#if !defined(G753_HOST_EMUL)
// integer-pipe roofline probe: dependent Montgomery products per thread
template <int VARIANT>
__global__ void __launch_bounds__(256) k_mac_probe(const Fq* seed, Fq* sink, int iters) {
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  Fq x = seed[t & 255];
  Fq y = seed[(t + 7) & 255];
  if (VARIANT == 0) {
    for (int k = 0; k < iters; k++) x = fq_mul<1>(x, y);
  } else if (VARIANT == 1) {
    for (int k = 0; k < iters; k++) x = fq_sqr<1>(x);
  } else if (VARIANT == 3) {
    for (int k = 0; k < iters; k++) x = fq_add<1>(fq_inv<1>(x), y);   // safegcd inversion chain
  } else {
    // 24 independent 64-bit accumulators, each a dependent IMAD.WIDE chain: 576 wide MACs / iter
    unsigned long long acc[NL];
#pragma unroll
    for (int i = 0; i < NL; i++) acc[i] = x.l[i];
    for (int k = 0; k < iters; k++) {
#pragma unroll
      for (int r = 0; r < NL; r++) {
#pragma unroll
        for (int i = 0; i < NL; i++) acc[i] = (unsigned long long)y.l[r] * (unsigned)(acc[i]) + acc[i];
      }
    }
#pragma unroll
    for (int i = 0; i < NL; i++) x.l[i] = (uint32_t)acc[i] ^ (uint32_t)(acc[i] >> 32);
  }
  if (x.l[0] == 0x12345678u && x.l[5] == 0x9abcdef0u) sink[t & 255] = x;  // keep the chain live
}

// latency probes of the warp-cooperative arithmetic (one warp): 5 = dependent coop_mul chain, 6 = dependent
// fused linear operations, 7 = the G1 doubling micro-program, 8 = the G1 addition (head + tail)
__global__ void __launch_bounds__(32) k_coop_probe(int variant, const Fq* seed, Fq* sink, int iters) {
  extern __shared__ uint4 coop_probe_smem[];
  CoopEc<0> ec;
  ec.init((uint32_t*)coop_probe_smem);
  typedef CoopGroup<0> Gp;
  ec.w.load(Gp::P, seed, 4);
  ec.w.load(Gp::Q, seed + 4, 4);
  if (variant == 5) {
    uint32_t A[3], B[3];
    ec.w.ld(A, Gp::P);
    ec.w.ld(B, Gp::Q);
    for (int k = 0; k < iters; k++) coop_mul(A, A, B, ec.w.n, ec.w.np);
    ec.w.st(Gp::P, A);
  } else if (variant == 6) {
    uint32_t A[3], B[3], Z[3] = {0, 0, 0};
    ec.w.ld(A, Gp::P);
    ec.w.ld(B, Gp::Q);
    for (int k = 0; k < iters; k++) coop_add3(A, A, B, Z, 0);
    ec.w.st(Gp::P, A);
  } else if (variant == 7) {
    for (int k = 0; k < iters; k++) ec.dbl();
  } else {
    for (int k = 0; k < iters; k++) ec.add_q();
  }
  ec.w.store(sink, Gp::P, 4);
}

// The OTHER multiplier pipe (bounded experiment): the instruction mix of a 753-bit Montgomery
// product on 15 limbs of 52 bits through the FP64 unit.  A 52 x 52 -> 104 bit limb product is two DFMAs in
// round-to-zero (hi = fma(a, b, 2^104) keeps the top 52 bits in its mantissa, lo = fma(a, b, 2^104 + 2^52 - hi)
// the low 52) and one DADD; the raw bit patterns are summed as 64-bit integers column by column.  One
// iteration = the 15 x 15 limb products of a x b plus the 15 x 15 of the reduction M x p (M folded from the
// low columns) = 450 limb products, the count of one Montgomery product; the columns feed the next
// iteration, so the chain is dependent like the IMAD probe's.
__global__ void __launch_bounds__(256) k_dfma_probe(const double* seed, double* sink, int iters) {
  constexpr int L = 15;
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  double a[L], b[L], p[L];
#pragma unroll
  for (int i = 0; i < L; i++) {
    a[i] = seed[(t + i) & 255];
    b[i] = seed[(t + 31 * i + 7) & 255];
    p[i] = seed[(17 * i + 3) & 255];
  }
  const double C1 = 0x1p104, C2 = 0x1p104 + 0x1p52;
  for (int k = 0; k < iters; k++) {
    long long col[2 * L];
#pragma unroll
    for (int c = 0; c < 2 * L; c++) col[c] = 0;
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
#pragma unroll
      for (int i = 0; i < L; i++) {
        // second pass: the multiplier limb is derived from the running low column (the "M" of the reduction)
        const double x = pass == 0 ? a[i] : (double)(col[i] & 0xFFFFFFFFFFFFFll);
#pragma unroll
        for (int j = 0; j < L; j++) {
          const double y = pass == 0 ? b[j] : p[j];
          const double hi = __fma_rz(x, y, C1);
          const double lo = __fma_rz(x, y, C2 - hi);
          col[i + j + 1] += __double_as_longlong(hi);
          col[i + j] += __double_as_longlong(lo);
        }
      }
    }
    // fold the upper columns back into the operand (keeps the chain dependent; exact values do not matter)
#pragma unroll
    for (int i = 0; i < L; i++) a[i] = (double)((col[L + i] ^ col[i]) & 0xFFFFFFFFFFFFFll);
  }
  double acc = 0;
#pragma unroll
  for (int i = 0; i < L; i++) acc += a[i];
  if (acc == 1234.5) sink[t & 255] = acc;
}
#endif

#if !defined(G753_HOST_EMUL)
// test hook for the warp-cooperative field arithmetic (coop.cuh): four elements per warp pass, raw
// (lazily reduced, not canonical) results so that they can be compared limb for limb with the Python model
template <int FID>
__global__ void __launch_bounds__(32)
k_coop_op(int op, unsigned k, const Fq* __restrict__ a, const Fq* __restrict__ b, unsigned n, Fq* __restrict__ out) {
  extern __shared__ uint4 coop_test_smem[];
  CoopWarp<FID> w;
  w.init((uint32_t*)coop_test_smem, 12);
  const unsigned o = coop_octet(), l = coop_lane();
  for (unsigned base = 0; base < n; base += 4) {
    const unsigned cnt = n - base < 4 ? n - base : 4;
    w.load(0, a + base, cnt);
    w.load(4, b + base, cnt);
    uint32_t A[3], B[3], R[3];
    w.ld(A, o);
    w.ld(B, 4 + o);
    if (op == 0) {
      coop_mul(R, A, B, w.n, w.np);
    } else if (op == 1) {
      uint32_t Z[3] = {0, 0, 0};
      coop_add3(R, A, B, Z, 0);
    } else if (op == 2) {
      uint32_t Y[3] = {~B[0], ~B[1], ~B[2]};
      const uint32_t* kpl = w.kp + k * 24 + 3 * l;
      uint32_t Z[3] = {kpl[0], kpl[1], kpl[2]};
      coop_add3(R, A, Y, Z, l == 0 ? 1u : 0u);
    } else {
      R[0] = A[0];
      R[1] = A[1];
      R[2] = A[2];
    }
    __syncwarp();
    w.st(8 + o, R);
    __syncwarp();
    if (op == 3) w.canonicalize(8, cnt);
    if (op == 4) {   // is_zero of a value below 2 p -> 0 / 1 in limb 0
      bool z[4];
      for (unsigned i = 0; i < 4; i++) z[i] = w.is_zero(8 + i);
      uint32_t Zr[3] = {z[o] ? 1u : 0u, 0, 0};
      __syncwarp();
      if (l == 0) w.st(8 + o, Zr);
      else {
        uint32_t zero3[3] = {0, 0, 0};
        w.st(8 + o, zero3);
      }
      __syncwarp();
    }
    w.store(out + base, 8, cnt);
    __syncwarp();
  }
}
#endif

// ------------------------------------------------------------------------------------
// extern "C"
// ------------------------------------------------------------------------------------
extern "C" {

const char* g753_last_error(void) { return g_last_error; }
const char* g753_version(void) { return "g753 0.2 (sm_100a)"; }
#ifndef G753_SOURCE_HASH
#define G753_SOURCE_HASH "unrecorded"
#endif
const char* g753_source_hash(void) { return G753_SOURCE_HASH; }

int g753_device_count(int* count) {
  if (!count) return fail(G753_ERR_BAD_ARG, "null count");
#if defined(G753_HOST_EMUL)
  *count = 1;
  return G753_OK;
#else
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    return cuda_fail(e, "cudaGetDeviceCount");
  }
  return G753_OK;
#endif
}

int g753_ctx_create(int device, g753_ctx** out) {
  if (!out) return fail(G753_ERR_BAD_ARG, "null out");
  *out = nullptr;
#if !defined(G753_HOST_EMUL)
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail(G753_ERR_NO_DEVICE, "no CUDA device (this library has no CPU path)");
  if (device < 0 || device >= count) return fail(G753_ERR_BAD_ARG, "device index out of range");
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
#endif
  g753_ctx* ctx = new (std::nothrow) g753_ctx();
  if (!ctx) return fail(G753_ERR_OOM, "host allocation failed");
  ctx->device = device;
#if !defined(G753_HOST_EMUL)
  e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete ctx;
    return cuda_fail(e, "cudaStreamCreate");
  }
  ctx->ev_ok = true;
  for (int i = 0; i <= MSM_PHASES; i++)
    if (cudaEventCreate(&ctx->ev[i]) != cudaSuccess) ctx->ev_ok = false;
  ctx->copy_ok = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < g753_ctx::MAX_CHUNKS && ctx->copy_ok; i++)
    if (cudaEventCreateWithFlags(&ctx->chunk_ev[i], cudaEventDisableTiming) != cudaSuccess) ctx->copy_ok = false;
  if (ctx->copy_ok && cudaEventCreateWithFlags(&ctx->ready_ev, cudaEventDisableTiming) != cudaSuccess) ctx->copy_ok = false;
  (void)cudaGetLastError();
#endif
  const char* fc = getenv("G753_MSM_C");
  if (fc) ctx->forced_c = atoi(fc);
  const char* fa = getenv("G753_MSM_AFFINE");
  if (fa) ctx->forced_affine = atoi(fa) ? 1 : 0;
  const char* tb = getenv("G753_TREE_BATCH");
  if (tb && atoi(tb) > 0) ctx->tree_batch = atoi(tb);
  const char* tw = getenv("G753_TREE_WAVES");
  if (tw && atoi(tw) > 0) ctx->tree_waves = atoi(tw);
  *out = ctx;
  return G753_OK;
}

int g753_ctx_destroy(g753_ctx* ctx) {
  if (!ctx) return G753_OK;
  use_device(ctx);
  stream_sync(ctx->stream);
  ctx->scratch.release();
  ctx->scratch_io.release();
  for (int f = 0; f < 2; f++) {
    for (auto& kv : ctx->tables[f]) kv.second.release();
    for (auto& kv : ctx->mixed[f]) kv.second.release();
  }
#if !defined(G753_HOST_EMUL)
  if (ctx->ev_ok)
    for (int i = 0; i <= MSM_PHASES; i++) cudaEventDestroy(ctx->ev[i]);
  if (ctx->copy_ok) {
    for (int i = 0; i < g753_ctx::MAX_CHUNKS; i++) cudaEventDestroy(ctx->chunk_ev[i]);
    cudaEventDestroy(ctx->ready_ev);
    cudaStreamDestroy(ctx->copy_stream);
  }
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
#endif
  delete ctx;
  return G753_OK;
}

int g753_group_coord_limbs(int group) { return 12 * group_k(group); }

int g753_bases_upload(g753_ctx* ctx, int group, const uint64_t* coords, const uint8_t* infinity, size_t n,
                      g753_bases** out) {
  CHECK_CTX(ctx);
  if (!out) return fail(G753_ERR_BAD_ARG, "null out");
  *out = nullptr;
  const int k = group_k(group);
  if (k == 0) return fail(G753_ERR_BAD_ARG, "unknown group");
  if (n && !coords) return fail(G753_ERR_BAD_ARG, "null coords");
  if (n > 0x7fffffffull) return fail(G753_ERR_BAD_ARG, "too many bases");
  std::lock_guard<std::mutex> lock(ctx->mu);
  g753_bases* b = new (std::nothrow) g753_bases();
  if (!b) return fail(G753_ERR_OOM, "host allocation failed");
  b->group = group;
  b->n = n;
  const size_t pt_bytes = (size_t)2 * k * 96;
  int rc = dev_alloc(&b->d_points, pt_bytes * n);
  if (rc == G753_OK && n) rc = h2d(b->d_points, coords, pt_bytes * n, ctx->stream);
  if (rc == G753_OK && infinity && n) {
    rc = dev_alloc((void**)&b->d_inf, n);
    if (rc == G753_OK) rc = h2d(b->d_inf, infinity, n, ctx->stream);
    if (rc == G753_OK) {
      G753_LAUNCH(k_bases_sanitize, div_up(n, 256), 256, ctx->stream, (Fq*)b->d_points, b->d_inf, (unsigned)n,
                  (unsigned)(2 * k));
      ctx->launches++;
      rc = launch_check("k_bases_sanitize");
    }
  }
  if (rc == G753_OK) rc = stream_sync(ctx->stream);
  if (rc != G753_OK) {
    dev_free(b->d_points);
    dev_free(b->d_inf);
    delete b;
    return rc;
  }
  *out = b;
  return G753_OK;
}

int g753_bases_upload_wire(g753_ctx* ctx, int group, const uint8_t* wire, size_t n, g753_bases** out) {
  CHECK_CTX(ctx);
  if (!out) return fail(G753_ERR_BAD_ARG, "null out");
  *out = nullptr;
  const int k = group_k(group);
  if (k == 0) return fail(G753_ERR_BAD_ARG, "unknown group");
  if (n && !wire) return fail(G753_ERR_BAD_ARG, "null data");
  if (n > 0x7fffffffull) return fail(G753_ERR_BAD_ARG, "too many bases");
  std::lock_guard<std::mutex> lock(ctx->mu);
  g753_bases* b = new (std::nothrow) g753_bases();
  if (!b) return fail(G753_ERR_OOM, "host allocation failed");
  b->group = group;
  b->n = n;
  const size_t rec = (size_t)2 * k * 96 + 1;
  uint8_t* d_wire = nullptr;
  unsigned* d_err = nullptr;
  unsigned h_err = 0;
  int rc = dev_alloc(&b->d_points, (size_t)2 * k * 96 * n);
  if (rc == G753_OK) rc = dev_alloc((void**)&b->d_inf, n ? n : 1);
  if (rc == G753_OK && n) rc = dev_alloc((void**)&d_wire, rec * n + 256);
  if (rc == G753_OK && n) {
    d_err = (unsigned*)(d_wire + ((rec * n + 15) & ~(size_t)15));
    rc = dev_memset(d_err, 0, sizeof(unsigned), ctx->stream);
  }
  if (rc == G753_OK && n) rc = h2d(d_wire, wire, rec * n, ctx->stream);
  if (rc == G753_OK && n) {
    const bool mnt4 = (group == G753_MNT4_G1 || group == G753_MNT4_G2);   // base field id = 0 for MNT4
    if (mnt4)
      G753_LAUNCH(k_bases_from_wire<0>, div_up(n * 2 * k, 128), 128, ctx->stream, d_wire, (unsigned)n, (unsigned)(2 * k),
                  (Fq*)b->d_points, b->d_inf, d_err);
    else
      G753_LAUNCH(k_bases_from_wire<1>, div_up(n * 2 * k, 128), 128, ctx->stream, d_wire, (unsigned)n, (unsigned)(2 * k),
                  (Fq*)b->d_points, b->d_inf, d_err);
    ctx->launches++;
    rc = launch_check("k_bases_from_wire");
    if (rc == G753_OK) rc = d2h(&h_err, d_err, sizeof(unsigned), ctx->stream);
  }
  if (rc == G753_OK) rc = stream_sync(ctx->stream);
  dev_free(d_wire);
  if (rc == G753_OK && h_err)
    rc = fail(G753_ERR_BAD_ARG, (h_err & 1u) ? "wire data: a coordinate is not below the field modulus (Fp768::read rejects it)"
                                              : "wire data: an infinity flag byte is neither 0 nor 1 (bool::read rejects it)");
  if (rc != G753_OK) {
    dev_free(b->d_points);
    dev_free(b->d_inf);
    delete b;
    return rc;
  }
  *out = b;
  return G753_OK;
}

int g753_bases_free(g753_ctx* ctx, g753_bases* b) {
  if (!b) return G753_OK;
  if (ctx) {
    use_device(ctx);
    stream_sync(ctx->stream);
  }
  dev_free(b->d_points);
  dev_free(b->d_inf);
  delete b;
  return G753_OK;
}

size_t g753_bases_len(const g753_bases* b) { return b ? b->n : 0; }

int g753_ctx_set_stream(g753_ctx* ctx, void* cuda_stream) {
  CHECK_CTX(ctx);
  std::lock_guard<std::mutex> lock(ctx->mu);
  G753_TRY(stream_sync(ctx->stream));
#if !defined(G753_HOST_EMUL)
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
#endif
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return G753_OK;
}

int g753_bases_generate(g753_ctx* ctx, int group, const uint64_t* gen_xy, uint64_t seed, size_t n,
                        g753_bases** out) {
  CHECK_CTX(ctx);
  if (!out || !gen_xy) return fail(G753_ERR_BAD_ARG, "null pointer");
  *out = nullptr;
  const int k = group_k(group);
  if (k == 0) return fail(G753_ERR_BAD_ARG, "unknown group");
  if (n > 0x7fffffffull) return fail(G753_ERR_BAD_ARG, "too many bases");
  std::lock_guard<std::mutex> lock(ctx->mu);
  g753_bases* b = new (std::nothrow) g753_bases();
  if (!b) return fail(G753_ERR_OOM, "host allocation failed");
  b->group = group;
  b->n = n;
  int rc = dev_alloc(&b->d_points, (size_t)2 * k * 96 * n);
  if (rc == G753_OK) {
    switch (group) {
      case G753_MNT4_G1: rc = bases_generate_impl<0>(ctx, gen_xy, seed, n, b->d_points); break;
      case G753_MNT4_G2: rc = bases_generate_impl<1>(ctx, gen_xy, seed, n, b->d_points); break;
      case G753_MNT6_G1: rc = bases_generate_impl<2>(ctx, gen_xy, seed, n, b->d_points); break;
      case G753_MNT6_G2: rc = bases_generate_impl<3>(ctx, gen_xy, seed, n, b->d_points); break;
    }
  }
  if (rc != G753_OK) {
    dev_free(b->d_points);
    delete b;
    return rc;
  }
  *out = b;
  return G753_OK;
}

int g753_bases_precompute(g753_ctx* ctx, g753_bases* b, unsigned copies) {
  CHECK_CTX(ctx);
  if (!b) return fail(G753_ERR_BAD_ARG, "null bases");
  if (copies == 0) {
    // auto: as many copies as fit a per-key memory budget (default 6 GiB, G753_KEY_BUDGET_MB): the
    // set-up cost is ~W*c doublings per point whatever the count, only memory grows with it
    size_t budget_mb = 6144;
    const char* env = getenv("G753_KEY_BUDGET_MB");
    if (env && atol(env) > 0) budget_mb = (size_t)atol(env);
    const size_t key_bytes = (size_t)2 * group_k(b->group) * 96 * (b->n ? b->n : 1);
    size_t fit = (budget_mb << 20) / key_bytes;
    copies = fit < 2 ? 1 : fit > 64 ? 64 : (unsigned)fit;
  }
  if (copies > 64) return fail(G753_ERR_BAD_ARG, "too many copies");
  if (b->copies > 1 || copies == 1) return G753_OK;
  std::lock_guard<std::mutex> lock(ctx->mu);
  switch (b->group) {
    case G753_MNT4_G1: return bases_precompute_impl<0>(ctx, b, copies);
    case G753_MNT4_G2: return bases_precompute_impl<1>(ctx, b, copies);
    case G753_MNT6_G1: return bases_precompute_impl<2>(ctx, b, copies);
    case G753_MNT6_G2: return bases_precompute_impl<3>(ctx, b, copies);
  }
  return fail(G753_ERR_BAD_ARG, "unknown group");
}

int g753_bases_update(g753_ctx* ctx, g753_bases* b, size_t first, size_t count, const uint64_t* coords,
                      const uint8_t* infinity) {
  CHECK_CTX(ctx);
  if (!b || (count && !coords)) return fail(G753_ERR_BAD_ARG, "null pointer");
  if (first > b->n || count > b->n - first) return fail(G753_ERR_BAD_ARG, "bases slice out of range");
  if (b->copies > 1) return fail(G753_ERR_BAD_ARG, "key has precomputed copies");
  if (count == 0) return G753_OK;
  std::lock_guard<std::mutex> lock(ctx->mu);
  const int k = group_k(b->group);
  const size_t pt_bytes = (size_t)2 * k * 96;
  if (!b->d_inf) {
    G753_TRY(dev_alloc((void**)&b->d_inf, b->n));
    G753_TRY(dev_memset(b->d_inf, 0, b->n, ctx->stream));
  }
  G753_TRY(h2d((char*)b->d_points + first * pt_bytes, coords, count * pt_bytes, ctx->stream));
  if (infinity)
    G753_TRY(h2d(b->d_inf + first, infinity, count, ctx->stream));
  else
    G753_TRY(dev_memset(b->d_inf + first, 0, count, ctx->stream));
  G753_LAUNCH(k_bases_sanitize, div_up(count, 256), 256, ctx->stream, (Fq*)b->d_points + first * 2 * k,
              b->d_inf + first, (unsigned)count, (unsigned)(2 * k));
  ctx->launches++;
  G753_TRY(launch_check("k_bases_sanitize"));
  return stream_sync(ctx->stream);  // the host buffers may be reused as soon as this returns
}

int g753_bases_download(g753_ctx* ctx, const g753_bases* b, size_t first, size_t count, uint64_t* coords) {
  CHECK_CTX(ctx);
  if (!b || (count && !coords)) return fail(G753_ERR_BAD_ARG, "null pointer");
  if (first > b->n || count > b->n - first) return fail(G753_ERR_BAD_ARG, "bases slice out of range");
  const size_t pt_bytes = (size_t)2 * group_k(b->group) * 96;
  G753_TRY(d2h(coords, (const char*)b->d_points + first * pt_bytes, count * pt_bytes, ctx->stream));
  return stream_sync(ctx->stream);
}

static int msm_check(g753_ctx* ctx, const g753_bases* b, size_t first, size_t count) {
  if (!b) return fail(G753_ERR_BAD_ARG, "null bases");
  if (first > b->n || count > b->n - first) return fail(G753_ERR_BAD_ARG, "bases slice out of range");
  (void)ctx;
  return G753_OK;
}

#if !defined(G753_HOST_EMUL)
static void collect_phases(g753_ctx* ctx) {
  if (!ctx->ev_ok || !ctx->phases_valid) return;   // no MSM with count > 0 has run on this context yet
  for (int i = 0; i < MSM_PHASES; i++) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]) == cudaSuccess) ctx->phase_ms[i] = ms;
    else (void)cudaGetLastError();
  }
}
#endif

int g753_msm_dev(g753_ctx* ctx, const g753_bases* b, size_t first, size_t count, const void* d_scalars,
                 void* d_out_xyz) {
  CHECK_CTX(ctx);
  G753_TRY(msm_check(ctx, b, first, count));
  if ((count && !d_scalars) || !d_out_xyz) return fail(G753_ERR_BAD_ARG, "null pointer");
  std::lock_guard<std::mutex> lock(ctx->mu);
  return msm_any(ctx, b, first, count, (const uint32_t*)d_scalars, d_out_xyz);
}

int g753_msm(g753_ctx* ctx, const g753_bases* b, size_t first, size_t count, const uint64_t* scalars,
             uint64_t* out_xyz) {
  CHECK_CTX(ctx);
  G753_TRY(msm_check(ctx, b, first, count));
  if ((count && !scalars) || !out_xyz) return fail(G753_ERR_BAD_ARG, "null pointer");
  std::lock_guard<std::mutex> lock(ctx->mu);
  const size_t out_bytes = (size_t)3 * group_k(b->group) * 96;
  G753_TRY(ctx->scratch_io.reserve(count * 96 + out_bytes + 1024));
  Carver cv(ctx->scratch_io.ptr);
  uint32_t* d_scalars = cv.take<uint32_t>(count * NL);
  uint32_t* d_out = cv.take<uint32_t>(out_bytes / 4);
#if !defined(G753_HOST_EMUL)
  // upload the scalars in pieces on the copy stream; the digit extraction of piece j waits for
  // piece j only, so the transfer overlaps the start of the pipeline
  // Only for PINNED (page-locked / registered) caller memory: a cudaMemcpyAsync from pageable memory
  // blocks the host until the piece is staged, so nothing would overlap and the plain path is used.
  unsigned chunks = 1;
  if (ctx->copy_ok && count >= ((size_t)1 << 18)) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, scalars) == cudaSuccess && attr.type == cudaMemoryTypeHost)
      chunks = g753_ctx::MAX_CHUNKS;
    else
      (void)cudaGetLastError();
  }
  if (chunks > 1) {
    // the staging buffer is free once the work queued so far on the main stream is done
    cudaError_t e = cudaEventRecord(ctx->ready_ev, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ready_ev, 0);
    if (e != cudaSuccess) return cuda_fail(e, "g753_msm: ordering the scalar upload");
    for (unsigned j = 0; j < chunks; j++) {
      const size_t lo = count * j / chunks, hi = count * (j + 1) / chunks;
      G753_TRY(h2d((char*)d_scalars + lo * 96, (const char*)scalars + lo * 96, (hi - lo) * 96, ctx->copy_stream));
      e = cudaEventRecord(ctx->chunk_ev[j], ctx->copy_stream);
      if (e != cudaSuccess) return cuda_fail(e, "g753_msm: cudaEventRecord");
    }
    ctx->scalar_chunks = chunks;
    int rc = msm_any(ctx, b, first, count, d_scalars, d_out);
    ctx->scalar_chunks = 1;
    G753_TRY(rc);
  } else
#endif
  {
    if (count) G753_TRY(h2d(d_scalars, scalars, count * 96, ctx->stream));
    G753_TRY(msm_any(ctx, b, first, count, d_scalars, d_out));
  }
  G753_TRY(d2h(out_xyz, d_out, out_bytes, ctx->stream));
  G753_TRY(stream_sync(ctx->stream));
  return G753_OK;
}

int g753_msm_host(g753_ctx* ctx, int group, const uint64_t* coords, const uint8_t* infinity, size_t n_bases,
                  const uint64_t* scalars, size_t n_scalars, uint64_t* out_xyz) {
  CHECK_CTX(ctx);
  // zip-truncation of the reference: scalars.iter().zip(bases) (variable_base.rs:36)
  const size_t count = n_bases < n_scalars ? n_bases : n_scalars;
  g753_bases* b = nullptr;
  G753_TRY(g753_bases_upload(ctx, group, coords, infinity, count, &b));
  int rc = g753_msm(ctx, b, 0, count, scalars, out_xyz);
  g753_bases_free(ctx, b);
  return rc;
}

int g753_points_sum_dev(g753_ctx* ctx, int group, const void* d_points_xyz, size_t count, void* d_out_xyz) {
  CHECK_CTX(ctx);
  if (!d_points_xyz || !d_out_xyz || count > 0xffffffffull) return fail(G753_ERR_BAD_ARG, "bad argument");
  std::lock_guard<std::mutex> lock(ctx->mu);
  switch (group) {
    case G753_MNT4_G1: points_sum_launch<0>(ctx, d_points_xyz, count, d_out_xyz); break;
    case G753_MNT4_G2: points_sum_launch<1>(ctx, d_points_xyz, count, d_out_xyz); break;
    case G753_MNT6_G1: points_sum_launch<2>(ctx, d_points_xyz, count, d_out_xyz); break;
    case G753_MNT6_G2: points_sum_launch<3>(ctx, d_points_xyz, count, d_out_xyz); break;
    default: return fail(G753_ERR_BAD_ARG, "unknown group");
  }
  ctx->launches++;
  return launch_check("k_points_sum");
}

int g753_batch_normalize(g753_ctx* ctx, int group, const uint64_t* xyz, size_t count, uint64_t* xy,
                         uint8_t* infinity) {
  CHECK_CTX(ctx);
  if (count && (!xyz || !xy || !infinity)) return fail(G753_ERR_BAD_ARG, "null pointer");
  if (count > 0x7fffffffull) return fail(G753_ERR_BAD_ARG, "too many points");
  if (count == 0) return G753_OK;
  std::lock_guard<std::mutex> lock(ctx->mu);
  switch (group) {
    case G753_MNT4_G1: return batch_normalize_impl<0>(ctx, xyz, count, xy, infinity);
    case G753_MNT4_G2: return batch_normalize_impl<1>(ctx, xyz, count, xy, infinity);
    case G753_MNT6_G1: return batch_normalize_impl<2>(ctx, xyz, count, xy, infinity);
    case G753_MNT6_G2: return batch_normalize_impl<3>(ctx, xyz, count, xy, infinity);
  }
  return fail(G753_ERR_BAD_ARG, "unknown group");
}

int g753_fixed_base_msm(g753_ctx* ctx, int group, const uint64_t* base_xy, const uint64_t* scalars, size_t n,
                        uint64_t* out_xy, uint8_t* out_infinity) {
  CHECK_CTX(ctx);
  if (!base_xy || (n && (!scalars || !out_xy || !out_infinity))) return fail(G753_ERR_BAD_ARG, "null pointer");
  if (n > 0x7fffffffull) return fail(G753_ERR_BAD_ARG, "too many scalars");
  if (n == 0) return G753_OK;
  std::lock_guard<std::mutex> lock(ctx->mu);
  switch (group) {
    case G753_MNT4_G1: return fixed_base_impl<0>(ctx, base_xy, scalars, n, out_xy, out_infinity);
    case G753_MNT4_G2: return fixed_base_impl<1>(ctx, base_xy, scalars, n, out_xy, out_infinity);
    case G753_MNT6_G1: return fixed_base_impl<2>(ctx, base_xy, scalars, n, out_xy, out_infinity);
    case G753_MNT6_G2: return fixed_base_impl<3>(ctx, base_xy, scalars, n, out_xy, out_infinity);
  }
  return fail(G753_ERR_BAD_ARG, "unknown group");
}

// ---- NTT ---------------------------------------------------------------------------------
int g753_domain_check(int field, unsigned log_n) {
  if (field != 0 && field != 1) return fail(G753_ERR_BAD_ARG, "unknown field");
  if (log_n >= G753_FIELD_CONSTANTS[field].two_adicity || log_n > NTT_MAX_LOG)
    return fail(G753_ERR_DOMAIN, "domain too large for this field (EvaluationDomain::new -> None)");
  return G753_OK;
}

static int ntt_any(g753_ctx* ctx, int field, Fq* d_data, Fq* d_tmp, unsigned log_n, int mode) {
  if (mode < G753_FFT || mode > G753_COSET_IFFT) return fail(G753_ERR_BAD_ARG, "unknown transform");
  G753_TRY(g753_domain_check(field, log_n));
  return field == 0 ? ntt_field<0>(ctx, d_data, d_tmp, log_n, mode) : ntt_field<1>(ctx, d_data, d_tmp, log_n, mode);
}

int g753_ntt_dev(g753_ctx* ctx, int field, void* d_data, unsigned log_n, int mode) {
  CHECK_CTX(ctx);
  if (!d_data) return fail(G753_ERR_BAD_ARG, "null data");
  G753_TRY(g753_domain_check(field, log_n));
  std::lock_guard<std::mutex> lock(ctx->mu);
  const size_t bytes = sizeof(Fq) << log_n;
  G753_TRY(ctx->scratch_io.reserve(bytes + 1024));
  return ntt_any(ctx, field, (Fq*)d_data, (Fq*)ctx->scratch_io.ptr, log_n, mode);
}

int g753_ntt(g753_ctx* ctx, int field, uint64_t* data, unsigned log_n, int mode) {
  CHECK_CTX(ctx);
  if (!data) return fail(G753_ERR_BAD_ARG, "null data");
  G753_TRY(g753_domain_check(field, log_n));
  std::lock_guard<std::mutex> lock(ctx->mu);
  const size_t bytes = sizeof(Fq) << log_n;
  G753_TRY(ctx->scratch_io.reserve(2 * Carver::pad(bytes) + 1024));
  Carver cv(ctx->scratch_io.ptr);
  Fq* d_data = cv.take<Fq>((size_t)1 << log_n);
  Fq* d_tmp = cv.take<Fq>((size_t)1 << log_n);
  G753_TRY(h2d(d_data, data, bytes, ctx->stream));
  G753_TRY(ntt_any(ctx, field, d_data, d_tmp, log_n, mode));
  G753_TRY(d2h(data, d_data, bytes, ctx->stream));
  return stream_sync(ctx->stream);
}

int g753_witness_map_dev(g753_ctx* ctx, int field, void* d_a, void* d_b, void* d_c, unsigned log_n,
                         const uint64_t* d123_mont, void* d_h) {
  CHECK_CTX(ctx);
  if (!d_a || !d_b || !d_c || !d123_mont || !d_h) return fail(G753_ERR_BAD_ARG, "null pointer");
  G753_TRY(g753_domain_check(field, log_n));
  std::lock_guard<std::mutex> lock(ctx->mu);
  // d1, d2, d3 ride along as a kernel argument: no allocation, copy or synchronisation here - the
  // call only queues work on the context's stream
  Fq3 d;
  memcpy(&d, d123_mont, sizeof(d));
  return field == 0 ? witness_map_field<0>(ctx, (Fq*)d_a, (Fq*)d_b, (Fq*)d_c, log_n, d, (Fq*)d_h)
                    : witness_map_field<1>(ctx, (Fq*)d_a, (Fq*)d_b, (Fq*)d_c, log_n, d, (Fq*)d_h);
}

int g753_witness_map_tail_dev(g753_ctx* ctx, int field, void* d_a, const void* d_b, const void* d_c, unsigned log_n,
                              const uint64_t* d123_mont, void* d_h) {
  CHECK_CTX(ctx);
  if (!d_a || !d_b || !d_c || !d123_mont || !d_h) return fail(G753_ERR_BAD_ARG, "null pointer");
  G753_TRY(g753_domain_check(field, log_n));
  std::lock_guard<std::mutex> lock(ctx->mu);
  Fq3 d;
  memcpy(&d, d123_mont, sizeof(d));
  return field == 0 ? witness_map_tail_field<0>(ctx, (Fq*)d_a, (const Fq*)d_b, (const Fq*)d_c, log_n, d, (Fq*)d_h)
                    : witness_map_tail_field<1>(ctx, (Fq*)d_a, (const Fq*)d_b, (const Fq*)d_c, log_n, d, (Fq*)d_h);
}

int g753_witness_map(g753_ctx* ctx, int field, const uint64_t* a, const uint64_t* b, const uint64_t* c,
                     unsigned log_n, const uint64_t* d123_mont, uint64_t* h) {
  CHECK_CTX(ctx);
  if (!a || !b || !c || !d123_mont || !h) return fail(G753_ERR_BAD_ARG, "null pointer");
  G753_TRY(g753_domain_check(field, log_n));
  const size_t n = (size_t)1 << log_n, bytes = sizeof(Fq) * n;
  void *d_a = nullptr, *d_b = nullptr, *d_c = nullptr, *d_h = nullptr;
  int rc = dev_alloc(&d_a, bytes);
  if (rc == G753_OK) rc = dev_alloc(&d_b, bytes);
  if (rc == G753_OK) rc = dev_alloc(&d_c, bytes);
  if (rc == G753_OK) rc = dev_alloc(&d_h, bytes + sizeof(Fq));
  if (rc == G753_OK) rc = h2d(d_a, a, bytes, ctx->stream);
  if (rc == G753_OK) rc = h2d(d_b, b, bytes, ctx->stream);
  if (rc == G753_OK) rc = h2d(d_c, c, bytes, ctx->stream);
  if (rc == G753_OK) rc = g753_witness_map_dev(ctx, field, d_a, d_b, d_c, log_n, d123_mont, d_h);
  if (rc == G753_OK) rc = d2h(h, d_h, bytes + sizeof(Fq), ctx->stream);
  if (rc == G753_OK) rc = stream_sync(ctx->stream);
  dev_free(d_a);
  dev_free(d_b);
  dev_free(d_c);
  dev_free(d_h);
  return rc;
}

int g753_ntt_shard_create(g753_ctx* ctx, int field, unsigned log_n, unsigned world, unsigned rank,
                          g753_ntt_shard** out) {
  CHECK_CTX(ctx);
  if (!out) return fail(G753_ERR_BAD_ARG, "null out");
  *out = nullptr;
  G753_TRY(g753_domain_check(field, log_n));
  if (world == 0 || (world & (world - 1)) || rank >= world) return fail(G753_ERR_BAD_ARG, "world must be a power of two");
  unsigned log_w = 0;
  while ((1u << log_w) < world) log_w++;
  const unsigned log_n2 = log_n / 2, log_n1 = log_n - log_n2;   // n1 >= n2
  if (log_n2 < log_w) return fail(G753_ERR_BAD_ARG, "domain too small to shard over this many ranks");
  std::lock_guard<std::mutex> lock(ctx->mu);
  g753_ntt_shard* p = new (std::nothrow) g753_ntt_shard();
  if (!p) return fail(G753_ERR_OOM, "host allocation failed");
  p->field = field;
  p->log_n = log_n;
  p->log_n1 = log_n1;
  p->log_n2 = log_n2;
  p->world = world;
  p->rank = rank;
  p->n1 = (size_t)1 << log_n1;
  p->n2 = (size_t)1 << log_n2;
  p->cols = p->n2 / world;
  p->rows = p->n1 / world;
  int rc = field == 0 ? shard_build<0>(ctx, p) : shard_build<1>(ctx, p);
  if (rc != G753_OK) {
    p->release();
    delete p;
    return rc;
  }
  *out = p;
  return G753_OK;
}

int g753_ntt_shard_destroy(g753_ctx* ctx, g753_ntt_shard* plan) {
  if (!plan) return G753_OK;
  if (ctx) {
    use_device(ctx);
    stream_sync(ctx->stream);
  }
  plan->release();
  delete plan;
  return G753_OK;
}

int g753_ntt_shard_shape(const g753_ntt_shard* plan, size_t* n1, size_t* n2, size_t* cols, size_t* rows) {
  if (!plan) return fail(G753_ERR_BAD_ARG, "null plan");
  if (n1) *n1 = plan->n1;
  if (n2) *n2 = plan->n2;
  if (cols) *cols = plan->cols;
  if (rows) *rows = plan->rows;
  return G753_OK;
}

int g753_ntt_shard_step1(g753_ctx* ctx, const g753_ntt_shard* plan, void* d_data, void* d_send, int mode) {
  CHECK_CTX(ctx);
  if (!plan || !d_data || !d_send) return fail(G753_ERR_BAD_ARG, "null pointer");
  if (mode < G753_FFT || mode > G753_COSET_IFFT) return fail(G753_ERR_BAD_ARG, "unknown transform");
  std::lock_guard<std::mutex> lock(ctx->mu);
  return plan->field == 0 ? shard_step1<0>(ctx, plan, (Fq*)d_data, (Fq*)d_send, mode)
                          : shard_step1<1>(ctx, plan, (Fq*)d_data, (Fq*)d_send, mode);
}

int g753_ntt_shard_step1_fused(g753_ctx* ctx, const g753_ntt_shard* plan, void* d_data, void* const* peer_z,
                               int mode) {
  CHECK_CTX(ctx);
  if (!plan || !d_data || !peer_z) return fail(G753_ERR_BAD_ARG, "null pointer");
  if (mode < G753_FFT || mode > G753_COSET_IFFT) return fail(G753_ERR_BAD_ARG, "unknown transform");
#if defined(G753_HOST_EMUL)
  return fail(G753_ERR_NO_DEVICE, "peer-memory exchange needs GPUs");
#else
  if (plan->world > 8 || plan->log_n1 == 0) return fail(G753_ERR_BAD_ARG, "fused exchange: 2..8 ranks, n1 >= 2");
  for (unsigned h = 0; h < plan->world; h++)
    if (!peer_z[h]) return fail(G753_ERR_BAD_ARG, "null peer buffer");
  std::lock_guard<std::mutex> lock(ctx->mu);
  return plan->field == 0 ? shard_step1_fused<0>(ctx, plan, (Fq*)d_data, peer_z, mode)
                          : shard_step1_fused<1>(ctx, plan, (Fq*)d_data, peer_z, mode);
#endif
}

int g753_ntt_shard_step2_local(g753_ctx* ctx, const g753_ntt_shard* plan, void* d_z, int mode) {
  CHECK_CTX(ctx);
  if (!plan || !d_z) return fail(G753_ERR_BAD_ARG, "null pointer");
  if (mode < G753_FFT || mode > G753_COSET_IFFT) return fail(G753_ERR_BAD_ARG, "unknown transform");
  std::lock_guard<std::mutex> lock(ctx->mu);
  return plan->field == 0 ? shard_step2_local<0>(ctx, plan, (Fq*)d_z, mode) : shard_step2_local<1>(ctx, plan, (Fq*)d_z, mode);
}

int g753_ntt_shard_step2(g753_ctx* ctx, const g753_ntt_shard* plan, const void* d_recv, void* d_data, int mode) {
  CHECK_CTX(ctx);
  if (!plan || !d_data || !d_recv) return fail(G753_ERR_BAD_ARG, "null pointer");
  if (mode < G753_FFT || mode > G753_COSET_IFFT) return fail(G753_ERR_BAD_ARG, "unknown transform");
  std::lock_guard<std::mutex> lock(ctx->mu);
  return plan->field == 0 ? shard_step2<0>(ctx, plan, (const Fq*)d_recv, (Fq*)d_data, mode)
                          : shard_step2<1>(ctx, plan, (const Fq*)d_recv, (Fq*)d_data, mode);
}

int g753_domain_check_mixed(int field, uint64_t n) {
  if (field != 0 && field != 1) return fail(G753_ERR_BAD_ARG, "unknown field");
  if (!mixed_split(field, n, nullptr, nullptr, nullptr))
    return fail(G753_ERR_DOMAIN, "no mixed-radix domain of this size (needs n = 2^a * m | p - 1, m odd <= 1024)");
  return G753_OK;
}

int g753_ntt_mixed_dev(g753_ctx* ctx, int field, void* d_data, uint64_t n, int mode) {
  CHECK_CTX(ctx);
  if (!d_data) return fail(G753_ERR_BAD_ARG, "null data");
  if (mode < G753_FFT || mode > G753_COSET_IFFT) return fail(G753_ERR_BAD_ARG, "unknown transform");
  G753_TRY(g753_domain_check_mixed(field, n));
  std::lock_guard<std::mutex> lock(ctx->mu);
  G753_TRY(ctx->scratch_io.reserve(sizeof(Fq) * n + 1024));
  return field == 0 ? mixed_run<0>(ctx, (Fq*)d_data, (Fq*)ctx->scratch_io.ptr, n, mode)
                    : mixed_run<1>(ctx, (Fq*)d_data, (Fq*)ctx->scratch_io.ptr, n, mode);
}

int g753_ntt_mixed(g753_ctx* ctx, int field, uint64_t* data, uint64_t n, int mode) {
  CHECK_CTX(ctx);
  if (!data) return fail(G753_ERR_BAD_ARG, "null data");
  G753_TRY(g753_domain_check_mixed(field, n));
  void* d = nullptr;
  G753_TRY(dev_alloc(&d, sizeof(Fq) * n));
  int rc = h2d(d, data, sizeof(Fq) * n, ctx->stream);
  if (rc == G753_OK) rc = g753_ntt_mixed_dev(ctx, field, d, n, mode);
  if (rc == G753_OK) rc = d2h(data, d, sizeof(Fq) * n, ctx->stream);
  if (rc == G753_OK) rc = stream_sync(ctx->stream);
  dev_free(d);
  return rc;
}

int g753_domain_constant(g753_ctx* ctx, int field, unsigned log_n, int which, uint64_t* out) {
  CHECK_CTX(ctx);
  if (!out || which < 0 || which > 4) return fail(G753_ERR_BAD_ARG, "bad argument");
  G753_TRY(g753_domain_check(field, log_n));
  std::lock_guard<std::mutex> lock(ctx->mu);
  // the constants alone: no O(n) twiddle tables for a domain that may never be transformed
  Fq* consts = nullptr;
  int rc = field == 0 ? ntt_consts_build<0>(&consts, log_n, ctx->stream, &ctx->launches)
                      : ntt_consts_build<1>(&consts, log_n, ctx->stream, &ctx->launches);
  if (rc == G753_OK) {
    const unsigned L = log_n ? log_n : 1;
    // consts layout: see k_ntt_setup
    const unsigned idx[5] = {0u, 1u, 1u + L, 1u + 3 * L, 1u + 4 * L};
    rc = d2h(out, consts + idx[which], sizeof(Fq), ctx->stream);
    if (rc == G753_OK) rc = stream_sync(ctx->stream);
  }
  dev_free(consts);
  return rc;
}

int g753_vec_op_dev(g753_ctx* ctx, int field, int op, void* d_a, const void* d_b, size_t n) {
  CHECK_CTX(ctx);
  if (!d_a || (field != 0 && field != 1)) return fail(G753_ERR_BAD_ARG, "bad argument");
  if (n == 0) return G753_OK;
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (field == 0)
    G753_LAUNCH(k_vec_op<0>, div_up(n, 256), 256, ctx->stream, (Fq*)d_a, (const Fq*)d_b, op, n);
  else
    G753_LAUNCH(k_vec_op<1>, div_up(n, 256), 256, ctx->stream, (Fq*)d_a, (const Fq*)d_b, op, n);
  ctx->launches++;
  return launch_check("k_vec_op");
}

int g753_vec_scale_dev(g753_ctx* ctx, int field, void* d_a, const uint64_t* k_mont, size_t n) {
  CHECK_CTX(ctx);
  if (!d_a || !k_mont || (field != 0 && field != 1)) return fail(G753_ERR_BAD_ARG, "bad argument");
  if (n == 0) return G753_OK;
  std::lock_guard<std::mutex> lock(ctx->mu);
  // the factor is a kernel argument (96 bytes by value): nothing is allocated, copied or waited for
  Fq k;
  memcpy(&k, k_mont, sizeof(k));
  if (field == 0)
    G753_LAUNCH(k_vec_scale_val<0>, div_up(n, 256), 256, ctx->stream, (Fq*)d_a, k, n);
  else
    G753_LAUNCH(k_vec_scale_val<1>, div_up(n, 256), 256, ctx->stream, (Fq*)d_a, k, n);
  ctx->launches++;
  return launch_check("k_vec_scale");
}

// ---- plumbing ------------------------------------------------------------------------------
int g753_dev_alloc(g753_ctx* ctx, size_t bytes, void** d_ptr) {
  CHECK_CTX(ctx);
  if (!d_ptr) return fail(G753_ERR_BAD_ARG, "null out");
  return dev_alloc(d_ptr, bytes);
}
int g753_dev_free(g753_ctx* ctx, void* d_ptr) {
  CHECK_CTX(ctx);
  stream_sync(ctx->stream);
  dev_free(d_ptr);
  return G753_OK;
}
int g753_h2d(g753_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
  CHECK_CTX(ctx);
  G753_TRY(h2d(d_dst, h_src, bytes, ctx->stream));
  return stream_sync(ctx->stream);
}
int g753_d2h(g753_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
  CHECK_CTX(ctx);
  G753_TRY(d2h(h_dst, d_src, bytes, ctx->stream));
  return stream_sync(ctx->stream);
}
int g753_d2d(g753_ctx* ctx, void* d_dst, const void* d_src, size_t bytes) {
  CHECK_CTX(ctx);
  return d2d(d_dst, d_src, bytes, ctx->stream);
}
int g753_sync(g753_ctx* ctx) {
  CHECK_CTX(ctx);
  return stream_sync(ctx->stream);
}
int g753_ctx_wait(g753_ctx* ctx, g753_ctx* other) {
  CHECK_CTX(ctx);
  if (!other) return fail(G753_ERR_BAD_ARG, "null context");
#if !defined(G753_HOST_EMUL)
  if (other->device != ctx->device) return fail(G753_ERR_BAD_ARG, "contexts live on different devices");
  cudaEvent_t ev;
  cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  if (e != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
  e = cudaEventRecord(ev, other->stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ev, 0);
  cudaEventDestroy(ev);  // released once the wait has been satisfied
  if (e != cudaSuccess) return cuda_fail(e, "cudaStreamWaitEvent");
#endif
  return G753_OK;
}
void* g753_stream(g753_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// ---- introspection -------------------------------------------------------------------------
int g753_field_op(g753_ctx* ctx, int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out,
                  size_t n) {
  CHECK_CTX(ctx);
  if (!a || !out || (field != 0 && field != 1)) return fail(G753_ERR_BAD_ARG, "bad argument");
  if (n == 0) return G753_OK;
  void *d_a = nullptr, *d_b = nullptr;
  int rc = dev_alloc(&d_a, n * 96);
  if (rc == G753_OK && b) rc = dev_alloc(&d_b, n * 96);
  if (rc == G753_OK) rc = h2d(d_a, a, n * 96, ctx->stream);
  if (rc == G753_OK && b) rc = h2d(d_b, b, n * 96, ctx->stream);
  if (rc == G753_OK) rc = g753_vec_op_dev(ctx, field, op, d_a, d_b, n);
  if (rc == G753_OK) rc = d2h(out, d_a, n * 96, ctx->stream);
  if (rc == G753_OK) rc = stream_sync(ctx->stream);
  dev_free(d_a);
  dev_free(d_b);
  return rc;
}

int g753_point_op(g753_ctx* ctx, int group, int op, const uint64_t* a, const uint64_t* b, uint64_t* out_xyz) {
  CHECK_CTX(ctx);
  if (!a || !out_xyz || (op != 1 && !b)) return fail(G753_ERR_BAD_ARG, "null pointer");
  std::lock_guard<std::mutex> lock(ctx->mu);
  switch (group) {
    case G753_MNT4_G1: return point_op_impl<0>(ctx, op, a, b, out_xyz);
    case G753_MNT4_G2: return point_op_impl<1>(ctx, op, a, b, out_xyz);
    case G753_MNT6_G1: return point_op_impl<2>(ctx, op, a, b, out_xyz);
    case G753_MNT6_G2: return point_op_impl<3>(ctx, op, a, b, out_xyz);
  }
  return fail(G753_ERR_BAD_ARG, "unknown group");
}

int g753_ext_op(g753_ctx* ctx, int group, int lanes, int op, const uint64_t* a, const uint64_t* b, uint64_t* out,
                size_t n) {
  CHECK_CTX(ctx);
  if (!a || !out || (lanes != 0 && lanes != 1)) return fail(G753_ERR_BAD_ARG, "bad argument");
  if (n == 0) return G753_OK;
  if (n > ((size_t)1 << 24)) return fail(G753_ERR_BAD_ARG, "too many elements");
  std::lock_guard<std::mutex> lock(ctx->mu);
  switch (group) {
    case G753_MNT4_G1: return ext_op_impl<0>(ctx, lanes, op, a, b, out, n);
    case G753_MNT4_G2: return ext_op_impl<1>(ctx, lanes, op, a, b, out, n);
    case G753_MNT6_G1: return ext_op_impl<2>(ctx, lanes, op, a, b, out, n);
    case G753_MNT6_G2: return ext_op_impl<3>(ctx, lanes, op, a, b, out, n);
  }
  return fail(G753_ERR_BAD_ARG, "unknown group");
}

int g753_coop_op(g753_ctx* ctx, int field, int op, unsigned k, const uint64_t* a, const uint64_t* b, uint64_t* out,
                 size_t n) {
  CHECK_CTX(ctx);
  if (!a || !b || !out || (field != 0 && field != 1) || op < 0 || op > 4 || k > 11) return fail(G753_ERR_BAD_ARG, "bad argument");
  if (n == 0) return G753_OK;
#if defined(G753_HOST_EMUL)
  return fail(G753_ERR_NO_DEVICE, "the warp-cooperative arithmetic runs on a GPU only");
#else
  std::lock_guard<std::mutex> lock(ctx->mu);
  const size_t bytes = sizeof(Fq) * n;
  G753_TRY(ctx->scratch_io.reserve(3 * Carver::pad(bytes) + 1024));
  Carver cv(ctx->scratch_io.ptr);
  Fq* da = cv.take<Fq>(n);
  Fq* db = cv.take<Fq>(n);
  Fq* dout = cv.take<Fq>(n);
  G753_TRY(h2d(da, a, bytes, ctx->stream));
  G753_TRY(h2d(db, b, bytes, ctx->stream));
  const size_t smem = sizeof(uint32_t) * (COOP_CONST_WORDS + 12 * 24);
  if (field == 0) k_coop_op<0><<<1, 32, smem, ctx->stream>>>(op, k, da, db, (unsigned)n, dout);
  else k_coop_op<1><<<1, 32, smem, ctx->stream>>>(op, k, da, db, (unsigned)n, dout);
  ctx->launches++;
  G753_TRY(launch_check("k_coop_op"));
  G753_TRY(d2h(out, dout, bytes, ctx->stream));
  return stream_sync(ctx->stream);
#endif
}

int g753_mac_probe(g753_ctx* ctx, int variant, int blocks, int threads, int iters, float* ms) {
  CHECK_CTX(ctx);
  if (!ms || blocks <= 0 || threads <= 0 || threads > 256 || iters <= 0) return fail(G753_ERR_BAD_ARG, "bad argument");
#if defined(G753_HOST_EMUL)
  (void)variant;
  return fail(G753_ERR_NO_DEVICE, "probe needs a GPU");
#else
  std::lock_guard<std::mutex> lock(ctx->mu);
  Fq* d_seed = nullptr;
  G753_TRY(dev_alloc((void**)&d_seed, sizeof(Fq) * 512));
  std::vector<uint32_t> h(256 * NL);
  uint64_t s = 0x9E3779B97F4A7C15ull;
  for (size_t i = 0; i < h.size(); i++) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    h[i] = (uint32_t)(s >> 33);
    if (i % NL == NL - 1) h[i] &= 0xffff;  // < p
  }
  int rc = h2d(d_seed, h.data(), h.size() * 4, ctx->stream);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  if (rc == G753_OK) {
    cudaEventRecord(e0, ctx->stream);
    if (variant >= 5 && variant <= 8) {
      k_coop_probe<<<1, 32, coop_smem_bytes<0>(1), ctx->stream>>>(variant, d_seed, d_seed + 256, iters);
    } else if (variant == 4) {
      // 256 doubles below 2^52 (exact integers) over the same seed buffer
      std::vector<double> hd(256);
      for (int i = 0; i < 256; i++) hd[i] = (double)((((uint64_t)h[2 * i] << 32) | h[2 * i + 1]) & 0xFFFFFFFFFFFFFull);
      rc = h2d(d_seed, hd.data(), hd.size() * sizeof(double), ctx->stream);
      if (rc == G753_OK) rc = stream_sync(ctx->stream);
      cudaEventRecord(e0, ctx->stream);
      k_dfma_probe<<<blocks, threads, 0, ctx->stream>>>((const double*)d_seed, (double*)(d_seed + 256), iters);
    } else if (variant == 0) k_mac_probe<0><<<blocks, threads, 0, ctx->stream>>>(d_seed, d_seed + 256, iters);
    else if (variant == 1) k_mac_probe<1><<<blocks, threads, 0, ctx->stream>>>(d_seed, d_seed + 256, iters);
    else if (variant == 3) k_mac_probe<3><<<blocks, threads, 0, ctx->stream>>>(d_seed, d_seed + 256, iters);
    else k_mac_probe<2><<<blocks, threads, 0, ctx->stream>>>(d_seed, d_seed + 256, iters);
    cudaEventRecord(e1, ctx->stream);
    ctx->launches++;
    rc = launch_check("k_mac_probe");
  }
  if (rc == G753_OK) rc = stream_sync(ctx->stream);
  if (rc == G753_OK) cudaEventElapsedTime(ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  dev_free(d_seed);
  return rc;
#endif
}

int g753_debug_scratch(g753_ctx* ctx, void* h_dst, size_t bytes, size_t* cap) {
  CHECK_CTX(ctx);
  if (cap) *cap = ctx->scratch.cap;
  if (!h_dst || bytes == 0) return G753_OK;
  if (bytes > ctx->scratch.cap) bytes = ctx->scratch.cap;
  G753_TRY(d2h(h_dst, ctx->scratch.ptr, bytes, ctx->stream));
  return stream_sync(ctx->stream);
}

uint64_t g753_launch_count(const g753_ctx* ctx) { return ctx ? ctx->launches : 0; }

int g753_last_msm_phases(g753_ctx* ctx, float* ms, int cap) {
  if (!ctx || !ms) return 0;
#if !defined(G753_HOST_EMUL)
  if (use_device(ctx) == G753_OK && stream_sync(ctx->stream) == G753_OK) collect_phases(ctx);
#endif
  int k = cap < MSM_PHASES ? cap : MSM_PHASES;
  for (int i = 0; i < k; i++) ms[i] = ctx->phase_ms[i];
  return k;
}

int g753_last_msm_plan(const g753_ctx* ctx, unsigned* plan5) {
  if (!ctx || !plan5) return fail(G753_ERR_BAD_ARG, "null pointer");
  for (int i = 0; i < 5; i++) plan5[i] = ctx->last_plan[i];
  return G753_OK;
}

}  // extern "C"
