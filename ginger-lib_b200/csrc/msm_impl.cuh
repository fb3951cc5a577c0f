// Per-group (curve) instantiations of the MSM pipeline and the group-law test hook.
// Compiled once per group (msm_g0.cu .. msm_g3.cu) so the four instantiations build in parallel.
#pragma once
#include "ctx.cuh"
#include "ec_slots.cuh"
#include "msm.cuh"

using namespace g753;

template <int GID>
int msm_dispatch(g753_ctx* ctx, const g753_bases* b, size_t first, size_t count,
                        const uint32_t* d_scalars, void* d_out) {
  MsmHooks hooks;
  hooks.launches = &ctx->launches;
#if !defined(G753_HOST_EMUL)
  hooks.mark = phase_mark;
  hooks.wait_chunk = chunk_wait;
  hooks.scalar_chunks = ctx->scalar_chunks;
  hooks.user = ctx;
#endif
  MsmKey key;
  key.bases = (const Fq*)b->d_points + first * 2 * MsmCfg<GID>::K;
  key.inf = b->d_inf ? b->d_inf + first : nullptr;
  key.copy_stride = key.inf_stride = b->n;
  key.affine = ctx->forced_affine;
  key.tree_batch = ctx->tree_batch;
  key.tree_waves = ctx->tree_waves;
  // the precomputed copies pay off when the slice is a sizeable part of the key they were sized
  // for; short slices (the prover's input-query views) run the plain pipeline on copy 0
  if (b->copies > 1 && count * 4 >= b->n) {
    key.copies = b->copies;
    key.c = (int)b->c;
    key.rows = b->rows;
  } else {
    key.copies = 1;
    key.c = ctx->forced_c;
  }
  if (count) {
    const MsmPlan pl = msm_plan(msm_cost<GID>(), count, key.copies, key.c, key.rows);
    ctx->last_plan[0] = pl.c;
    ctx->last_plan[1] = pl.W;
    ctx->last_plan[2] = pl.rows;
    ctx->last_plan[3] = pl.copies;
  }
  hooks.form = &ctx->last_plan[4];
  return msm_run<GID>(ctx->scratch, ctx->stream, key, d_scalars, count, (Fq*)d_out, hooks);
}

// group-law test hook on the product's own slot code (ec_slots.cuh); affine (0, 0) = infinity
template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS)
k_point_op(int op, const Fq* a, const Fq* b, const uint32_t* scalar, Fq* out) {
  typedef EcS<SC> E;
  typedef typename E::M M;
  constexpr int K = E::K, P = 0, Q = E::PT, S = 2 * E::PT;
  if (E::M::T::item() != 0) return;
  auto load_affine = [&](int D, const Fq* q) {
    E::set_inf(D);
    M::ldg(S, q);
    M::ldg(S + K, q + K);
    if (!(M::is_zero(S) && M::is_zero(S + K))) E::madd_g(D, q, false, S);
  };
  if (op == 0) {
    load_affine(P, a);
    M::ldg(S, b);
    M::ldg(S + K, b + K);
    if (!(M::is_zero(S) && M::is_zero(S + K))) E::madd_g(P, b, false, S);
  } else if (op == 1) {
    load_affine(P, a);
    E::dbl(P, S);
  } else if (op == 3) {  // 2a + 2b through the full (XYZZ + XYZZ) addition
    load_affine(P, a);
    E::dbl(P, S);
    load_affine(Q, b);
    E::dbl(Q, S);
    E::add(P, Q, S);
  } else {  // MSB-first double-and-add, as the reference's mul_assign
    E::set_inf(P);
    M::ldg(S, a);
    M::ldg(S + K, a + K);
    const bool a_inf = M::is_zero(S) && M::is_zero(S + K);
    bool started = false;
    for (int i = NL * 32 - 1; i >= 0 && !a_inf; i--) {
      if (started) E::dbl(P, S);
      if ((scalar[i >> 5] >> (i & 31)) & 1) {
        E::madd_g(P, a, false, S);
        started = true;
      }
    }
  }
  E::to_projective(P, S);
  for (int i = 0; i < 3; i++) M::stg(out + i * K, P + i * K);
}

template <int GID>
int point_op_impl(g753_ctx* ctx, int op, const uint64_t* a, const uint64_t* b, uint64_t* out) {
  constexpr int K = MsmCfg<GID>::K, T = 32;
  typedef typename MsmCfg<GID>::template SC<T> SC;
  typedef EcS<SC> E;
  const size_t aff = sizeof(Fq) * 2 * K, prj = sizeof(Fq) * 3 * K;
  G753_TRY(ctx->scratch_io.reserve(2 * aff + 96 + prj + 1024));
  Carver cv(ctx->scratch_io.ptr);
  Fq* da = cv.take<Fq>(2 * K);
  Fq* db = cv.take<Fq>(2 * K);
  uint32_t* ds = cv.take<uint32_t>(NL);
  Fq* dout = cv.take<Fq>(3 * K);
  G753_TRY(h2d(da, a, aff, ctx->stream));
  if (op == 0 || op == 3) G753_TRY(h2d(db, b, aff, ctx->stream));
  if (op == 2) G753_TRY(h2d(ds, b, 96, ctx->stream));
  G753_LAUNCH_SMEM(k_point_op<SC>, 1, T * MsmCfg<GID>::TP, (slot_bytes<E, T>(2 * E::PT + E::ADD_SCRATCH)), ctx->stream, op, da, db,
                   ds, dout);
  ctx->launches++;
  G753_TRY(launch_check("k_point_op"));
  G753_TRY(d2h(out, dout, prj, ctx->stream));
  return stream_sync(ctx->stream);
}

// Tower-arithmetic test hook: the coordinate field of the group (Fq, Fq2 or Fq3) through the very
// tower code the kernels run - on the device the lane-cooperative Tw2C / Tw3C (slots.cuh) in the lane
// counts of the accumulation kernels (lanes = 0) or of the reduction kernels (lanes = 1).  One column
// per element pair.  ops: 0 mul, 1 add, 2 sub, 3 sqr, 4 neg, 5 inv, 12 dbl, 20 / 21 mul with the
// destination aliasing a / b, 23 sqr in place (the in-place forms the curve formulas rely on).
template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS)
k_ext_op(int op, const Fq* __restrict__ a, const Fq* __restrict__ b, unsigned n, Fq* __restrict__ out) {
  typedef typename SC::M M;
  constexpr int K = M::K, A = 0, B = K, D = 2 * K, TMP = 3 * K;
  if (M::T::idle()) return;
  unsigned i = M::T::item();
  if (i >= n) return;
  M::ldg(A, a + (size_t)i * K);
  M::ldg(B, (b ? b : a) + (size_t)i * K);
  int res = D;
  switch (op) {
    case 0: M::mul(D, A, B, TMP); break;
    case 1: M::add(D, A, B); break;
    case 2: M::sub(D, A, B); break;
    case 3: M::sqr(D, A, TMP); break;
    case 4: M::neg(D, A); break;
    case 5: M::inv(D, A, TMP); break;
    case 12: M::dbl(D, A); break;
    case 20: M::mul(A, A, B, TMP); res = A; break;
    case 21: M::mul(B, A, B, TMP); res = B; break;
    case 23: M::sqr(A, A, TMP); res = A; break;
    default: M::set_zero(D);
  }
  M::stg(out + (size_t)i * K, res);
}

template <int GID>
int ext_op_impl(g753_ctx* ctx, int lanes, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  typedef MsmCfg<GID> Cfg;
  constexpr int K = Cfg::K, T = 32;
  const size_t bytes = sizeof(Fq) * K * n;
  G753_TRY(ctx->scratch_io.reserve(3 * Carver::pad(bytes) + 1024));
  Carver cv(ctx->scratch_io.ptr);
  Fq* da = cv.take<Fq>(K * n);
  Fq* db = cv.take<Fq>(K * n);
  Fq* dout = cv.take<Fq>(K * n);
  G753_TRY(h2d(da, a, bytes, ctx->stream));
  if (b) G753_TRY(h2d(db, b, bytes, ctx->stream));
  if (lanes == 0) {
    constexpr int TA = Cfg::TPA == 3 ? 40 : T;      // three lanes per column: ten columns per warp
    typedef typename Cfg::template SC<TA, Cfg::TPA> SC;
    G753_LAUNCH_SMEM(k_ext_op<SC>, div_up(n, TA), (Lay<TA, Cfg::TPA>::THREADS), (slot_bytes<EcS<SC>, TA>(3 * K + SC::M::NTMP + 1)),
                     ctx->stream, op, da, b ? db : (const Fq*)nullptr, (unsigned)n, dout);
  } else {
    typedef typename Cfg::template SC<T, Cfg::TP> SC;
    G753_LAUNCH_SMEM(k_ext_op<SC>, div_up(n, T), T * Cfg::TP, (slot_bytes<EcS<SC>, T>(3 * K + SC::M::NTMP + 1)), ctx->stream, op, da,
                     b ? db : (const Fq*)nullptr, (unsigned)n, dout);
  }
  ctx->launches++;
  G753_TRY(launch_check("k_ext_op"));
  G753_TRY(d2h(out, dout, bytes, ctx->stream));
  return stream_sync(ctx->stream);
}

template <int GID>
void points_sum_launch(g753_ctx* ctx, const void* d_pts, size_t count, void* d_out) {
#if defined(G753_HOST_EMUL)
  constexpr int T = 32;  // columns; only column 0 works
  typedef typename MsmCfg<GID>::template SC<T> SC;
  typedef EcS<SC> E;
  G753_LAUNCH_SMEM(k_points_sum<SC>, 1, T * MsmCfg<GID>::TP, (slot_bytes<E, T>(2 * E::PT + E::ADD_SCRATCH)), ctx->stream,
                   (const Fq*)d_pts, (unsigned)count, (Fq*)d_out);
#else
  // one warp, four field products at a time (coop.cuh): the fold is a latency chain
  G753_LAUNCH_SMEM(k_points_sum_coop<GID>, 1, 32, coop_smem_bytes<GID>(1), ctx->stream, (const Fq*)d_pts, (unsigned)count,
                   (Fq*)d_out);
#endif
}


// Synthetic proving-key bases for benchmarks and full-size parity checks:
// P_i = a_i * G with a_i = splitmix64(seed, i) | 1, normalised to affine.  The discrete logs a_i
// are reproducible on the host, so sum_i s_i P_i can be checked at any size as
// (sum_i s_i a_i mod r) * G.
G753_HD uint64_t splitmix64_at(uint64_t seed, uint64_t i) {
  uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS)
k_bases_generate(const Fq* __restrict__ gen, uint64_t seed, unsigned n, Fq* __restrict__ out) {
  typedef EcS<SC> E;
  typedef typename E::M M;
  constexpr int K = E::K, P = 0, S = E::PT;
  unsigned i = E::M::T::item();
  if (i >= n) return;
  const uint64_t a = splitmix64_at(seed, i) | 1ull;
  E::set_inf(P);
  bool started = false;
  for (int b = 63; b >= 0; b--) {
    if (started) E::dbl(P, S);
    if ((a >> b) & 1) {
      E::madd_g(P, gen, false, S);
      started = true;
    }
  }
  // a_i < 2^64 << r and G has prime order r: the result is never the point at infinity
  E::to_affine(P, S);
  M::stg(out + (size_t)i * 2 * K, P);
  M::stg(out + (size_t)i * 2 * K + K, P + K);
}

template <int GID>
int bases_generate_impl(g753_ctx* ctx, const uint64_t* gen_xy, uint64_t seed, size_t n, void* d_points) {
  constexpr int K = MsmCfg<GID>::K, T = 64;
  typedef typename MsmCfg<GID>::template SC<T> SC;
  typedef EcS<SC> E;
  if (n == 0) return G753_OK;
  Fq* d_gen = nullptr;
  G753_TRY(dev_alloc((void**)&d_gen, sizeof(Fq) * 2 * K));
  int rc = h2d(d_gen, gen_xy, sizeof(Fq) * 2 * K, ctx->stream);
  if (rc == G753_OK) {
    G753_LAUNCH_SMEM(k_bases_generate<SC>, div_up(n, T), T * MsmCfg<GID>::TP, (slot_bytes<E, T>(E::PT + E::ADD_SCRATCH)), ctx->stream,
                     d_gen, seed, (unsigned)n, (Fq*)d_points);
    ctx->launches++;
    rc = launch_check("k_bases_generate");
  }
  if (rc == G753_OK) rc = stream_sync(ctx->stream);
  dev_free(d_gen);
  return rc;
}

// Precomputed key copies (g753_bases_precompute): copy j of point i is 2^(j * shift) * P_i in
// affine form, shift = rows * c.  Window w = j * rows + r of a scalar then selects a bucket of
// row r with the base taken from copy j, so the MSM has `rows` bucket rows instead of W: the
// bucket reduction and the serial Horner fold shrink by W / rows.
template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS)
k_bases_precompute(Fq* __restrict__ pts, uint8_t* __restrict__ inf, unsigned n, unsigned copies, unsigned shift) {
  typedef EcS<SC> E;
  typedef typename E::M M;
  constexpr int K = E::K, P = 0, S = E::PT;
  unsigned i = E::M::T::item();
  if (i >= n) return;
  bool is_inf = inf[i] != 0;
  if (!is_inf) {
    M::ldg(P, pts + (size_t)i * 2 * K);
    M::ldg(P + K, pts + (size_t)i * 2 * K + K);
    M::set_one(P + 2 * K);
    M::set_one(P + 3 * K);
  }
  for (unsigned j = 1; j < copies; j++) {
    Fq* out = pts + ((size_t)j * n + i) * 2 * K;
    if (!is_inf) {
      for (unsigned k = 0; k < shift; k++) E::dbl(P, S);
      is_inf = E::is_inf(P);
    }
    if (is_inf) {
      M::set_zero(S);
      M::stg(out, S);
      M::stg(out + K, S);
      inf[(size_t)j * n + i] = 1;
    } else {
      E::to_affine(P, S);
      M::stg(out, P);
      M::stg(out + K, P + K);
      inf[(size_t)j * n + i] = 0;
    }
  }
}

template <int GID>
int bases_precompute_impl(g753_ctx* ctx, g753_bases* b, unsigned copies) {
  constexpr int K = MsmCfg<GID>::K, T = 64;
  typedef typename MsmCfg<GID>::template SC<T> SC;
  typedef EcS<SC> E;
  const size_t n = b->n;
  const MsmPlan pl = msm_plan(msm_cost<GID>(), n, copies, ctx->forced_c);
  if (pl.copies <= 1 || n == 0) return G753_OK;
  if ((uint64_t)pl.copies * n > 0x7fffffffull) return fail(G753_ERR_BAD_ARG, "copies * n exceeds the 31-bit base index");
  const size_t pt_bytes = sizeof(Fq) * 2 * K;
  void* d_new = nullptr;
  uint8_t* d_inf = nullptr;
  G753_TRY(dev_alloc(&d_new, pt_bytes * n * pl.copies));
  int rc = dev_alloc((void**)&d_inf, n * pl.copies);
  if (rc == G753_OK) rc = d2d(d_new, b->d_points, pt_bytes * n, ctx->stream);
  if (rc == G753_OK) rc = b->d_inf ? d2d(d_inf, b->d_inf, n, ctx->stream) : dev_memset(d_inf, 0, n, ctx->stream);
  if (rc == G753_OK) {
    G753_LAUNCH_SMEM(k_bases_precompute<SC>, div_up(n, T), T * MsmCfg<GID>::TP, (slot_bytes<E, T>(E::PT + E::ADD_SCRATCH)), ctx->stream,
                     (Fq*)d_new, d_inf, (unsigned)n, pl.copies, pl.rows * pl.c);
    ctx->launches++;
    rc = launch_check("k_bases_precompute");
  }
  if (rc == G753_OK) rc = stream_sync(ctx->stream);
  if (rc != G753_OK) {
    dev_free(d_new);
    dev_free(d_inf);
    return rc;
  }
  dev_free(b->d_points);
  dev_free(b->d_inf);
  b->d_points = d_new;
  b->d_inf = d_inf;
  b->copies = pl.copies;
  b->c = pl.c;
  b->rows = pl.rows;
  return G753_OK;
}

// GroupProjective::batch_normalization / into_affine (short_weierstrass_projective.rs:402-442,
// 663-678): homogeneous (X:Y:Z) -> affine (X/Z, Y/Z); Z == 0 -> GroupAffine::zero() = (0, 1, true).  One thread
// per point with its own inversion: the prover normalises 3 points per proof.
template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS)
k_batch_normalize(const Fq* __restrict__ xyz, unsigned n, Fq* __restrict__ xy, uint8_t* __restrict__ inf) {
  typedef EcS<SC> E;
  typedef typename E::M M;
  constexpr int K = E::K, X = 0, Y = K, Z = 2 * K, ZI = 3 * K, S = 4 * K;
  unsigned i = E::M::T::item();
  if (i >= n) return;
  M::ldg(Z, xyz + (size_t)i * 3 * K + 2 * K);
  if (M::is_zero(Z)) {
    M::set_zero(X);
    M::stg(xy + (size_t)i * 2 * K, X);
    M::set_one(Y);  // GroupAffine::zero() = (0, 1, infinity) (short_weierstrass_projective.rs:130-132)
    M::stg(xy + (size_t)i * 2 * K + K, Y);
    inf[i] = 1;
    return;
  }
  M::ldg(X, xyz + (size_t)i * 3 * K);
  M::ldg(Y, xyz + (size_t)i * 3 * K + K);
  M::inv(ZI, Z, S);
  M::mul(X, X, ZI, S);
  M::mul(Y, Y, ZI, S);
  M::stg(xy + (size_t)i * 2 * K, X);
  M::stg(xy + (size_t)i * 2 * K + K, Y);
  inf[i] = 0;
}

template <int GID>
int batch_normalize_impl(g753_ctx* ctx, const uint64_t* xyz, size_t count, uint64_t* xy, uint8_t* infinity) {
  constexpr int K = MsmCfg<GID>::K, T = 32;
  typedef typename MsmCfg<GID>::template SC<T> SC;
  typedef EcS<SC> E;
  const size_t prj = sizeof(Fq) * 3 * K, aff = sizeof(Fq) * 2 * K;
  // staged through the context's grow-only I/O scratch: no cudaMalloc / cudaFree (a device-wide
  // synchronisation) on the prover's critical path
  G753_TRY(ctx->scratch_io.reserve((prj + aff + 1) * count + 1024));
  Carver cv(ctx->scratch_io.ptr);
  Fq* d_in = cv.take<Fq>(3 * K * count);
  Fq* d_out = cv.take<Fq>(2 * K * count);
  uint8_t* d_inf = cv.take<uint8_t>(count);
  G753_TRY(h2d(d_in, xyz, prj * count, ctx->stream));
  G753_LAUNCH_SMEM(k_batch_normalize<SC>, div_up(count, T), T * MsmCfg<GID>::TP, (slot_bytes<E, T>(4 * K + E::M::NTMP + 1)), ctx->stream,
                   d_in, (unsigned)count, d_out, d_inf);
  ctx->launches++;
  G753_TRY(launch_check("k_batch_normalize"));
  G753_TRY(d2h(xy, d_out, aff * count, ctx->stream));
  G753_TRY(d2h(infinity, d_inf, count, ctx->stream));
  return stream_sync(ctx->stream);
}

// FixedBaseMSM (algebra/src/msm/fixed_base.rs:6-80) + batch_normalization + into_affine
// (proof-systems/src/groth16/generator.rs:225-319 computes every query of a key this way): out[i] =
// scalars[i] * G in affine form.  The window table (8-bit windows: entry (w, j) = j * 2^(8w) * G, affine)
// is built per call on the device; the reference picks its window from the number of scalars
// (fixed_base.rs:7-13) - the resulting points do not depend on it.
constexpr unsigned FB_WINDOW = 8, FB_WINDOWS = (SCALAR_BITS + FB_WINDOW - 1) / FB_WINDOW, FB_ENTRIES = (1u << FB_WINDOW) - 1;

// window bases: wb[w] = 2^(8 w) * G, affine (one serial chain of 8 doublings per window)
template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS)
k_fb_bases(const Fq* __restrict__ g, Fq* __restrict__ wb) {
  typedef EcS<SC> E;
  typedef typename E::M M;
  constexpr int K = E::K, P = 0, S = E::PT;
  if (E::M::T::item() != 0) return;
  E::set_inf(P);
  E::madd_g(P, g, false, S);
  for (unsigned w = 0; w < FB_WINDOWS; w++) {
    Fq* o = wb + (size_t)w * 2 * K;
    if (E::is_inf(P)) {
      M::set_zero(S);
      M::stg(o, S);
      M::stg(o + K, S);
    } else {
      E::to_affine(P, S);
      M::stg(o, P);
      M::stg(o + K, P + K);
    }
    for (unsigned k = 0; k < FB_WINDOW; k++) E::dbl(P, S);
  }
}
// table[w * FB_ENTRIES + j - 1] = j * wb[w], affine: one thread per window, running sum
template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS)
k_fb_rows(const Fq* __restrict__ wb, Fq* __restrict__ table) {
  typedef EcS<SC> E;
  typedef typename E::M M;
  constexpr int K = E::K, P = 0, Q = E::PT, S = 2 * E::PT;
  unsigned w = E::M::T::item();
  if (w >= FB_WINDOWS) return;
  const Fq* base = wb + (size_t)w * 2 * K;
  M::ldg(S, base);
  M::ldg(S + K, base + K);
  const bool base_inf = M::is_zero(S) && M::is_zero(S + K);
  E::set_inf(P);
  for (unsigned j = 1; j <= FB_ENTRIES; j++) {
    Fq* o = table + ((size_t)w * FB_ENTRIES + j - 1) * 2 * K;
    if (!base_inf) E::madd_g(P, base, false, S);
    if (E::is_inf(P)) {
      M::set_zero(S);
      M::stg(o, S);
      M::stg(o + K, S);
      continue;
    }
    E::copy(Q, P);
    E::to_affine(Q, S);
    M::stg(o, Q);
    M::stg(o + K, Q + K);
  }
}

template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS)
k_fixed_base(const Fq* __restrict__ table, const uint32_t* __restrict__ scalars, unsigned n, Fq* __restrict__ out_xy,
             uint8_t* __restrict__ out_inf) {
  typedef EcS<SC> E;
  typedef typename E::M M;
  constexpr int K = E::K, P = 0, S = E::PT;
  unsigned i = E::M::T::item();
  if (i >= n) return;
  const uint32_t* s = scalars + (size_t)i * NL;
  E::set_inf(P);
  for (unsigned w = 0; w < FB_WINDOWS; w++) {
    const unsigned bit = w * FB_WINDOW;
    const uint32_t d = (s[bit >> 5] >> (bit & 31)) & 0xffu;      // 8-bit windows never straddle a limb
    if (!d) continue;
    const Fq* q = table + (size_t)(w * FB_ENTRIES + d - 1) * 2 * K;
    M::ldg(S, q);
    M::ldg(S + K, q + K);
    if (M::is_zero(S) && M::is_zero(S + K)) continue;            // table entry at infinity
    E::madd_g(P, q, false, S);
  }
  Fq* o = out_xy + (size_t)i * 2 * K;
  if (E::is_inf(P)) {            // GroupAffine::zero() = (0, 1, true)
    M::set_zero(S);
    M::stg(o, S);
    M::set_one(S);
    M::stg(o + K, S);
    out_inf[i] = 1;
    return;
  }
  E::to_affine(P, S);
  M::stg(o, P);
  M::stg(o + K, P + K);
  out_inf[i] = 0;
}

template <int GID>
int fixed_base_impl(g753_ctx* ctx, const uint64_t* base_xy, const uint64_t* scalars, size_t n, uint64_t* out_xy,
                    uint8_t* out_inf) {
  constexpr int K = MsmCfg<GID>::K, T = 32;
  typedef typename MsmCfg<GID>::template SC<T> SC;
  typedef EcS<SC> E;
  const size_t aff = sizeof(Fq) * 2 * K;
  const size_t entries = (size_t)FB_WINDOWS * FB_ENTRIES;
  G753_TRY(ctx->scratch.reserve(aff * (entries + 1 + FB_WINDOWS + n) + 96 * n + n + 8192));
  Carver cv(ctx->scratch.ptr);
  Fq* d_g = cv.take<Fq>(2 * K);
  Fq* d_table = cv.take<Fq>(2 * K * entries);
  uint32_t* d_sc = cv.take<uint32_t>(NL * n);
  Fq* d_out = cv.take<Fq>(2 * K * n);
  uint8_t* d_inf = cv.take<uint8_t>(n);
  G753_TRY(h2d(d_g, base_xy, aff, ctx->stream));
  G753_TRY(h2d(d_sc, scalars, 96 * n, ctx->stream));
  const size_t smem = slot_bytes<E, T>(2 * E::PT + E::ADD_SCRATCH);
  Fq* d_wb = cv.take<Fq>(2 * K * FB_WINDOWS);
  G753_LAUNCH_SMEM(k_fb_bases<SC>, 1, T * MsmCfg<GID>::TP, smem, ctx->stream, d_g, d_wb);
  G753_LAUNCH_SMEM(k_fb_rows<SC>, div_up(FB_WINDOWS, T), T * MsmCfg<GID>::TP, smem, ctx->stream, d_wb, d_table);
  ctx->launches++;
  G753_LAUNCH_SMEM(k_fixed_base<SC>, div_up(n, T), T * MsmCfg<GID>::TP, smem, ctx->stream, d_table, d_sc, (unsigned)n, d_out,
                   d_inf);
  ctx->launches += 2;
  G753_TRY(launch_check("k_fixed_base"));
  G753_TRY(d2h(out_xy, d_out, aff * n, ctx->stream));
  G753_TRY(d2h(out_inf, d_inf, n, ctx->stream));
  return stream_sync(ctx->stream);
}

#define G753_INSTANTIATE_GROUP(GID)                                                                        \
  template int msm_dispatch<GID>(g753_ctx*, const g753_bases*, size_t, size_t, const uint32_t*, void*);   \
  template int point_op_impl<GID>(g753_ctx*, int, const uint64_t*, const uint64_t*, uint64_t*);            \
  template int ext_op_impl<GID>(g753_ctx*, int, int, const uint64_t*, const uint64_t*, uint64_t*, size_t);  \
  template void points_sum_launch<GID>(g753_ctx*, const void*, size_t, void*);                               \
  template int bases_generate_impl<GID>(g753_ctx*, const uint64_t*, uint64_t, size_t, void*);                \
  template int bases_precompute_impl<GID>(g753_ctx*, g753_bases*, unsigned);                                 \
  template int batch_normalize_impl<GID>(g753_ctx*, const uint64_t*, size_t, uint64_t*, uint8_t*);        \
  template int fixed_base_impl<GID>(g753_ctx*, const uint64_t*, const uint64_t*, size_t, uint64_t*, uint8_t*);
