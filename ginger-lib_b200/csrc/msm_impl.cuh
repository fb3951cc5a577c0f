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
  hooks.user = ctx;
#endif
  const Fq* pts = (const Fq*)b->d_points + first * 2 * MsmCfg<GID>::K;
  const uint8_t* inf = b->d_inf ? b->d_inf + first : nullptr;
  return msm_run<GID>(ctx->scratch, ctx->stream, pts, inf, d_scalars, count, (Fq*)d_out, ctx->forced_c, hooks);
}

// group-law test hook on the product's own slot code (ec_slots.cuh); affine (0, 0) = infinity
template <class SC>
__global__ void __launch_bounds__(SC::M::T)
k_point_op(int op, const Fq* a, const Fq* b, const uint32_t* scalar, Fq* out) {
  typedef EcS<SC> E;
  typedef typename E::M M;
  constexpr int K = E::K, P = 0, Q = E::PT, S = 2 * E::PT;
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
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
  G753_LAUNCH_SMEM(k_point_op<SC>, 1, T, (slot_bytes<E, T>(2 * E::PT + E::ADD_SCRATCH)), ctx->stream, op, da, db,
                   ds, dout);
  ctx->launches++;
  G753_TRY(launch_check("k_point_op"));
  G753_TRY(d2h(out, dout, prj, ctx->stream));
  return stream_sync(ctx->stream);
}

template <int GID>
void points_sum_launch(g753_ctx* ctx, const void* d_pts, size_t count, void* d_out) {
  constexpr int T = 32;
  typedef typename MsmCfg<GID>::template SC<T> SC;
  typedef EcS<SC> E;
  G753_LAUNCH_SMEM(k_points_sum<SC>, 1, T, (slot_bytes<E, T>(2 * E::PT + E::ADD_SCRATCH)), ctx->stream,
                   (const Fq*)d_pts, (unsigned)count, (Fq*)d_out);
}


#define G753_INSTANTIATE_GROUP(GID)                                                                        \
  template int msm_dispatch<GID>(g753_ctx*, const g753_bases*, size_t, size_t, const uint32_t*, void*);   \
  template int point_op_impl<GID>(g753_ctx*, int, const uint64_t*, const uint64_t*, uint64_t*);            \
  template void points_sum_launch<GID>(g753_ctx*, const void*, size_t, void*);
