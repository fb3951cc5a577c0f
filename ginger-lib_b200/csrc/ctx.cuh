// Context / handle definitions shared by the translation units of libg753.so.
#pragma once
#include <map>
#include <mutex>
#include <new>
#include <vector>

#include "device.cuh"
#include "ntt.cuh"
#include "ntt_mixed.cuh"

using namespace g753;

enum { MSM_PHASES = 5 };

struct g753_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  Scratch scratch;      // MSM workspace
  Scratch scratch_io;   // host-API staging (scalars / NTT ping-pong)
  std::map<unsigned, NttTables> tables[2];  // per field, keyed by log_n
  std::map<uint64_t, MixedTables> mixed[2]; // per field, keyed by the mixed-radix size N
  uint64_t launches = 0;
  float phase_ms[MSM_PHASES] = {0, 0, 0, 0, 0};
  bool phases_valid = false;  // every phase event of the last MSM was recorded (count > 0, this context)
  unsigned last_plan[5] = {0, 0, 0, 0, 0};  // window bits, windows, bucket rows, key copies, accumulation form of the last MSM
  int forced_c = 0;
  int forced_affine = -1;  // G753_MSM_AFFINE: 0 / 1 force the accumulation form, unset = the group's default
  int tree_batch = 0;      // G753_TREE_BATCH: output slots per thread of the addition tree (0 = default)
  int tree_waves = 0;      // G753_TREE_WAVES: waves of blocks per level before batches grow (0 = default)
  std::mutex mu;
  unsigned scalar_chunks = 1;  // > 1 while g753_msm feeds the scalars of the running MSM in pieces
#if !defined(G753_HOST_EMUL)
  cudaEvent_t ev[MSM_PHASES + 1];
  bool ev_ok = false;
  enum { MAX_CHUNKS = 8 };
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t chunk_ev[MAX_CHUNKS];
  cudaEvent_t ready_ev;
  bool copy_ok = false;
#endif
};

struct g753_bases {
  int group = 0;
  size_t n = 0;
  void* d_points = nullptr;   // copies x n affine points; copy j = 2^(j * rows * c) * P_i
  uint8_t* d_inf = nullptr;   // infinity flags, copies x n (null: no base is infinite)
  unsigned copies = 1;        // > 1 after g753_bases_precompute
  unsigned c = 0, rows = 0;   // window bits / bucket rows the copies were built for
};

static int group_k(int group) {
  switch (group) {
    case G753_MNT4_G1: return 1;
    case G753_MNT4_G2: return 2;
    case G753_MNT6_G1: return 1;
    case G753_MNT6_G2: return 3;
    default: return 0;
  }
}

#if !defined(G753_HOST_EMUL)
static inline int use_device(g753_ctx* ctx) {
  cudaError_t e = cudaSetDevice(ctx->device);
  return e == cudaSuccess ? G753_OK : cuda_fail(e, "cudaSetDevice");
}
static inline void phase_mark(void* user, int phase) {
  g753_ctx* ctx = (g753_ctx*)user;
  if (!ctx->ev_ok || phase > MSM_PHASES) return;
  if (phase == 0) ctx->phases_valid = false;
  const bool ok = cudaEventRecord(ctx->ev[phase], ctx->stream) == cudaSuccess;
  if (!ok) (void)cudaGetLastError();  // timing is best effort: never leaves a stale error behind
  if (phase == MSM_PHASES) ctx->phases_valid = ok;
}
static inline void chunk_wait(void* user, int chunk) {
  g753_ctx* ctx = (g753_ctx*)user;
  if (ctx->copy_ok && ctx->scalar_chunks > 1 && chunk < g753_ctx::MAX_CHUNKS) {
    cudaError_t e = cudaStreamWaitEvent(ctx->stream, ctx->chunk_ev[chunk], 0);
    if (e != cudaSuccess) {
      // ordering is not optional: fall back to a full wait for the copy stream
      (void)cudaGetLastError();
      cudaStreamSynchronize(ctx->copy_stream);
    }
  }
}
#else
static inline int use_device(g753_ctx*) { return G753_OK; }
#endif

#define CHECK_CTX(ctx)                                              \
  do {                                                              \
    if (!(ctx)) return fail(G753_ERR_BAD_ARG, "null context");      \
    G753_TRY(use_device(ctx));                                      \
  } while (0)


// per-group entry points (one translation unit per group: msm_g0.cu .. msm_g3.cu)
template <int GID>
int msm_dispatch(g753_ctx* ctx, const g753_bases* b, size_t first, size_t count, const uint32_t* d_scalars,
                 void* d_out);
template <int GID>
int point_op_impl(g753_ctx* ctx, int op, const uint64_t* a, const uint64_t* b, uint64_t* out);
template <int GID>
int ext_op_impl(g753_ctx* ctx, int lanes, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);
template <int GID>
void points_sum_launch(g753_ctx* ctx, const void* d_pts, size_t count, void* d_out);
template <int GID>
int bases_generate_impl(g753_ctx* ctx, const uint64_t* gen_xy, uint64_t seed, size_t n, void* d_points);
template <int GID>
int bases_precompute_impl(g753_ctx* ctx, g753_bases* b, unsigned copies);
template <int GID>
int batch_normalize_impl(g753_ctx* ctx, const uint64_t* xyz, size_t count, uint64_t* xy, uint8_t* infinity);
template <int GID>
int fixed_base_impl(g753_ctx* ctx, const uint64_t* base_xy, const uint64_t* scalars, size_t n, uint64_t* out_xy,
                    uint8_t* out_inf);
