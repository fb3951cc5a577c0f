// Multi-scalar multiplication pipeline: sum_i s_i * P_i for one of the four prover groups.
//
// Replaces VariableBaseMSM::msm_inner (algebra/src/msm/variable_base.rs:10-83) with a
// different algorithm that yields the same group element:
//   reference: unsigned c-bit windows, 2^c - 1 buckets, one CPU task per window, serial
//              running-sum reduction, c from scalars.len() (variable_base.rs:14-18);
//   here:      signed-digit windows (2^(c-1) buckets), a counting sort of (digit, point)
//              pairs per window, one accumulation pass over the sorted runs, a multi-level
//              parallel running-sum reduction, Horner window fold.
// Reference semantics preserved: zero scalars and infinity bases contribute
// nothing, duplicate bases hit the doubling branch, P + (-P) gives infinity, count == 0
// returns infinity.  (The reference's scalar == 1 fast path is an optimisation, not a
// semantic: 1 * P is accumulated through window 0 like any other digit.)
//
// Kernels (all barrier-free, so the host-emulation test build can run them):
//   k_msm_digits     K4  scalar -> signed window digits + per-window histogram
//   k_scan_*         K4  exclusive scan of the histograms (bucket offsets)
//   k_msm_scatter    K4  counting-sort scatter of point indices into bucket order
//   k_bucket_acc     K5  one accumulator per bucket over its sorted run (mixed additions)
//   k_reduce_level   K6  sum_b b * B_b by segmented running sums, log_s(B) levels
//   k_window_combine K6  Horner fold of the window sums, XYZZ -> homogeneous projective
//   k_points_sum     K6  fold of per-shard partial results (multi-GPU)
#pragma once
#include "device.cuh"
#include "ec.cuh"

namespace g753 {

constexpr unsigned SCALAR_BITS = 753;      // FpParameters::MODULUS_BITS of both scalar fields
constexpr unsigned SCAN_CHUNK = 256;
constexpr unsigned REDUCE_SEG_LOG = 5;     // running-sum segment = 32 buckets
constexpr unsigned REDUCE_SEG = 1u << REDUCE_SEG_LOG;
constexpr unsigned MSM_MAX_C = 20;

struct MsmPlan {
  unsigned c;  // window bits
  unsigned W;  // windows: W * c >= SCALAR_BITS + 1 so the signed recoding never carries out
  unsigned B;  // buckets per window = 2^(c-1); bucket b holds digit magnitude b, 0 = discard
};

// field multiplications: mixed add 10, full add 14 (XYZZ); reduction does 2 full adds / bucket
static inline MsmPlan msm_plan(size_t n, int forced_c = 0) {
  MsmPlan best{0, 0, 0};
  double best_cost = 0;
  for (unsigned c = 3; c <= MSM_MAX_C; c++) {
    if (forced_c && (int)c != forced_c) continue;
    unsigned W = (SCALAR_BITS + 1 + c - 1) / c;
    double B = (double)(1u << (c - 1));
    double cost = (double)W * (10.0 * (double)n + 28.0 * B);
    if (best.c == 0 || cost < best_cost) {
      best = MsmPlan{c, W, 1u << (c - 1)};
      best_cost = cost;
    }
  }
  return best;
}

// ------------------------------------------------------------------------------------
// K4: digits + histogram
// ------------------------------------------------------------------------------------
__global__ void k_msm_digits(const uint32_t* __restrict__ scalars, const uint8_t* __restrict__ inf,
                             unsigned n, unsigned c, unsigned W, unsigned B,
                             uint32_t* __restrict__ digits, uint32_t* __restrict__ hist) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* s = scalars + (size_t)i * NL;
  bool skip = inf != nullptr && inf[i] != 0;
  const uint32_t half = 1u << (c - 1);
  const uint32_t mask = (1u << c) - 1;
  uint32_t carry = 0;
  for (unsigned w = 0; w < W; w++) {
    unsigned bit = w * c;
    unsigned limb = bit >> 5, sh = bit & 31;
    uint64_t two = 0;
    if (limb < (unsigned)NL) two = s[limb];
    if (limb + 1 < (unsigned)NL) two |= (uint64_t)s[limb + 1] << 32;
    uint32_t raw = ((uint32_t)(two >> sh) & mask) + carry;
    uint32_t mag, neg;
    if (raw > half) {
      mag = (1u << c) - raw;
      neg = 0x80000000u;
      carry = 1;
    } else {
      mag = raw;
      neg = 0;
      carry = 0;
    }
    if (skip) mag = 0;
    digits[(size_t)w * n + i] = mag ? (mag | neg) : 0u;
    if (mag) atomicAdd(&hist[(size_t)w * (B + 1) + mag], 1u);
  }
}

// exclusive scan of each window's histogram, in three barrier-free steps
__global__ void k_scan_chunks(const uint32_t* __restrict__ hist, unsigned len, unsigned n_chunks,
                              unsigned W, uint32_t* __restrict__ chunk_sums) {
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= W * n_chunks) return;
  unsigned w = t / n_chunks, ch = t % n_chunks;
  unsigned lo = ch * SCAN_CHUNK, hi = lo + SCAN_CHUNK < len ? lo + SCAN_CHUNK : len;
  uint32_t sum = 0;
  for (unsigned k = lo; k < hi; k++) sum += hist[(size_t)w * len + k];
  chunk_sums[t] = sum;
}
__global__ void k_scan_tops(uint32_t* __restrict__ chunk_sums, unsigned n_chunks, unsigned W) {
  unsigned w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  uint32_t run = 0;
  for (unsigned k = 0; k < n_chunks; k++) {
    uint32_t v = chunk_sums[(size_t)w * n_chunks + k];
    chunk_sums[(size_t)w * n_chunks + k] = run;
    run += v;
  }
}
__global__ void k_scan_apply(const uint32_t* __restrict__ hist, const uint32_t* __restrict__ chunk_sums,
                             unsigned len, unsigned n_chunks, unsigned W,
                             uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursor) {
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= W * n_chunks) return;
  unsigned w = t / n_chunks, ch = t % n_chunks;
  unsigned lo = ch * SCAN_CHUNK, hi = lo + SCAN_CHUNK < len ? lo + SCAN_CHUNK : len;
  uint32_t run = chunk_sums[t];
  for (unsigned k = lo; k < hi; k++) {
    size_t idx = (size_t)w * len + k;
    uint32_t v = hist[idx];
    offsets[idx] = run;
    cursor[idx] = run;
    run += v;
  }
}

__global__ void k_msm_scatter(const uint32_t* __restrict__ digits, unsigned n, unsigned W, unsigned B,
                              uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)W * n) return;
  unsigned w = (unsigned)(t / n), i = (unsigned)(t % n);
  uint32_t d = digits[t];
  uint32_t mag = d & 0x7fffffffu;
  if (!mag) return;
  uint32_t pos = atomicAdd(&cursor[(size_t)w * (B + 1) + mag], 1u);
  sorted[(size_t)w * n + pos] = i | (d & 0x80000000u);
}

// ------------------------------------------------------------------------------------
// K5: bucket accumulation
// ------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(128)
k_bucket_acc(const Affine<C>* __restrict__ bases, const uint32_t* __restrict__ sorted,
             const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ ends, unsigned n,
             unsigned W, unsigned B, Xyzz<C>* __restrict__ buckets) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)W * (B + 1)) return;
  unsigned w = (unsigned)(t / (B + 1));
  Xyzz<C> acc = xyzz_inf<C>();
  uint32_t lo = offsets[t], hi = ends[t];
  if (t % (B + 1) == 0) hi = lo;  // bucket 0 is the discard bucket
  for (uint32_t k = lo; k < hi; k++) {
    uint32_t e = sorted[(size_t)w * n + k];
    Affine<C> q = bases[e & 0x7fffffffu];
    if (e & 0x80000000u) q.y = C::F::neg(q.y);
    xyzz_madd<C>(acc, q);
  }
  buckets[t] = acc;
}

// ------------------------------------------------------------------------------------
// K6: bucket reduction  S_w = sum_b b * B_{w,b}
// One level turns  G(X, Y, f) = sum_i Y_i + f * sum_i i * X_i  over n_in entries into the same
// problem over n_out = ceil(n_in / s) entries with f' = f * s:
//   R_j  = sum_{i in seg j} X_i
//   Y'_j = sum_{i in seg j} Y_i + f * sum_{i in seg j} (i - j s) X_i        (running sums)
// ------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(128)
k_reduce_level(const Xyzz<C>* __restrict__ X, const Xyzz<C>* __restrict__ Y, unsigned n_in,
               unsigned log2f, unsigned W, unsigned n_out, Xyzz<C>* __restrict__ R,
               Xyzz<C>* __restrict__ Yout) {
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= W * n_out) return;
  unsigned w = t / n_out, j = t % n_out;
  unsigned lo = j * REDUCE_SEG;
  unsigned hi = lo + REDUCE_SEG < n_in ? lo + REDUCE_SEG : n_in;
  const Xyzz<C>* x = X + (size_t)w * n_in;
  Xyzz<C> running = xyzz_inf<C>();
  Xyzz<C> acc = xyzz_inf<C>();
  for (unsigned i = hi - 1; i > lo; i--) {
    xyzz_add<C>(running, x[i]);
    xyzz_add<C>(acc, running);
  }
  xyzz_add<C>(running, x[lo]);
  for (unsigned k = 0; k < log2f; k++) xyzz_dbl<C>(acc);
  if (Y != nullptr) {
    const Xyzz<C>* y = Y + (size_t)w * n_in;
    for (unsigned i = lo; i < hi; i++) xyzz_add<C>(acc, y[i]);
  }
  R[t] = running;
  Yout[t] = acc;
}

// Horner fold over the W window sums (stride between windows given), then convert to the
// reference's homogeneous projective layout.  One thread: W*c doublings are a serial chain.
template <class C>
__global__ void k_window_combine(const Xyzz<C>* __restrict__ sums, unsigned stride, unsigned W,
                                 unsigned c, typename C::F* __restrict__ out_xyz) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Xyzz<C> total = xyzz_inf<C>();
  for (int w = (int)W - 1; w >= 0; w--) {
    if (w != (int)W - 1)
      for (unsigned k = 0; k < c; k++) xyzz_dbl<C>(total);
    xyzz_add<C>(total, sums[(size_t)w * stride]);
  }
  typename C::F X, Y, Z;
  xyzz_to_projective<C>(total, X, Y, Z);
  out_xyz[0] = X;
  out_xyz[1] = Y;
  out_xyz[2] = Z;
}

template <class C>
__global__ void k_write_infinity(typename C::F* __restrict__ out_xyz) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  typedef typename C::F F;
  out_xyz[0] = F::zero();
  out_xyz[1] = F::one();
  out_xyz[2] = F::zero();
}

// homogeneous projective (X:Y:Z) -> XYZZ
template <class C>
G753_HD Xyzz<C> xyzz_from_projective(const typename C::F& X, const typename C::F& Y, const typename C::F& Z) {
  typedef typename C::F F;
  if (F::is_zero(Z)) return xyzz_inf<C>();
  Xyzz<C> r;
  r.zz = F::sqr(Z);
  r.zzz = F::mul(r.zz, Z);
  r.x = F::mul(X, Z);
  r.y = F::mul(Y, r.zz);
  return r;
}

// sum of `count` projective points (the multi-GPU fold; count is the number of ranks)
template <class C>
__global__ void k_points_sum(const typename C::F* __restrict__ pts, unsigned count,
                             typename C::F* __restrict__ out_xyz) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Xyzz<C> total = xyzz_inf<C>();
  for (unsigned k = 0; k < count; k++) {
    Xyzz<C> p = xyzz_from_projective<C>(pts[3 * k], pts[3 * k + 1], pts[3 * k + 2]);
    xyzz_add<C>(total, p);
  }
  typename C::F X, Y, Z;
  xyzz_to_projective<C>(total, X, Y, Z);
  out_xyz[0] = X;
  out_xyz[1] = Y;
  out_xyz[2] = Z;
}

// zero the coordinates of bases flagged infinite so that (0, 0) is the only encoding the
// accumulation kernels ever see
template <class C>
__global__ void k_bases_sanitize(Affine<C>* __restrict__ bases, const uint8_t* __restrict__ inf, unsigned n) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (inf[i]) {
    bases[i].x = C::F::zero();
    bases[i].y = C::F::zero();
  }
}

// ------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------
struct MsmPhaseTimer;  // defined by the product build (CUDA events); a no-op in emulation

struct MsmWorkspaceSizes {
  size_t digits, hist, offsets, cursor, chunk_sums, sorted, buckets, level, total;
  unsigned n_chunks, level_entries;
};

template <class C>
static inline MsmWorkspaceSizes msm_workspace(const MsmPlan& pl, size_t n) {
  MsmWorkspaceSizes s;
  size_t len = (size_t)pl.B + 1;
  s.n_chunks = div_up(len, SCAN_CHUNK);
  s.digits = Carver::pad(sizeof(uint32_t) * pl.W * n);
  s.sorted = Carver::pad(sizeof(uint32_t) * pl.W * n);
  s.hist = Carver::pad(sizeof(uint32_t) * pl.W * len);
  s.offsets = s.hist;
  s.cursor = s.hist;
  s.chunk_sums = Carver::pad(sizeof(uint32_t) * pl.W * s.n_chunks);
  s.buckets = Carver::pad(sizeof(Xyzz<C>) * pl.W * len);
  // reduction levels: n_out entries per window per level, two arrays (R, Y') per level
  unsigned entries = 0;
  for (size_t m = len; m > 1;) {
    m = div_up(m, REDUCE_SEG);
    entries += (unsigned)m;
  }
  if (entries == 0) entries = 1;
  s.level_entries = entries;
  s.level = Carver::pad(sizeof(Xyzz<C>) * pl.W * entries) * 2;
  s.total = s.digits + s.sorted + s.hist * 3 + s.chunk_sums + s.buckets + s.level + 4096;
  return s;
}

struct MsmHooks {  // phase timing hooks; the emulation build leaves them null
  void (*mark)(void* user, int phase) = nullptr;
  void* user = nullptr;
  uint64_t* launches = nullptr;
};

#define G753_MSM_LAUNCH(hooks, ...)            \
  do {                                         \
    G753_LAUNCH(__VA_ARGS__);                  \
    if ((hooks).launches) ++*(hooks).launches; \
  } while (0)

// d_bases: `count` affine points (device); d_inf: their infinity flags (device, may be null);
// d_scalars: count x 24 u32 canonical (device); d_out: 3 field elements (device)
template <class C>
static int msm_run(Scratch& scratch, cudaStream_t stream, const Affine<C>* d_bases, const uint8_t* d_inf,
                   const uint32_t* d_scalars, size_t count, typename C::F* d_out, int forced_c,
                   MsmHooks hooks) {
  typedef Xyzz<C> P;
  if (count == 0) {
    G753_MSM_LAUNCH(hooks, k_write_infinity<C>, 1, 1, stream, d_out);
    return launch_check("k_write_infinity");
  }
  if (count > 0x7fffffffull) return G753_ERR_BAD_ARG;
  const unsigned n = (unsigned)count;
  const MsmPlan pl = msm_plan(n, forced_c);
  const MsmWorkspaceSizes ws = msm_workspace<C>(pl, n);
  G753_TRY(scratch.reserve(ws.total));
  Carver cv(scratch.ptr);
  const size_t len = (size_t)pl.B + 1;
  uint32_t* digits = cv.take<uint32_t>((size_t)pl.W * n);
  uint32_t* sorted = cv.take<uint32_t>((size_t)pl.W * n);
  uint32_t* hist = cv.take<uint32_t>(pl.W * len);
  uint32_t* offsets = cv.take<uint32_t>(pl.W * len);
  uint32_t* cursor = cv.take<uint32_t>(pl.W * len);
  uint32_t* chunk_sums = cv.take<uint32_t>((size_t)pl.W * ws.n_chunks);
  P* buckets = cv.take<P>(pl.W * len);
  P* lvl_r = cv.take<P>((size_t)pl.W * ws.level_entries);
  P* lvl_y = cv.take<P>((size_t)pl.W * ws.level_entries);

  if (hooks.mark) hooks.mark(hooks.user, 0);
  G753_TRY(dev_memset(hist, 0, sizeof(uint32_t) * pl.W * len, stream));
  G753_MSM_LAUNCH(hooks, k_msm_digits, div_up(n, 256), 256, stream, d_scalars, d_inf, n, pl.c, pl.W, pl.B,
                  digits, hist);
  if (hooks.mark) hooks.mark(hooks.user, 1);
  G753_MSM_LAUNCH(hooks, k_scan_chunks, div_up((size_t)pl.W * ws.n_chunks, 128), 128, stream, hist,
                  (unsigned)len, ws.n_chunks, pl.W, chunk_sums);
  G753_MSM_LAUNCH(hooks, k_scan_tops, div_up(pl.W, 64), 64, stream, chunk_sums, ws.n_chunks, pl.W);
  G753_MSM_LAUNCH(hooks, k_scan_apply, div_up((size_t)pl.W * ws.n_chunks, 128), 128, stream, hist,
                  chunk_sums, (unsigned)len, ws.n_chunks, pl.W, offsets, cursor);
  G753_MSM_LAUNCH(hooks, k_msm_scatter, div_up((size_t)pl.W * n, 256), 256, stream, digits, n, pl.W, pl.B,
                  cursor, sorted);
  if (hooks.mark) hooks.mark(hooks.user, 2);
  G753_MSM_LAUNCH(hooks, k_bucket_acc<C>, div_up(pl.W * len, 128), 128, stream, d_bases, sorted, offsets,
                  cursor, n, pl.W, pl.B, buckets);
  if (hooks.mark) hooks.mark(hooks.user, 3);
  // reduction levels
  const P* X = buckets;
  const P* Y = nullptr;
  unsigned n_in = (unsigned)len, log2f = 0;
  size_t lvl_off = 0;
  const P* window_sums = nullptr;
  for (;;) {
    unsigned n_out = div_up(n_in, REDUCE_SEG);
    P* R = lvl_r + lvl_off;
    P* Yo = lvl_y + lvl_off;
    G753_MSM_LAUNCH(hooks, k_reduce_level<C>, div_up((size_t)pl.W * n_out, 128), 128, stream, X, Y, n_in,
                    log2f, pl.W, n_out, R, Yo);
    lvl_off += (size_t)pl.W * n_out;
    X = R;
    Y = Yo;
    n_in = n_out;
    log2f += REDUCE_SEG_LOG;
    if (n_out == 1) {
      window_sums = Yo;
      break;
    }
  }
  if (hooks.mark) hooks.mark(hooks.user, 4);
  G753_MSM_LAUNCH(hooks, k_window_combine<C>, 1, 32, stream, window_sums, 1u, pl.W, pl.c, d_out);
  if (hooks.mark) hooks.mark(hooks.user, 5);
  return launch_check("msm_run");
}

}  // namespace g753
