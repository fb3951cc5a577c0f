// Multi-scalar multiplication pipeline: sum_i s_i * P_i for one of the four prover groups.
//
// Replaces VariableBaseMSM::msm_inner (algebra/src/msm/variable_base.rs:10-83) with a
// different algorithm that yields the same group element:
//   reference: unsigned c-bit windows, 2^c - 1 buckets, one CPU task per window, serial
//              running-sum reduction, c from scalars.len() (variable_base.rs:14-18);
//   here:      signed-digit windows (2^(c-1) buckets), a counting sort of (digit, point)
//              pairs per window, accumulation of the sorted runs - a pairwise tree of affine
//              additions with shared inversions, or length-balanced XYZZ running sums for short
//              inputs -, a multi-level parallel running-sum reduction, Horner window fold.
// Reference semantics preserved: zero scalars and infinity bases contribute
// nothing, duplicate bases hit the doubling branch, P + (-P) gives infinity, count == 0
// returns infinity.  (The reference's scalar == 1 fast path is an optimisation, not a
// semantic: 1 * P is accumulated through window 0 like any other digit.)
//
// Kernels (all barrier-free, so the host-emulation test build can run them):
//   k_msm_digits      K4  scalar -> signed window digits + per-window histogram
//   k_scan_*          K4  exclusive scan of the histograms (bucket offsets)
//   k_msm_scatter     K4  counting-sort scatter of point indices into bucket order
//   k_tree_max/_round/_finish  K5t  the runs summed level by level, 6 products per addition (the default
//                         from 2^22 window entries)
//   k_item_*          K5  (short inputs) cut every bucket's run into work items of <= ITEM_LEN points and order
//                         the items by length (longest first), so the 32 lanes of a warp run
//                         the same number of mixed additions
//   k_bucket_acc      K5  one thread per item: XYZZ accumulator in shared-memory slots
//   k_bucket_fixup    K5  buckets cut into several items: sum the partial results
//   k_reduce_level    K6  sum_b b * B_b by segmented running sums, log_s(B) levels
//   k_window_combine  K6  Horner fold of the window sums, XYZZ -> homogeneous projective
//   k_points_sum      K6  fold of per-shard partial results (multi-GPU)
#pragma once
#include <type_traits>

#include "coop.cuh"
#include "device.cuh"
#include "ec_slots.cuh"

namespace g753 {

constexpr unsigned SCALAR_BITS = 753;      // FpParameters::MODULUS_BITS of both scalar fields
constexpr unsigned SCAN_CHUNK = 256;
constexpr unsigned REDUCE_SEG_LOG = 3;     // running-sum segment = 8 buckets: short serial chains, more levels
constexpr unsigned REDUCE_SEG = 1u << REDUCE_SEG_LOG;
constexpr unsigned MSM_MAX_C = 20;
constexpr unsigned ITEM_LEN = 128;         // longest run one thread accumulates
constexpr unsigned ITEM_REP = 16;          // replicated length counters (spreads the atomics)

// per-group launch shapes: NC_* columns (curve operations) per block, TP lanes per column; chosen so
// that the slot footprint (accumulate: 8 slots on G1, 7K + NTMP on G2; reduce: 12K + NTMP slots of 96 B per column)
// leaves >= 8 warps per SM inside the 227 KB of shared memory
template <int GID> struct MsmCfg;
// TPA lanes per column in the accumulation kernels, TP in all others (reduction, fold, key set-up);
// the TEST-ONLY host build runs one lane everywhere
#if defined(G753_HOST_EMUL)
#define G753_TP2A 1
#define G753_TP2 1
#define G753_TP3A 1
#define G753_TP3 1
#else
#define G753_TP2A 2
#define G753_TP2 4
#define G753_TP3A 3
#define G753_TP3 8
#endif
#ifndef G753_NC_ACC1
#define G753_NC_ACC1 128  // columns per block of the prime-field accumulation kernel
#endif
template <> struct MsmCfg<0> {
  static constexpr int K = 1, TP = 1, TPA = 1, NC_ACC = G753_NC_ACC1, NC_RED = 96;
  template <int NC, int LANES = 1> using SC = SCurveM4G1<Lay<NC, LANES>>;
};
template <> struct MsmCfg<1> {
  static constexpr int K = 2, TP = G753_TP2, TPA = G753_TP2A, NC_ACC = 16, NC_RED = 32;
  template <int NC, int LANES = G753_TP2> using SC = SCurveM4G2<Lay<NC, LANES>>;
};
template <> struct MsmCfg<2> {
  static constexpr int K = 1, TP = 1, TPA = 1, NC_ACC = G753_NC_ACC1, NC_RED = 96;
  template <int NC, int LANES = 1> using SC = SCurveM6G1<Lay<NC, LANES>>;
};
template <> struct MsmCfg<3> {
  static constexpr int K = 3, TP = G753_TP3, TPA = G753_TP3A, NC_ACC = G753_TP3A == 3 ? 40 : 32, NC_RED = 16;
  template <int NC, int LANES = G753_TP3> using SC = SCurveM6G2<Lay<NC, LANES>>;
};

struct MsmPlan {
  unsigned c;       // window bits
  unsigned W;       // windows: W * c >= SCALAR_BITS + 1 so the signed recoding never carries out
  unsigned B;       // buckets per window = 2^(c-1); bucket b holds digit magnitude b, 0 = discard
  unsigned rows;    // bucket rows = ceil(W / copies): window w = j * rows + r lands in row r and
                    // uses copy j of the key (copy j holds 2^(j * rows * c) * P_i); copies = 1: rows = W
  unsigned copies;  // key copies actually used = ceil(W / rows)
};

// Cost model in units of one base-field multiplication of the saturated accumulation kernel (7 G/s
// measured).  Per point and window: one mixed addition (10 products of the coordinate field).  Per bucket:
// the running-sum reduction, 2 full additions on a kernel that runs at about half rate - 50 reproduces
// the measured reduce phases (2^22 points, 5 rows of 2^19 buckets: 17 ms; 2^19 points, one row: 4.9 ms).
// Per bucket row: the Horner fold on ONE cooperating warp (coop.cuh), c doublings + one addition of the
// window sum, a latency chain measured at ~3.5 us per G1 doubling = 25 000 units (Fq2: x 4, Fq3: x 8.7;
// the one-thread fold it replaces cost 200 000 per doubling).
struct MsmCost {
  double madd, bucket, fold_dbl, fold_add;
};
template <int GID>
static inline MsmCost msm_cost() {
  const double tower = MsmCfg<GID>::K == 1 ? 1.0 : MsmCfg<GID>::K == 2 ? 3.4 : 9.5;   // coordinate-field cost ratio
  const double fold = MsmCfg<GID>::K == 1 ? 1.0 : MsmCfg<GID>::K == 2 ? 4.0 : 8.7;
  return MsmCost{10.0 * tower, 50.0 * tower, 25000.0 * fold, 45000.0 * fold};
}
static inline MsmPlan msm_plan(const MsmCost& k, size_t n, unsigned copies = 1, int forced_c = 0, unsigned forced_rows = 0) {
  MsmPlan best{0, 0, 0, 0, 0};
  double best_cost = 0;
  if (copies == 0) copies = 1;
  for (unsigned c = 3; c <= MSM_MAX_C; c++) {
    if (forced_c && (int)c != forced_c) continue;
    unsigned W = (SCALAR_BITS + 1 + c - 1) / c;
    unsigned rows = forced_rows ? forced_rows : (W + copies - 1) / copies;
    double B = (double)(1u << (c - 1));
    double cost = (double)W * k.madd * (double)n + (double)rows * k.bucket * B +
                  (double)rows * ((double)c * k.fold_dbl + k.fold_add);
    if (best.c == 0 || cost < best_cost) {
      best = MsmPlan{c, W, 1u << (c - 1), rows, (W + rows - 1) / rows};
      best_cost = cost;
    }
  }
  return best;
}

// ------------------------------------------------------------------------------------
// K4: digits + histogram
// ------------------------------------------------------------------------------------
// window w = j * rows + r: histogram row r; inf (may be null) holds one flag per (copy, point):
// inf[j * inf_stride + i]
// handles scalars [first, last) of the n (chunks let the upload of the next scalars overlap)
static __global__ void k_msm_digits(const uint32_t* __restrict__ scalars, const uint8_t* __restrict__ inf,
                             size_t inf_stride, unsigned n, unsigned first, unsigned last, unsigned c, unsigned W,
                             unsigned B, unsigned rows, uint32_t* __restrict__ digits, uint32_t* __restrict__ hist) {
  unsigned i = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= last) return;
  const uint32_t* s = scalars + (size_t)i * NL;
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
    if (inf != nullptr && inf[(size_t)(w / rows) * inf_stride + i] != 0) mag = 0;
    digits[(size_t)w * n + i] = mag ? (mag | neg) : 0u;
    if (mag) atomicAdd(&hist[(size_t)(w % rows) * (B + 1) + mag], 1u);
  }
}

// exclusive scan of each row (`W` rows of `len` counters), in three barrier-free steps
static __global__ void k_scan_chunks(const uint32_t* __restrict__ hist, unsigned len, unsigned n_chunks,
                              unsigned W, uint32_t* __restrict__ chunk_sums) {
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= W * n_chunks) return;
  unsigned w = t / n_chunks, ch = t % n_chunks;
  unsigned lo = ch * SCAN_CHUNK, hi = lo + SCAN_CHUNK < len ? lo + SCAN_CHUNK : len;
  uint32_t sum = 0;
  for (unsigned k = lo; k < hi; k++) sum += hist[(size_t)w * len + k];
  chunk_sums[t] = sum;
}
// row totals go to totals[w] when given
static __global__ void k_scan_tops(uint32_t* __restrict__ chunk_sums, unsigned n_chunks, unsigned W,
                            uint32_t* __restrict__ totals) {
  unsigned w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  uint32_t run = 0;
  for (unsigned k = 0; k < n_chunks; k++) {
    uint32_t v = chunk_sums[(size_t)w * n_chunks + k];
    chunk_sums[(size_t)w * n_chunks + k] = run;
    run += v;
  }
  if (totals) totals[w] = run;
}
static __global__ void k_scan_apply(const uint32_t* __restrict__ hist, const uint32_t* __restrict__ chunk_sums,
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
    if (cursor) cursor[idx] = run;
    run += v;
  }
}

// row r of `sorted` holds row_cap = copies * n entries: base index (j * copy_stride + i) | sign
static __global__ void k_msm_scatter(const uint32_t* __restrict__ digits, unsigned n, unsigned W, unsigned B,
                              unsigned rows, size_t copy_stride, size_t row_cap,
                              uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)W * n) return;
  unsigned w = (unsigned)(t / n), i = (unsigned)(t % n);
  uint32_t d = digits[t];
  uint32_t mag = d & 0x7fffffffu;
  if (!mag) return;
  const unsigned r = w % rows, j = w / rows;
  uint32_t pos = atomicAdd(&cursor[(size_t)r * (B + 1) + mag], 1u);
  sorted[(size_t)r * row_cap + pos] = (uint32_t)(j * copy_stride + i) | (d & 0x80000000u);
}

// ------------------------------------------------------------------------------------
// K5a: work items.  Bucket t = r * (B+1) + b owns sorted[r*row_cap + offsets[t] .. r*row_cap + ends[t]).
// It is cut into ceil(count / ITEM_LEN) items; item ids are item_off[t] + j.  The result of an
// item goes to points[t] when the bucket has a single item, else to points[NB + item id]
// (summed into points[t] by k_bucket_fixup).  len_hist[key * ITEM_REP + r] counts items of
// length ITEM_LEN - key.
// ------------------------------------------------------------------------------------
static __global__ void k_item_count(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ ends,
                             unsigned NB, unsigned B, uint32_t* __restrict__ item_cnt,
                             uint32_t* __restrict__ len_hist) {
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= NB) return;
  uint32_t cnt = (t % (B + 1) == 0) ? 0u : ends[t] - offsets[t];  // bucket 0 is the discard bucket
  uint32_t items = (cnt + ITEM_LEN - 1) / ITEM_LEN;
  item_cnt[t] = items;
  if (!items) return;
  const unsigned r = t % ITEM_REP;
  if (items > 1) atomicAdd(&len_hist[0 * ITEM_REP + r], items - 1);  // full-length items
  uint32_t last = cnt - (items - 1) * ITEM_LEN;
  atomicAdd(&len_hist[(ITEM_LEN - last) * ITEM_REP + r], 1u);
}

// single-thread exclusive scan of the ITEM_LEN * ITEM_REP length counters
static __global__ void k_item_len_scan(const uint32_t* __restrict__ len_hist, uint32_t* __restrict__ len_cursor,
                                uint32_t* __restrict__ item_total) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  uint32_t run = 0;
  for (unsigned k = 0; k < ITEM_LEN * ITEM_REP; k++) {
    len_cursor[k] = run;
    run += len_hist[k];
  }
  *item_total = run;
}

struct MsmItem {  // 16 bytes
  uint32_t start;  // first entry, as a flat index into sorted[]
  uint32_t len;    // entries to accumulate (1..ITEM_LEN)
  uint32_t dest;   // index into points[]
  uint32_t pad;
};

static __global__ void k_item_emit(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ ends,
                            const uint32_t* __restrict__ item_cnt, const uint32_t* __restrict__ item_off,
                            unsigned NB, unsigned B, size_t row_cap, uint32_t* __restrict__ len_cursor,
                            MsmItem* __restrict__ items, uint32_t* __restrict__ part_bucket) {
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= NB) return;
  const uint32_t cnt_items = item_cnt[t];
  if (!cnt_items) return;
  const unsigned w = t / (B + 1);
  const unsigned r = t % ITEM_REP;
  const uint32_t lo = offsets[t], hi = ends[t];
  const uint32_t first_id = item_off[t];
  for (uint32_t j = 0; j < cnt_items; j++) {
    uint32_t s = lo + j * ITEM_LEN;
    uint32_t len = hi - s < ITEM_LEN ? hi - s : ITEM_LEN;
    uint32_t pos = atomicAdd(&len_cursor[(ITEM_LEN - len) * ITEM_REP + r], 1u);
    MsmItem it;
    it.start = (uint32_t)((size_t)w * row_cap + s);
    it.len = len;
    it.dest = cnt_items == 1 ? t : NB + first_id + j;
    it.pad = 0;
    items[pos] = it;
    part_bucket[first_id + j] = t;
  }
}

// ------------------------------------------------------------------------------------
// K5b: bucket accumulation.  One thread per item; the accumulator and the formula temporaries
// live in shared-memory slots (ec_slots.cuh), bases are gathered straight from HBM.
// ------------------------------------------------------------------------------------
template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS)
k_bucket_acc(const Fq* __restrict__ bases, const uint32_t* __restrict__ sorted,
             const MsmItem* __restrict__ items, const uint32_t* __restrict__ item_total,
             Fq* __restrict__ points) {
  typedef EcS<SC> E;
  if (E::M::T::idle()) return;
  unsigned pos = E::M::T::item();
  if (pos >= *item_total) return;
  const MsmItem it = items[pos];
  E::set_inf(0);
  uint32_t e = sorted[it.start];
  for (uint32_t k = 0; k < it.len; k++) {
    const Fq* q = bases + (size_t)(e & 0x7fffffffu) * (2 * E::K);
    const bool neg = (e >> 31) != 0;
    if (k + 1 < it.len) {
      // the gather of the NEXT base is the one long-latency load of the loop: pull its lines into
      // L2 while this addition runs (no registers, no shared memory)
      e = sorted[it.start + k + 1];
#if defined(__CUDA_ARCH__)
      const char* nq = (const char*)(bases + (size_t)(e & 0x7fffffffu) * (2 * E::K));
#pragma unroll
      for (int off = 0; off < (int)(2 * E::K * sizeof(Fq)); off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nq + off));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(nq + 2 * E::K * sizeof(Fq) - 1));
#endif
    }
    E::madd_acc_g(0, q, neg, E::PT);
  }
  E::stg(points + (size_t)it.dest * E::PT, 0);
}

// ------------------------------------------------------------------------------------
// K5t: bucket accumulation as a PAIRWISE TREE of affine additions with shared inversions.
// An affine addition is lambda = (y2 - y1) / (x2 - x1), x3 = lambda^2 - x1 - x2,
// y3 = lambda (x1 - x3) - y1: 1 inversion + 2 products + 1 squaring.  Inverting the PRODUCT of many
// denominators once (Montgomery's trick, 3 products per denominator; the reference's
// batch_normalization does the same for its Z inversions, short_weierstrass_projective.rs:402-442)
// makes an addition 6 products + (one safegcd inversion ~ 45 products) / batch, against the 10 of the
// XYZZ mixed addition.  The trick needs many INDEPENDENT additions, which a running bucket sum does not
// offer; a tree does: level 0 of bucket t is its run of sorted (point, sign) entries, level L + 1 holds
// the sums of the pairs (2l, 2l + 1) of level L (an odd last element is carried over), and after
// ceil(log2(count)) rounds one element is left.  The additions of a round are all independent, so a
// thread takes `batch` consecutive OUTPUT slots whatever buckets they belong to:
//   phase 1 (forward)   den_j = x2 - x1, the running product of the denominators BEFORE slot j is
//                       parked in the slot's own output cell;
//   one inversion of the product of all denominators of the thread;
//   phase 2 (backward)  1 / den_j = parked prefix * inverse, inverse *= den_j, finish the addition and
//                       overwrite the cell with (x3, y3).
// Exceptional pairs take no part in the shared inversion except the doubling (denominator 2 y1):
// P + (-P) stores the point at infinity as (0, 0) (not on these curves, b != 0), an operand at
// infinity passes the other one through.
// No index arrays, no scans: level L of bucket t starts at S_L[t], S_0[t] = row * row_cap + offsets[t],
// S_{L+1}[t] = floor((S_L[t] + t) / 2) - one slot of padding per bucket and level keeps the runs
// disjoint (S_L[t+1] - S_L[t] >= count_L[t] implies the same one level up) - and holds
// count_L[t] = ceil(count_0[t] / 2^L) elements; the bucket of an output slot is found by bisection on
// S_{L+1}[.].  Level L >= 1 lives in one of two ping-pong arrays of affine points.  The addition of a
// bucket's LAST pair writes the sum straight to buckets[t] (XYZZ: x, y, 1, 1) and later rounds skip the
// bucket; k_tree_finish copies the buckets that hold a single entry.  The number of rounds depends on the
// fullest bucket (known on the device only): the host launches the ceil(log2(row_cap)) rounds a single
// bucket could need and a round whose input level has no bucket with two elements left returns at once.
// ------------------------------------------------------------------------------------
// output slots per thread when there is enough work for several waves of blocks: the safegcd inversion runs ~170 000
// instructions against ~9 700 per addition, so a batch of 64 spends a fifth of the instruction stream inverting
constexpr unsigned TREE_BATCH = 256;
constexpr uint64_t TREE_MIN_ENTRIES = (uint64_t)1 << 22;   // windows x points from which the tree is the default form

G753_HD uint32_t tree_start(uint32_t s0, uint32_t t, unsigned level) {
  uint64_t s = s0;
  for (unsigned k = 0; k < level; k++) s = (s + t) >> 1;
  return (uint32_t)s;
}
G753_HD uint32_t tree_count(uint32_t c0, unsigned level) {
  return level >= 32 ? (c0 ? 1u : 0u) : (uint32_t)(((uint64_t)c0 + ((1ull << level) - 1)) >> level);
}
// fullest bucket (grid-stride; one atomic per thread at most)
static __global__ void k_tree_max(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ ends, unsigned NB,
                                  uint32_t* __restrict__ max_count) {
  uint32_t m = 0;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < NB; t += (size_t)gridDim.x * blockDim.x) {
    const uint32_t c = ends[t] - offsets[t];
    m = c > m ? c : m;
  }
  if (m > 1 && m > *(volatile uint32_t*)max_count) atomicMax(max_count, m);
}

struct TreeGeo {
  const uint32_t* offsets;   // [NB] start of bucket t's run within its row of sorted[]
  const uint32_t* ends;      // [NB] end of the run
  unsigned NB, len;          // buckets in all rows, buckets per row (B + 1)
  uint32_t row_cap;          // entries per row of sorted[]
};

// position of output slot j of level `level`: bucket t (largest t with S_level[t] <= j) and that
// bucket's geometry one level down.  The walk is sequential (up in phase 1, down in phase 2) and the
// offsets of the NEXT bucket are loaded one bucket ahead, so crossing a bucket boundary does not wait
// for memory.
struct TreeWalk {
  TreeGeo g;
  unsigned level;            // output level (>= 1)
  unsigned t;
  uint32_t s_in, c_in, s_out, c_out;
  uint32_t s_next;           // walking up: S_level[t + 1] (2^32 - 1 after the last bucket)
  uint32_t off_next, pre_a, pre_b;   // up: offsets[t+1], ends[t+1], offsets[t+2]; down: -, offsets[t-1], ends[t-1]
  G753_D uint32_t start_of(unsigned b, uint32_t off, unsigned lv) const {
    return tree_start((uint32_t)(b / g.len) * g.row_cap + off, b, lv);
  }
  G753_D void set(unsigned b, uint32_t off, uint32_t end) {
    t = b;
    s_in = start_of(b, off, level - 1);
    s_out = (uint32_t)(((uint64_t)s_in + b) >> 1);
    c_in = tree_count(end - off, level - 1);
    c_out = (c_in + 1) >> 1;
  }
  G753_D void preload_up() {
    pre_a = t + 1 < g.NB ? g.ends[t + 1] : 0u;
    pre_b = t + 2 < g.NB ? g.offsets[t + 2] : 0u;
  }
  G753_D void start_up(uint32_t j) {
    unsigned lo = 0, hi = g.NB;   // invariant: S[lo] <= j (S[0] = 0), S[hi] > j or hi == NB
    while (hi - lo > 1) {
      const unsigned mid = lo + (hi - lo) / 2;
      if (start_of(mid, g.offsets[mid], level) <= j) lo = mid; else hi = mid;
    }
    set(lo, g.offsets[lo], g.ends[lo]);
    off_next = lo + 1 < g.NB ? g.offsets[lo + 1] : 0u;
    s_next = lo + 1 < g.NB ? start_of(lo + 1, off_next, level) : 0xffffffffu;
    preload_up();
  }
  G753_D void seek_up(uint32_t j) {
    while (j >= s_next) {
      const unsigned b = t + 1;
      set(b, off_next, pre_a);
      off_next = pre_b;
      s_next = b + 1 < g.NB ? start_of(b + 1, off_next, level) : 0xffffffffu;
      preload_up();
    }
  }
  G753_D void preload_down() {
    pre_a = t > 0 ? g.offsets[t - 1] : 0u;
    pre_b = t > 0 ? g.ends[t - 1] : 0u;
  }
  G753_D void start_down() { preload_down(); }   // from the bucket the walk up ended in
  G753_D void seek_down(uint32_t j) {
    while (j < s_out) {
      set(t - 1, pre_a, pre_b);
      preload_down();
    }
  }
};

// what output slot j is made of, in two steps so that the loads of one step are issued an iteration before
// the next step needs them: TreeIdx (entry indices; at level 1 the two sorted[] entries, in flight) and
// TreeSlot (operand addresses)
struct TreeIdx {
  uint32_t a, b;   // level 1: sorted[src], sorted[src + 1]; above: a = src
  uint32_t t;      // bucket
  unsigned kind;   // 0 = nothing to do, 1 = single element carried over, 2 = pair; bit 6: the pair is the last of its bucket
};
struct TreeSlot {
  const Fq* p1;
  const Fq* p2;
  uint32_t t;
  unsigned kind;   // as TreeIdx; bit 4 / bit 5: negate y of p1 / p2
};
G753_D TreeIdx tree_idx(const TreeWalk& w, uint32_t j, const uint32_t* __restrict__ sorted) {
  TreeIdx s;
  s.a = s.b = 0;
  s.t = w.t;
  s.kind = 0;
  const uint32_t l = j - w.s_out;
  if (l >= w.c_out || w.c_in <= 1) return s;   // a hole, or a bucket that is finished (its sum is in buckets[t])
  const uint32_t src = w.s_in + 2 * l;
  const bool pair = 2 * l + 1 < w.c_in;
  s.kind = pair ? (w.c_in == 2 ? 66u : 2u) : 1u;
  if (w.level == 1) {
    s.a = sorted[src];
    if (pair) s.b = sorted[src + 1];
  } else {
    s.a = src;
  }
  return s;
}
template <int K>
G753_D TreeSlot tree_slot(const TreeIdx& i, bool lvl0, const Fq* __restrict__ bases, const Fq* __restrict__ in) {
  TreeSlot s;
  s.kind = i.kind;
  s.t = i.t;
  s.p1 = s.p2 = nullptr;
  if (i.kind == 0) return s;
  if (lvl0) {
    s.p1 = bases + (size_t)(i.a & 0x7fffffffu) * (2 * K);
    s.p2 = bases + (size_t)(i.b & 0x7fffffffu) * (2 * K);
    s.kind |= ((i.a >> 31) << 4) | ((i.b >> 31) << 5);
  } else {
    s.p1 = in + (size_t)i.a * (2 * K);
    s.p2 = s.p1 + 2 * K;
  }
  return s;
}
template <int BYTES>
G753_D void tree_prefetch(const void* p) {
#if defined(__CUDA_ARCH__)
  const char* q = (const char*)p;
#pragma unroll
  for (int off = 0; off < BYTES; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + off));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(q + BYTES - 1));
#else
  (void)p;
#endif
}

// Six element slots per column (x1, y1, x2, y2, T, INV; the denominator overwrites x2, x3 lands in y2's slot): on
// the prime-field curves 72 KB per 128-thread block, so THREE blocks share an SM (12 warps; the XYZZ kernel's
// eight slots allow two).  The hot loop calls ONE multiplier body (22 KB; squarings go through it too), which
// the 32 KB L1.5 instruction cache holds - with the dedicated squaring body beside it ncu showed a quarter of
// the issue slots waiting for instructions.
// warp votes of the round's loops (the TEST-ONLY host build runs one lane at a time)
G753_D unsigned tree_lanes() {
#if defined(__CUDA_ARCH__)
  return __activemask();
#else
  return 1u;
#endif
}
G753_D bool tree_any(unsigned lanes, bool v) {
#if defined(__CUDA_ARCH__)
  return __any_sync(lanes, v) != 0;
#else
  (void)lanes;
  return v;
#endif
}
G753_D void tree_join(unsigned lanes) {
#if defined(__CUDA_ARCH__)
  __syncwarp(lanes);
#else
  (void)lanes;
#endif
}

template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS, SC::M::K == 1 ? 3 : 1)
k_tree_round(const Fq* __restrict__ bases, const uint32_t* __restrict__ sorted, TreeGeo geo,
             const uint32_t* max_count, unsigned level, unsigned batch, const Fq* __restrict__ in,
             Fq* __restrict__ out, Fq* __restrict__ points) {
  typedef typename SC::M M;
  typedef typename M::T L;
  constexpr int K = M::K;
  enum { X1 = 0, Y1 = K, X2 = 2 * K, Y2 = 3 * K, T = 4 * K, INV = 5 * K, TMP = 6 * K };
  constexpr int ONE = M::NTMP >= K ? TMP : X2;   // where the doubling builds the curve coefficient a
  constexpr int EL = (int)(K * sizeof(Fq));      // bytes of one coordinate
  if (tree_count(*max_count, level - 1) <= 1) return;   // every bucket is down to one element already
  if (L::idle()) return;
  const uint64_t first64 = (uint64_t)L::item() * batch;
  TreeWalk w;
  w.g = geo;
  w.level = level;
  {
    const unsigned b = geo.NB - 1;
    w.set(b, geo.offsets[b], geo.ends[b]);
    if (first64 >= (uint64_t)w.s_out + w.c_out) return;   // beyond the last slot of the level
  }
  const uint32_t first = (uint32_t)first64;
  const uint64_t stop64 = first64 + batch;
  const uint32_t last = (uint32_t)(stop64 > 0xfffffffeull ? 0xfffffffeull : stop64);   // exclusive
  const bool lvl0 = level == 1;   // inputs are key points: never at infinity, y negated by the sign bit
  // max_count[1 + L] != 0: level L holds a point at infinity, stored as (0, 0) (a cancellation - with real keys
  // practically never); only then are the operands of the next level tested for it
  uint32_t* const inf_seen = (uint32_t*)max_count + 1;
  const bool may_inf = !lvl0 && inf_seen[level - 1] != 0;
  if (may_inf) inf_seen[level] = 1;   // infinity operands pass through

  auto fix_y = [&](const TreeSlot& s) {
    if (s.kind & 16u) M::neg(Y1, Y1);
    if (s.kind & 32u) M::neg(Y2, Y2);
  };
  auto load_y = [&](const TreeSlot& s) {
    const int d[2] = {Y1, Y2};
    const Fq* const g[2] = {s.p1 + K, s.p2 + K};
    t_ldg_many<M, 2>(d, g);
    fix_y(s);
  };
  // with (x1, x2) in X1, X2 (and, when have_y, the sign-corrected y in Y1, Y2):
  // 0 = ordinary addition, 1 = doubling, 2 = cancellation, 3 = first operand at infinity (result = second),
  // 4 = second at infinity (result = first).  Kinds 0, 1 leave the DENOMINATOR in X2 (x2 - x1, or 2 y1).
  auto classify = [&](const TreeSlot& s, bool have_y) -> int {
    if (may_inf) {
      const bool z1 = M::is_zero(X1), z2 = M::is_zero(X2);
      if (z1 || z2) {
        if (!have_y) load_y(s);
        if (z1 && M::is_zero(Y1)) return 3;
        if (z2 && M::is_zero(Y2)) return 4;
        have_y = true;
      }
    }
    M::sub(X2, X2, X1);
    if (!M::is_zero(X2)) return 0;
    if (!have_y) load_y(s);
    M::sub(X2, Y2, Y1);
    if (!M::is_zero(X2) || M::is_zero(Y1)) return 2;
    M::dbl(X2, Y1);
    return 1;
  };
  auto prefetch_x = [&](const TreeSlot& s) {
    if ((s.kind & 3u) != 2u) return;
    tree_prefetch<EL>(s.p1);
    tree_prefetch<EL>(s.p2);
  };
  auto prefetch_xy = [&](const TreeSlot& s, const Fq* cell) {
    if ((s.kind & 3u) == 0u) return;
    if ((s.kind & 3u) == 2u) {
      if (lvl0) {
        tree_prefetch<2 * EL>(s.p1);
        tree_prefetch<2 * EL>(s.p2);
      } else {
        tree_prefetch<4 * EL>(s.p1);
      }
      tree_prefetch<EL>(cell);
    } else {
      tree_prefetch<2 * EL>(s.p1);
    }
  };
  // prime fields: the dedicated squaring is fewer limb products but MORE instructions (2370 against 1362) and a
  // second 38 KB body for the instruction cache; the towers' squarings do save products
  auto square = [&](int d, int a) {
    if (K == 1) M::mul(d, a, a, TMP); else M::sqr(d, a, TMP);
  };
  // (x, y) in slots sx, sy: to the cell of the output level, or - the last addition of a bucket - to the bucket
  // itself as XYZZ (x, y, 1, 1); a sum at infinity leaves the bucket's all-zero limbs (ZZ == 0) alone
  auto put = [&](const TreeSlot& s, Fq* cell, int sx, int sy, int one, bool maybe_inf) {
    if (s.kind & 64u) {
      if (maybe_inf && M::is_zero(sx) && M::is_zero(sy)) return;
      Fq* o = points + (size_t)s.t * (4 * K);
      M::set_one(one);
      M::stg(o, sx);
      M::stg(o + K, sy);
      M::stg(o + 2 * K, one);
      M::stg(o + 3 * K, one);
    } else {
      M::stg(cell, sx);
      M::stg(cell + K, sy);
    }
  };
  TreeIdx none;
  none.a = none.b = none.t = 0;
  none.kind = 0;

  // ---- phase 1: denominators and their running product -----------------------------------------
  // Entry indices are read two slots ahead (their loads fly during a slot's arithmetic), turned into addresses
  // and prefetched into L2 one slot ahead.  A lane steps over the slots that are no pair (holes, finished buckets,
  // carried elements) on its own and meets the other lanes of its warp again at its next PAIR: on the upper levels
  // most slots are no pair, and a warp that walked them in lockstep paid a full addition for every slot in which
  // any of its lanes had one (measured: 1 ns per slot on every level, whatever the share of pairs).
  bool any = false, work = false;
  uint32_t lead = 0;   // slot of the first denominator: its prefix is 1 and is not stored
  const unsigned lanes = tree_lanes();   // the lanes of this warp that have slots
  {
    w.start_up(first);
    TreeSlot cur = tree_slot<K>(tree_idx(w, first, sorted), lvl0, bases, in);
    TreeSlot nxt = tree_slot<K>(none, lvl0, bases, in);
    TreeIdx ahead = none;
    if (first + 1 < last) {
      w.seek_up(first + 1);
      nxt = tree_slot<K>(tree_idx(w, first + 1, sorted), lvl0, bases, in);
      prefetch_x(nxt);
    }
    if (first + 2 < last) {
      w.seek_up(first + 2);
      ahead = tree_idx(w, first + 2, sorted);
    }
    uint32_t j = first;
    auto advance = [&]() {
      cur = nxt;
      nxt = tree_slot<K>(ahead, lvl0, bases, in);
      prefetch_x(nxt);
      j++;
      ahead = none;
      if (j + 2 < last) {
        w.seek_up(j + 2);
        ahead = tree_idx(w, j + 2, sorted);
      }
    };
    // every lane of the warp stays in the loop until the last one is through, and the lanes JOIN before each pair:
    // left to itself the compiler does not reconverge the lanes after the stepping loop, and a warp in pieces
    // runs the multiplier once per piece
    while (tree_any(lanes, j < last)) {
      while (j < last && (cur.kind & 3u) != 2u) {
        work = work || (cur.kind & 3u) != 0u;
        advance();
      }
      tree_join(lanes);
      if (j < last) {
        work = true;
        const int d[2] = {X1, X2};
        const Fq* const g[2] = {cur.p1, cur.p2};
        t_ldg_many<M, 2>(d, g);
        if (classify(cur, false) < 2) {
          if (!any) {
            M::copy(INV, X2);
            any = true;
            lead = j;
          } else {
            M::stg(out + (size_t)j * (2 * K), INV);
            M::mul(INV, INV, X2, TMP);
          }
        }
        advance();
      }
    }
  }
  tree_join(lanes);
  if (tree_any(lanes, any)) {        // (together: the lanes that returned early would split the warp for good)
    if (!any) M::set_one(INV);
    M::inv(INV, INV, TMP);
  }
  if (!tree_any(lanes, work)) return;   // holes and finished buckets only
  // ---- phase 2: walk back, peel the individual inverses off, finish the additions ----------------
  {
    w.start_down();                  // the walk up ended in the bucket of slot last - 1
    TreeSlot cur = tree_slot<K>(tree_idx(w, last - 1, sorted), lvl0, bases, in);
    TreeSlot nxt = tree_slot<K>(none, lvl0, bases, in);
    TreeIdx ahead = none;
    if (last - first >= 2) {
      w.seek_down(last - 2);
      nxt = tree_slot<K>(tree_idx(w, last - 2, sorted), lvl0, bases, in);
      prefetch_xy(nxt, out + (size_t)(last - 2) * (2 * K));
    }
    if (last - first >= 3) {
      w.seek_down(last - 3);
      ahead = tree_idx(w, last - 3, sorted);
    }
    uint32_t left = work ? last - first : 0u;   // slots not yet done (the current one is first + left - 1); a lane
                                                // without pairs or carried elements only keeps its warp company
    auto advance = [&]() {
      cur = nxt;
      nxt = tree_slot<K>(ahead, lvl0, bases, in);
      left--;
      if (left >= 2) prefetch_xy(nxt, out + (size_t)(first + left - 2) * (2 * K));
      ahead = none;
      if (left >= 3) {
        w.seek_down(first + left - 3);
        ahead = tree_idx(w, first + left - 3, sorted);
      }
    };
    while (tree_any(lanes, left > 0)) {
      while (left > 0 && (cur.kind & 3u) != 2u) {
        if ((cur.kind & 3u) == 1u) {          // carried over
          Fq* cell = out + (size_t)(first + left - 1) * (2 * K);
          const int d[2] = {X1, Y1};
          const Fq* const g[2] = {cur.p1, cur.p1 + K};
          t_ldg_many<M, 2>(d, g);
          if (cur.kind & 16u) M::neg(Y1, Y1);
          M::stg(cell, X1);
          M::stg(cell + K, Y1);
        }
        advance();
      }
      tree_join(lanes);
      if (left == 0) continue;
      const uint32_t j = first + left - 1;
      Fq* cell = out + (size_t)j * (2 * K);
      if (any && j != lead) {                 // with the product of the denominators before this one
        const int d[5] = {X1, Y1, X2, Y2, T};
        const Fq* const g[5] = {cur.p1, cur.p1 + K, cur.p2, cur.p2 + K, cell};
        t_ldg_many<M, 5>(d, g);
      } else {
        const int d[4] = {X1, Y1, X2, Y2};
        const Fq* const g[4] = {cur.p1, cur.p1 + K, cur.p2, cur.p2 + K};
        t_ldg_many<M, 4>(d, g);
      }
      fix_y(cur);
      const int kind = classify(cur, true);
      if (kind == 2) {
        if (!(cur.kind & 64u)) {
          M::set_zero(X1);
          M::stg(cell, X1);
          M::stg(cell + K, X1);
          inf_seen[level] = 1;
        }
      } else if (kind == 3) {
        put(cur, cell, X2, Y2, X1, true);
      } else if (kind == 4) {
        put(cur, cell, X1, Y1, X2, true);
      } else {
        if (j != lead) {
          M::mul(T, T, INV, TMP);             // 1 / den
          M::mul(INV, INV, X2, TMP);          // inverse of the product of the earlier ones
        } else {
          M::copy(T, INV);
        }
        if (kind == 1) {                      // lambda = (3 x1^2 + a) / (2 y1), x2 = x1
          square(Y2, X1);
          M::dbl(X2, Y2);
          M::add(Y2, Y2, X2);
          M::set_one(ONE);
          SC::mul_by_a(X2, ONE);
          M::add(Y2, Y2, X2);
          M::set_zero(X2);                    // "x2 - x1" for the x3 below
        } else {
          M::sub(Y2, Y2, Y1);
        }
        M::mul(T, T, Y2, TMP);                // lambda
        square(Y2, T);
        M::sub(Y2, Y2, X1);
        M::sub(Y2, Y2, X1);
        M::sub(Y2, Y2, X2);                   // x3 = lambda^2 - x1 - (x1 + (x2 - x1))
        M::sub(X1, X1, Y2);
        M::mul(X1, X1, T, TMP);
        M::sub(Y1, X1, Y1);                   // y3 = lambda (x1 - x3) - y1
        put(cur, cell, Y2, Y1, X2, false);
      }
      advance();
    }
  }
}
template <class M>
constexpr int tree_slots() { return 6 * M::K + M::NTMP; }

// buckets with ONE entry never enter a round: buckets[t] = that key point (sign applied) as XYZZ (x, y, 1, 1).
// (Fuller buckets are written by the round that adds their last pair; empty ones keep the all-zero limbs the
// host wrote, ZZ == 0.)
template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS)
k_tree_finish(const Fq* __restrict__ bases, const uint32_t* __restrict__ sorted, TreeGeo geo, Fq* __restrict__ points) {
  typedef typename SC::M M;
  typedef typename M::T L;
  constexpr int K = M::K;
  enum { X = 0, Y = K, ONE = 2 * K };
  if (L::idle()) return;
  const unsigned t = L::item();
  if (t >= geo.NB) return;
  if (geo.ends[t] - geo.offsets[t] != 1) return;
  const uint32_t e = sorted[(size_t)(t / geo.len) * geo.row_cap + geo.offsets[t]];
  const Fq* q = bases + (size_t)(e & 0x7fffffffu) * (2 * K);
  M::ldg(X, q);
  M::ldg(Y, q + K);
  if (e >> 31) M::neg(Y, Y);
  M::set_one(ONE);
  Fq* o = points + (size_t)t * (4 * K);
  M::stg(o, X);
  M::stg(o + K, Y);
  M::stg(o + 2 * K, ONE);
  M::stg(o + 3 * K, ONE);
}

// Buckets that were cut into several items: points[t] = sum of their partial results
// points[NB + item_off[t] + j], j < item_cnt[t].  A bucket can hold a large share of all points
// (the top window of a 753-bit scalar has only a few significant bits; real witnesses are full of
// 0 / 1 / small values), so the partials are summed by an in-place tree: at level l (stride
// s = FIX_FAN^l) the thread of partial j, j % (FIX_FAN s) == 0, adds partials j + s, j + 2s, ...
// into its own - every thread writes only inside its own segment - and after the last level
// partial 0 holds the bucket's sum.  One thread per item id; part_bucket[] maps ids to buckets.
constexpr unsigned FIX_FAN = 8;
template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS)
k_bucket_fixup_level(const uint32_t* __restrict__ item_cnt, const uint32_t* __restrict__ item_off,
                     const uint32_t* __restrict__ part_bucket, const uint32_t* __restrict__ item_total, unsigned NB,
                     unsigned stride, Fq* __restrict__ points) {
  typedef EcS<SC> E;
  if (E::M::T::idle()) return;
  unsigned q = E::M::T::item();
  if (q >= *item_total) return;
  const uint32_t t = part_bucket[q];
  const uint32_t cnt = item_cnt[t], first = item_off[t];
  const uint32_t j = q - first;
  if (cnt < 2 || j % (FIX_FAN * stride) != 0 || j + stride >= cnt) return;
  Fq* mine = points + (size_t)(NB + q) * E::PT;
  E::ldg(0, mine);
  for (unsigned k = 1; k < FIX_FAN; k++) {
    const uint32_t jj = j + k * stride;
    if (jj >= cnt) break;
    E::add_g(0, points + (size_t)(NB + first + jj) * E::PT, E::PT);
  }
  E::stg(mine, 0);
}
// points[t] = partial 0 of every multi-item bucket
template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS)
k_bucket_fixup(const uint32_t* __restrict__ item_cnt, const uint32_t* __restrict__ item_off,
               unsigned NB, Fq* __restrict__ points) {
  typedef EcS<SC> E;
  if (E::M::T::idle()) return;
  unsigned t = E::M::T::item();
  if (t >= NB) return;
  if (item_cnt[t] < 2) return;
  E::ldg(0, points + (size_t)(NB + item_off[t]) * E::PT);
  E::stg(points + (size_t)t * E::PT, 0);
}

// ------------------------------------------------------------------------------------
// K6: bucket reduction  S_w = sum_b b * B_{w,b}
// One level turns  G(X, Y, f) = sum_i Y_i + f * sum_i i * X_i  over n_in entries into the same
// problem over n_out = ceil(n_in / s) entries with f' = f * s:
//   R_j  = sum_{i in seg j} X_i
//   Y'_j = sum_{i in seg j} Y_i + f * sum_{i in seg j} (i - j s) X_i        (running sums)
// Row w of X / Y starts at w * x_stride / w * y_stride (in points).
// ------------------------------------------------------------------------------------
template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS)
k_reduce_level(const Fq* __restrict__ X, size_t x_stride, const Fq* __restrict__ Y, size_t y_stride,
               unsigned n_in, unsigned log2f, unsigned W, unsigned n_out, Fq* __restrict__ R,
               Fq* __restrict__ Yout) {
  typedef EcS<SC> E;
  constexpr int RUN = 0, ACC = E::PT, SCR = 2 * E::PT;
  unsigned t = E::M::T::item();
  if (t >= W * n_out) return;
  unsigned w = t / n_out, j = t % n_out;
  unsigned lo = j * REDUCE_SEG;
  unsigned hi = lo + REDUCE_SEG < n_in ? lo + REDUCE_SEG : n_in;
  const Fq* x = X + (size_t)w * x_stride * E::PT;
  E::set_inf(RUN);
  E::set_inf(ACC);
  for (unsigned i = hi - 1; i > lo; i--) {
    E::add_g(RUN, x + (size_t)i * E::PT, SCR);
    E::add(ACC, RUN, SCR);
  }
  E::add_g(RUN, x + (size_t)lo * E::PT, SCR);
  for (unsigned k = 0; k < log2f; k++) E::dbl(ACC, SCR);
  if (Y != nullptr) {
    const Fq* y = Y + (size_t)w * y_stride * E::PT;
    for (unsigned i = lo; i < hi; i++) E::add_g(ACC, y + (size_t)i * E::PT, SCR);
  }
  E::stg(R + (size_t)t * E::PT, RUN);
  E::stg(Yout + (size_t)t * E::PT, ACC);
}

// ---- reduction tail ----------------------------------------------------------------------------
// Once a row is down to n_in <= REDUCE_TAIL_MAX entries the remaining running-sum levels are pure
// latency (a dozen serial additions per level on a handful of threads).  The tail computes
//   G = sum_i Y_i + 2^log2f * sum_i i * X_i
// by bit planes instead: S_t = sum_{i : bit t of i set} X_i for every bit t, and sum_i Y_i, each
// by a pairwise tree (one addition of latency per halving, all planes and rows in one launch),
// then 2^(t + log2f) S_t per plane in parallel and a last tree over the planes.
constexpr unsigned REDUCE_TAIL_MAX = 16384;

// first halving: T[(row * P + plane) * half + p] = masked X (or Y on the last plane) of entries 2p, 2p+1
template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS)
k_tail_planes(const Fq* __restrict__ X, const Fq* __restrict__ Y, unsigned n_in, unsigned nbits, unsigned rows,
              Fq* __restrict__ T) {
  typedef EcS<SC> E;
  const unsigned P = nbits + 1, half = (n_in + 1) / 2;
  unsigned t = E::M::T::item();
  if (t >= rows * P * half) return;
  const unsigned p = t % half, plane = (t / half) % P, row = t / (half * P);
  const Fq* x = (plane < nbits ? X : Y) + (size_t)row * n_in * E::PT;
  E::set_inf(0);
  for (unsigned k = 0; k < 2; k++) {
    const unsigned i = 2 * p + k;
    if (i < n_in && (plane == nbits || ((i >> plane) & 1))) E::add_g(0, x + (size_t)i * E::PT, E::PT);
  }
  E::stg(T + (size_t)t * E::PT, 0);
}
// out[g * half + p] = in[g * m_in + 2p] + in[g * m_in + 2p + 1]
template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS)
k_pair_sum(const Fq* __restrict__ in, unsigned groups, unsigned m_in, Fq* __restrict__ out) {
  typedef EcS<SC> E;
  const unsigned half = (m_in + 1) / 2;
  unsigned t = E::M::T::item();
  if (t >= groups * half) return;
  const unsigned p = t % half, g = t / half;
  const Fq* x = in + (size_t)g * m_in * E::PT;
  E::set_inf(0);
  E::add_g(0, x + (size_t)(2 * p) * E::PT, E::PT);
  if (2 * p + 1 < m_in) E::add_g(0, x + (size_t)(2 * p + 1) * E::PT, E::PT);
  E::stg(out + (size_t)t * E::PT, 0);
}
// Z[row * P + plane] = 2^(plane + log2f) * S[row * P + plane] (the Y plane, plane == nbits, is copied)
template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS)
k_tail_scale(const Fq* __restrict__ S, unsigned nbits, unsigned rows, unsigned log2f, Fq* __restrict__ Z) {
  typedef EcS<SC> E;
  const unsigned P = nbits + 1;
  unsigned t = E::M::T::item();
  if (t >= rows * P) return;
  const unsigned plane = t % P;
  E::ldg(0, S + (size_t)t * E::PT);
  if (plane < nbits)
    for (unsigned k = 0; k < plane + log2f; k++) E::dbl(0, E::PT);
  E::stg(Z + (size_t)t * E::PT, 0);
}

// Horner fold over the W window sums (stride between windows given, in points), then convert
// to the reference's homogeneous projective layout.  One thread: W*c doublings are a serial chain.
template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS)
k_window_combine(const Fq* __restrict__ sums, unsigned stride, unsigned W, unsigned c,
                 Fq* __restrict__ out_xyz) {
  typedef EcS<SC> E;
  if (E::M::T::item() != 0) return;
  E::set_inf(0);
  for (int w = (int)W - 1; w >= 0; w--) {
    if (w != (int)W - 1)
      for (unsigned k = 0; k < c; k++) E::dbl(0, E::PT);
    E::add_g(0, sums + (size_t)w * stride * E::PT, E::PT);
  }
  E::to_projective(0, E::PT);
  for (int i = 0; i < 3; i++) E::M::stg(out_xyz + i * E::K, i * E::K);
}

#if !defined(G753_HOST_EMUL)
// ---- warp-cooperative forms of the latency-bound steps (coop.cuh): one WARP per point ---------------
constexpr int COOP_WARPS = 4;                 // warps (points) per block
constexpr unsigned COOP_MAX_ITEMS = 4096;     // above this many independent points one thread per point wins

// Horner fold over the W window sums, then the reference's homogeneous projective layout: the serial
// chain of (W - 1) c doublings of variable_base.rs:72-82 at ~3 us per doubling instead of ~40 us
template <int GID>
__global__ void __launch_bounds__(32)
k_window_combine_coop(const Fq* __restrict__ sums, unsigned stride, unsigned W, unsigned c, Fq* __restrict__ out_xyz) {
  constexpr int PT = 4 * CoopGroup<GID>::K;
  CoopEc<GID> ec;
  ec.init((uint32_t*)g_slots);
  ec.set_inf();
  for (int w = (int)W - 1; w >= 0; w--) {
    if (w != (int)W - 1)
      for (unsigned k = 0; k < c; k++) ec.dbl();
    ec.add_g(sums + (size_t)w * stride * PT);
  }
  ec.store_projective(out_xyz);
}
// sum of `count` homogeneous projective points (the multi-GPU fold)
template <int GID>
__global__ void __launch_bounds__(32)
k_points_sum_coop(const Fq* __restrict__ pts, unsigned count, Fq* __restrict__ out_xyz) {
  constexpr int K = CoopGroup<GID>::K;
  CoopEc<GID> ec;
  ec.init((uint32_t*)g_slots);
  ec.set_inf();
  for (unsigned k = 0; k < count; k++) ec.add_projective_g(pts + (size_t)k * 3 * K);
  ec.store_projective(out_xyz);
}
// k_pair_sum with one warp per output point
template <int GID>
__global__ void __launch_bounds__(32 * COOP_WARPS)
k_pair_sum_coop(const Fq* __restrict__ in, unsigned groups, unsigned m_in, Fq* __restrict__ out) {
  constexpr int PT = 4 * CoopGroup<GID>::K;
  CoopEc<GID> ec;
  ec.init((uint32_t*)g_slots);
  const unsigned half = (m_in + 1) / 2;
  const unsigned t = blockIdx.x * COOP_WARPS + (threadIdx.x >> 5);
  if (t >= groups * half) return;
  const unsigned p = t % half, g = t / half;
  const Fq* x = in + (size_t)g * m_in * PT;
  ec.set_inf();
  ec.add_g(x + (size_t)(2 * p) * PT);
  if (2 * p + 1 < m_in) ec.add_g(x + (size_t)(2 * p + 1) * PT);
  ec.store_xyzz(out + (size_t)t * PT);
}
// k_tail_scale with one warp per plane
template <int GID>
__global__ void __launch_bounds__(32 * COOP_WARPS)
k_tail_scale_coop(const Fq* __restrict__ S, unsigned nbits, unsigned rows, unsigned log2f, Fq* __restrict__ Z) {
  constexpr int PT = 4 * CoopGroup<GID>::K;
  CoopEc<GID> ec;
  ec.init((uint32_t*)g_slots);
  const unsigned P = nbits + 1;
  const unsigned t = blockIdx.x * COOP_WARPS + (threadIdx.x >> 5);
  if (t >= rows * P) return;
  const unsigned plane = t % P;
  ec.set_inf();
  ec.add_g(S + (size_t)t * PT);
  if (plane < nbits)
    for (unsigned k = 0; k < plane + log2f; k++) ec.dbl();
  ec.store_xyzz(Z + (size_t)t * PT);
}
#endif

template <class SC>
__global__ void k_write_infinity(Fq* __restrict__ out_xyz) {
  typedef EcS<SC> E;
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (int i = 0; i < 3 * E::K; i++) g_st(out_xyz + i, i == E::K ? fq_one<E::M::FIELD>() : fq_zero<E::M::FIELD>());
}

// sum of `count` projective points (the multi-GPU fold; count is the number of ranks)
template <class SC>
__global__ void __launch_bounds__(SC::M::T::THREADS)
k_points_sum(const Fq* __restrict__ pts, unsigned count, Fq* __restrict__ out_xyz) {
  typedef EcS<SC> E;
  constexpr int TOT = 0, CUR = E::PT, SCR = 2 * E::PT;
  if (E::M::T::item() != 0) return;
  E::set_inf(TOT);
  for (unsigned k = 0; k < count; k++) {
    E::from_projective_g(CUR, pts + (size_t)k * 3 * E::K, SCR);
    E::add(TOT, CUR, SCR);
  }
  E::to_projective(TOT, SCR);
  for (int i = 0; i < 3; i++) E::M::stg(out_xyz + i * E::K, TOT + i * E::K);
}

// ------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------
struct MsmWorkspace {
  unsigned n_chunks, nb_chunks, level_entries, max_items;
  size_t tail_points;  // points per ping-pong buffer of the reduction tail
  size_t row_cap;  // entries per row of sorted[]
  size_t tree_odd, tree_even;  // affine points in the two ping-pong arrays of the addition tree (0: not used)
  size_t total;
};
// upper bound on the slots of level L + 1 of the addition tree given the bound on level L
static inline size_t tree_level_bound(size_t prev, size_t NB) { return (prev + NB) / 2 + 1; }

template <int GID>
static inline MsmWorkspace msm_workspace(const MsmPlan& pl, size_t n, bool tree = false) {
  constexpr size_t PT_BYTES = sizeof(Fq) * 4 * MsmCfg<GID>::K;
  MsmWorkspace s;
  const size_t len = (size_t)pl.B + 1;
  const size_t NB = pl.rows * len;
  s.row_cap = (size_t)pl.copies * n;
  s.n_chunks = div_up(len, SCAN_CHUNK);
  s.nb_chunks = div_up(NB, SCAN_CHUNK);
  s.max_items = (unsigned)(NB + (size_t)pl.rows * s.row_cap / ITEM_LEN);
  unsigned entries = 0;
  for (size_t m = len; m > 1;) {
    m = div_up(m, REDUCE_SEG);
    entries += (unsigned)m;
  }
  if (entries == 0) entries = 1;
  s.level_entries = entries;
  // reduction tail: two ping-pong buffers of rows x (bits + 1) planes x ceil(n_in / 2) points, n_in being
  // the first level size <= REDUCE_TAIL_MAX (at least one running-sum level always runs first)
  s.tail_points = 0;
  {
    size_t m = div_up(len, REDUCE_SEG);
    while (m > REDUCE_TAIL_MAX) m = div_up(m, REDUCE_SEG);
    if (m > 1) {
      unsigned nbits = 0;
      while (((size_t)1 << nbits) < m) nbits++;
      s.tail_points = (size_t)pl.rows * (nbits + 1) * ((m + 1) / 2);
    }
  }
  size_t t = 0;
  t += Carver::pad(sizeof(uint32_t) * pl.W * n);                        // digits
  t += Carver::pad(sizeof(uint32_t) * pl.rows * s.row_cap);             // sorted
  t += Carver::pad(sizeof(uint32_t) * NB) * 5;                          // hist, offsets, cursor, item_cnt, item_off
  t += Carver::pad(sizeof(uint32_t) * pl.rows * s.n_chunks);            // chunk sums
  t += Carver::pad(sizeof(uint32_t) * s.nb_chunks);                     // item-count chunk sums
  t += Carver::pad(sizeof(uint32_t) * ITEM_LEN * ITEM_REP) * 2 + 512;   // length counters + total
  t += Carver::pad(sizeof(MsmItem) * s.max_items);
  t += Carver::pad(sizeof(uint32_t) * s.max_items);                     // item id -> bucket
  t += Carver::pad(PT_BYTES * (NB + s.max_items));                      // buckets + item partials
  t += Carver::pad(PT_BYTES * pl.rows * entries) * 2;                   // reduction levels
  t += Carver::pad(PT_BYTES * (s.tail_points + 1)) * 2;                 // reduction tail
  s.tree_odd = s.tree_even = 0;
  if (tree) {                                                            // levels 1, 3, ... / 2, 4, ... of the addition tree
    s.tree_odd = tree_level_bound((size_t)pl.rows * s.row_cap, NB);
    s.tree_even = tree_level_bound(s.tree_odd, NB);
    if (s.tree_odd < NB + 4) s.tree_odd = NB + 4;     // the bounds of the later levels tend to NB + 2
    if (s.tree_even < NB + 4) s.tree_even = NB + 4;
    t += Carver::pad(PT_BYTES / 2 * s.tree_odd) + Carver::pad(PT_BYTES / 2 * s.tree_even) + 512;
  }
  s.total = t + 8192;
  return s;
}

struct MsmHooks {  // phase timing hooks; the emulation build leaves them null
  void (*mark)(void* user, int phase) = nullptr;
  void (*wait_chunk)(void* user, int chunk) = nullptr;  // order the stream after the upload of scalar chunk j
  unsigned scalar_chunks = 1;
  void* user = nullptr;
  uint64_t* launches = nullptr;
  unsigned* form = nullptr;   // out: accumulation form that ran (0 = XYZZ running sums, 1 = affine addition tree)
};

#define G753_MSM_LAUNCH(hooks, ...)            \
  do {                                         \
    G753_LAUNCH(__VA_ARGS__);                  \
    if ((hooks).launches) ++*(hooks).launches; \
  } while (0)
#define G753_MSM_LAUNCH_SMEM(hooks, ...)       \
  do {                                         \
    G753_LAUNCH_SMEM(__VA_ARGS__);             \
    if ((hooks).launches) ++*(hooks).launches; \
  } while (0)

template <class E, int T>
constexpr size_t slot_bytes(int slots) {
  return (size_t)slots * sizeof(Fq) * T;
}

// The bases of one MSM call: `copies` tables of affine points (device, 2K Fq each), table j at
// bases + j * copy_stride points holding 2^(j * rows * c) * P_i (see g753_bases_precompute);
// copies == 1 is a plain key.  inf: infinity flags, one per (copy, point), may be null.
struct MsmKey {
  const Fq* bases = nullptr;
  const uint8_t* inf = nullptr;
  size_t copy_stride = 0;
  size_t inf_stride = 0;
  unsigned copies = 1;
  int c = 0;          // window bits the tables were built for (0 = choose per call)
  unsigned rows = 0;  // bucket rows the tables were built for (0 = derive)
  int affine = -1;    // accumulation: -1 = the group's default, 0 = XYZZ running sums, 1 = affine addition tree
  int tree_batch = 0; // output slots per thread of the addition tree (0 = TREE_BATCH)
  int tree_waves = 0; // waves of blocks a level is cut into before its batches grow (0 = 4)
};

// out[g * ceil(m_in / 2) + p] = in[g * m_in + 2p] + in[g * m_in + 2p + 1]: one thread per output point while
// there are many, one warp per point (coop.cuh) once the launch is latency-bound
template <int GID>
static inline void msm_pair_sum(MsmHooks& hooks, cudaStream_t stream, const Fq* in, unsigned groups, unsigned m_in, Fq* out) {
  typedef MsmCfg<GID> Cfg;
  constexpr int CR = Cfg::NC_RED, TR = CR * Cfg::TP;
  typedef typename Cfg::template SC<CR, Cfg::TP> SCR;
  typedef EcS<SCR> ER;
  constexpr size_t SMEM_RED = slot_bytes<ER, CR>(2 * ER::PT + ER::ADD_SCRATCH);
  const size_t items = (size_t)groups * div_up(m_in, 2);
#if !defined(G753_HOST_EMUL)
  if (items <= COOP_MAX_ITEMS) {
    G753_MSM_LAUNCH_SMEM(hooks, k_pair_sum_coop<GID>, div_up(items, COOP_WARPS), 32 * COOP_WARPS,
                         coop_smem_bytes<GID>(COOP_WARPS), stream, in, groups, m_in, out);
    return;
  }
#endif
  G753_MSM_LAUNCH_SMEM(hooks, k_pair_sum<SCR>, div_up(items, CR), TR, SMEM_RED, stream, in, groups, m_in, out);
}

// d_scalars: count x 24 u32 canonical (device); d_out: 3K Fq (device)
template <int GID>
static int msm_run(Scratch& scratch, cudaStream_t stream, const MsmKey& key, const uint32_t* d_scalars,
                   size_t count, Fq* d_out, MsmHooks hooks) {
  typedef MsmCfg<GID> Cfg;
  constexpr int CA = Cfg::NC_ACC, CR = Cfg::NC_RED;        // columns (curve operations) per block
  constexpr int TA = Lay<CA, Cfg::TPA>::THREADS, TR = CR * Cfg::TP;     // threads per block (whole warps)
  typedef typename Cfg::template SC<CA, Cfg::TPA> SCA;
  typedef typename Cfg::template SC<CR, Cfg::TP> SCR;
  typedef EcS<SCA> EA;
  typedef EcS<SCR> ER;
  constexpr size_t PT = 4 * Cfg::K;  // Fq per XYZZ point
  constexpr size_t SMEM_ACC = slot_bytes<EA, CA>(EA::PT + EA::ACC_SCRATCH);
  constexpr size_t SMEM_FIX = slot_bytes<EA, CA>(EA::PT + EA::ADD_SCRATCH);
  constexpr size_t SMEM_RED = slot_bytes<ER, CR>(2 * ER::PT + ER::ADD_SCRATCH);
  static_assert(SMEM_ACC <= 232448 && SMEM_FIX <= 232448 && SMEM_RED <= 232448, "slot footprint exceeds 227 KB");
  if (count == 0) {
    G753_MSM_LAUNCH(hooks, k_write_infinity<SCR>, 1, 1, stream, d_out);
    return launch_check("k_write_infinity");
  }
  if (count > 0x7fffffffull) return fail(G753_ERR_BAD_ARG, "msm: more than 2^31 - 1 scalars");
  const unsigned n = (unsigned)count;
  const MsmPlan pl = msm_plan(msm_cost<GID>(), n, key.copies, key.c, key.rows);
  if (pl.c == 0) return fail(G753_ERR_BAD_ARG, "msm: no window size satisfies the forced plan");
  constexpr size_t SMEM_TREE = slot_bytes<EA, CA>(tree_slots<typename EA::M>());
  static_assert(SMEM_TREE <= 232448, "slot footprint exceeds 227 KB");
  // the tree pays one inversion per thread and round - ~0.4 ms of latency per round whatever the size - so short
  // MSMs keep the running sums; G753_MSM_AFFINE=0 / 1 forces one form
  bool tree = key.affine < 0 ? (uint64_t)pl.W * n >= TREE_MIN_ENTRIES : key.affine > 0;
  MsmWorkspace ws = msm_workspace<GID>(pl, n, tree);
  if (tree && key.affine < 0 && !scratch.can_hold(ws.total)) {   // the tree's level arrays do not fit: running sums
    tree = false;
    ws = msm_workspace<GID>(pl, n, false);
  }
  if ((uint64_t)pl.W * n >= 0xffffffffull || (uint64_t)pl.rows * ws.row_cap >= 0xffffffffull ||
      (uint64_t)(pl.copies - 1) * key.copy_stride + n > 0x7fffffffull)
    return fail(G753_ERR_BAD_ARG, "msm: windows x points exceed the 32-bit index space of the sort");
  if (hooks.form) *hooks.form = tree ? 1u : 0u;
  G753_TRY(scratch.reserve(ws.total));
  Carver cv(scratch.ptr);
  const size_t len = (size_t)pl.B + 1;
  const unsigned R = pl.rows;
  const unsigned NB = (unsigned)(R * len);
  uint32_t* digits = cv.take<uint32_t>((size_t)pl.W * n);
  uint32_t* sorted = cv.take<uint32_t>((size_t)R * ws.row_cap);
  uint32_t* hist = cv.take<uint32_t>(NB);
  uint32_t* offsets = cv.take<uint32_t>(NB);
  uint32_t* cursor = cv.take<uint32_t>(NB);
  uint32_t* item_cnt = cv.take<uint32_t>(NB);
  uint32_t* item_off = cv.take<uint32_t>(NB);
  uint32_t* chunk_sums = cv.take<uint32_t>((size_t)R * ws.n_chunks);
  uint32_t* item_chunk_sums = cv.take<uint32_t>(ws.nb_chunks);
  uint32_t* len_hist = cv.take<uint32_t>(ITEM_LEN * ITEM_REP);
  uint32_t* len_cursor = cv.take<uint32_t>(ITEM_LEN * ITEM_REP);
  uint32_t* item_total = cv.take<uint32_t>(64);
  MsmItem* items = cv.take<MsmItem>(ws.max_items);
  uint32_t* part_bucket = cv.take<uint32_t>(ws.max_items);
  Fq* points = cv.take<Fq>(PT * ((size_t)NB + ws.max_items));
  Fq* lvl_r = cv.take<Fq>(PT * R * ws.level_entries);
  Fq* lvl_y = cv.take<Fq>(PT * R * ws.level_entries);
  Fq* tail_a = cv.take<Fq>(PT * (ws.tail_points + 1));
  Fq* tail_b = cv.take<Fq>(PT * (ws.tail_points + 1));
  Fq* tree_odd = tree ? cv.take<Fq>(PT / 2 * ws.tree_odd) : nullptr;
  Fq* tree_even = tree ? cv.take<Fq>(PT / 2 * ws.tree_even) : nullptr;
  uint32_t* tree_max = cv.take<uint32_t>(64);

  if (hooks.mark) hooks.mark(hooks.user, 0);
  G753_TRY(dev_memset(hist, 0, sizeof(uint32_t) * NB, stream));
  G753_TRY(dev_memset(len_hist, 0, sizeof(uint32_t) * ITEM_LEN * ITEM_REP, stream));
  // empty buckets are never written by the accumulation: all-zero limbs = ZZ == 0 = infinity
  G753_TRY(dev_memset(points, 0, sizeof(Fq) * PT * NB, stream));
  {
    // scalar chunks: the host-buffer entry point uploads the scalars in pieces on a copy stream and
    // makes this stream wait for piece j right before its digits are extracted
    const unsigned chunks = hooks.scalar_chunks > 1 ? hooks.scalar_chunks : 1;
    for (unsigned j = 0; j < chunks; j++) {
      const unsigned lo = (unsigned)((uint64_t)n * j / chunks), hi = (unsigned)((uint64_t)n * (j + 1) / chunks);
      if (hooks.wait_chunk) hooks.wait_chunk(hooks.user, (int)j);
      if (hi > lo)
        G753_MSM_LAUNCH(hooks, k_msm_digits, div_up(hi - lo, 256), 256, stream, d_scalars, key.inf, key.inf_stride, n, lo,
                        hi, pl.c, pl.W, pl.B, R, digits, hist);
    }
  }
  if (hooks.mark) hooks.mark(hooks.user, 1);
  G753_MSM_LAUNCH(hooks, k_scan_chunks, div_up((size_t)R * ws.n_chunks, 128), 128, stream, hist,
                  (unsigned)len, ws.n_chunks, R, chunk_sums);
  G753_MSM_LAUNCH(hooks, k_scan_tops, div_up(R, 64), 64, stream, chunk_sums, ws.n_chunks, R,
                  (uint32_t*)nullptr);
  G753_MSM_LAUNCH(hooks, k_scan_apply, div_up((size_t)R * ws.n_chunks, 128), 128, stream, hist,
                  chunk_sums, (unsigned)len, ws.n_chunks, R, offsets, cursor);
  G753_MSM_LAUNCH(hooks, k_msm_scatter, div_up((size_t)pl.W * n, 256), 256, stream, digits, n, pl.W, pl.B, R,
                  key.copy_stride, ws.row_cap, cursor, sorted);
  if (tree) {
    // addition tree (after the scatter, cursor[t] is the end of bucket t's run)
    G753_TRY(dev_memset(tree_max, 0, sizeof(uint32_t) * 64, stream));   // fullest bucket + one "infinity stored" flag per level
    G753_MSM_LAUNCH(hooks, k_tree_max, (unsigned)(NB < 148u * 1024 ? div_up(NB, 256) : 148u * 4), 256, stream, offsets, cursor,
                    NB, tree_max);
    if (hooks.mark) hooks.mark(hooks.user, 2);
    TreeGeo geo{offsets, cursor, NB, (unsigned)len, (uint32_t)ws.row_cap};
    size_t bound = (size_t)R * ws.row_cap;   // slots of the level below
    unsigned rounds = 0;
    while (((size_t)1 << rounds) < ws.row_cap) rounds++;   // a bucket holds at most row_cap entries
    const size_t top = key.tree_batch > 0 ? (size_t)key.tree_batch : (size_t)TREE_BATCH;
    int resident = 2;   // blocks per SM: columns the GPU runs at once = one wave
#if !defined(G753_HOST_EMUL)
    cudaFuncSetAttribute(k_tree_round<SCA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TREE);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_tree_round<SCA>, TA, SMEM_TREE) != cudaSuccess || resident < 1)
      resident = 2;
#endif
    const size_t wave = (size_t)148 * (size_t)resident * CA;
    for (unsigned level = 1; level <= rounds; level++) {
      bound = tree_level_bound(bound, NB);
      // slots per thread: enough threads for four waves of blocks first (a thread is a serial chain, and levels
      // end with their slowest thread), then longer batches - one inversion (~113 000 instructions against
      // ~9 700 per addition) per thread and round.  The upper levels are mostly holes and finished buckets: there
      // a batch is sized for ~48 expected PAIRS (uniform digits) as long as one wave of threads remains.
      // Measured at 2^22: exact single waves of longer batches on every short level lose 5 ms, 256 slots
      // everywhere 40 ms.
      size_t batch = top;
      if (key.tree_batch <= 0) {   // (G753_TREE_BATCH fixes the batch)
        const size_t four = div_up(bound, (key.tree_waves > 0 ? (size_t)key.tree_waves : 4) * wave), one = div_up(bound, wave);
        size_t pairs = ((size_t)pl.W * n) >> level;
        if (pairs < 1) pairs = 1;
        size_t by_pairs = 48 * bound / pairs;
        if (by_pairs > one) by_pairs = one;
        batch = four > by_pairs ? four : by_pairs;
        batch = batch < 4 ? 4 : batch > top ? top : batch;
        // ... and WHOLE waves: a few blocks more than the GPU holds run as one more wave (measured: 446 blocks on
        // 444 block slots took twice the time; at 2^22 levels 2 and 3 ran 4.01 waves)
        const size_t waves = div_up(bound, batch * wave);
        batch = div_up(bound, waves * wave);
        if (batch < 4) batch = 4;
      }
      G753_MSM_LAUNCH_SMEM(hooks, k_tree_round<SCA>, div_up(div_up(bound, batch), CA), TA, SMEM_TREE, stream, key.bases, sorted,
                           geo, tree_max, level, (unsigned)batch, (level & 1) ? tree_even : tree_odd,
                           (level & 1) ? tree_odd : tree_even, points);
    }
    G753_MSM_LAUNCH_SMEM(hooks, k_tree_finish<SCA>, div_up(NB, CA), TA, SMEM_TREE, stream, key.bases, sorted, geo, points);
  } else {
  // work items (after the scatter, cursor[t] is the end of bucket t's run)
    G753_MSM_LAUNCH(hooks, k_item_count, div_up(NB, 256), 256, stream, offsets, cursor, NB, pl.B, item_cnt,
                    len_hist);
    G753_MSM_LAUNCH(hooks, k_scan_chunks, div_up(ws.nb_chunks, 128), 128, stream, item_cnt, NB, ws.nb_chunks,
                    1u, item_chunk_sums);
    G753_MSM_LAUNCH(hooks, k_scan_tops, 1, 64, stream, item_chunk_sums, ws.nb_chunks, 1u, (uint32_t*)nullptr);
    G753_MSM_LAUNCH(hooks, k_scan_apply, div_up(ws.nb_chunks, 128), 128, stream, item_cnt, item_chunk_sums,
                    NB, ws.nb_chunks, 1u, item_off, (uint32_t*)nullptr);
    G753_MSM_LAUNCH(hooks, k_item_len_scan, 1, 32, stream, len_hist, len_cursor, item_total);
    G753_MSM_LAUNCH(hooks, k_item_emit, div_up(NB, 256), 256, stream, offsets, cursor, item_cnt, item_off, NB,
                    pl.B, ws.row_cap, len_cursor, items, part_bucket);
    if (hooks.mark) hooks.mark(hooks.user, 2);
    G753_MSM_LAUNCH_SMEM(hooks, k_bucket_acc<SCA>, div_up(ws.max_items, CA), TA, SMEM_ACC, stream, key.bases,
                           sorted, items, item_total, points);
    // a bucket holds at most row_cap points = row_cap / ITEM_LEN + 1 partials
    for (size_t stride = 1; stride <= ws.row_cap / ITEM_LEN; stride *= FIX_FAN)
      G753_MSM_LAUNCH_SMEM(hooks, k_bucket_fixup_level<SCA>, div_up(ws.max_items, CA), TA, SMEM_FIX, stream, item_cnt,
                           item_off, part_bucket, item_total, NB, (unsigned)stride, points);
    G753_MSM_LAUNCH_SMEM(hooks, k_bucket_fixup<SCA>, div_up(NB, CA), TA, SMEM_FIX, stream, item_cnt, item_off,
                         NB, points);
  }
  if (hooks.mark) hooks.mark(hooks.user, 3);
  // reduction levels
  const Fq* X = points;
  size_t x_stride = len;
  const Fq* Y = nullptr;
  size_t y_stride = 0;
  unsigned n_in = (unsigned)len, log2f = 0;
  size_t lvl_off = 0;
  const Fq* window_sums = nullptr;
  for (;;) {
    unsigned n_out = div_up(n_in, REDUCE_SEG);
    Fq* Rr = lvl_r + lvl_off * PT;
    Fq* Yo = lvl_y + lvl_off * PT;
    G753_MSM_LAUNCH_SMEM(hooks, k_reduce_level<SCR>, div_up((size_t)R * n_out, CR), TR, SMEM_RED, stream, X,
                         x_stride, Y, y_stride, n_in, log2f, R, n_out, Rr, Yo);
    lvl_off += (size_t)R * n_out;
    X = Rr;
    Y = Yo;
    x_stride = y_stride = n_out;
    n_in = n_out;
    log2f += REDUCE_SEG_LOG;
    if (n_out == 1) {
      window_sums = Yo;
      break;
    }
    if (n_out <= REDUCE_TAIL_MAX) {
      // tail: bit planes + pairwise trees (see k_tail_planes)
      unsigned nbits = 0;
      while ((1u << nbits) < n_out) nbits++;
      const unsigned P = nbits + 1;
      unsigned m = div_up(n_out, 2);
      G753_MSM_LAUNCH_SMEM(hooks, k_tail_planes<SCR>, div_up((size_t)R * P * m, CR), TR, SMEM_RED, stream, X, Y, n_out,
                           nbits, R, tail_a);
      Fq *cur = tail_a, *oth = tail_b;
      while (m > 1) {
        msm_pair_sum<GID>(hooks, stream, cur, R * P, m, oth);
        m = div_up(m, 2);
        Fq* t2 = cur;
        cur = oth;
        oth = t2;
      }
#if defined(G753_HOST_EMUL)
      G753_MSM_LAUNCH_SMEM(hooks, k_tail_scale<SCR>, div_up((size_t)R * P, CR), TR, SMEM_RED, stream, cur, nbits, R, log2f,
                           oth);
#else
      G753_MSM_LAUNCH_SMEM(hooks, k_tail_scale_coop<GID>, div_up((size_t)R * P, COOP_WARPS), 32 * COOP_WARPS,
                           coop_smem_bytes<GID>(COOP_WARPS), stream, cur, nbits, R, log2f, oth);
#endif
      {
        Fq* t2 = cur;
        cur = oth;
        oth = t2;
      }
      m = P;
      while (m > 1) {
        msm_pair_sum<GID>(hooks, stream, cur, R, m, oth);
        m = div_up(m, 2);
        Fq* t2 = cur;
        cur = oth;
        oth = t2;
      }
      window_sums = cur;
      break;
    }
  }
  if (hooks.mark) hooks.mark(hooks.user, 4);
#if defined(G753_HOST_EMUL)
  G753_MSM_LAUNCH_SMEM(hooks, k_window_combine<SCR>, 1, TR, SMEM_RED, stream, window_sums, 1u, R, pl.c,
                       d_out);
#else
  G753_MSM_LAUNCH_SMEM(hooks, k_window_combine_coop<GID>, 1, 32, coop_smem_bytes<GID>(1), stream, window_sums, 1u, R, pl.c,
                       d_out);
#endif
  if (hooks.mark) hooks.mark(hooks.user, 5);
  return launch_check("msm_run");
}

}  // namespace g753
