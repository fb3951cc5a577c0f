// Launch / memory plumbing shared by the MSM and NTT pipelines.
//
// Two builds of the same kernel source exist:
//   * the product: nvcc, sm_100a, real CUDA launches (this is what libg753.so contains);
//   * G753_HOST_EMUL: a TEST-ONLY g++ build (tests/host_emul) in which a "launch" is a
//     sequential loop over (block, thread) on the host and the PTX carry primitives are
//     emulated.  It exists so the indexing / orchestration logic of barrier-free kernels can
//     be checked in the CPU-only test tier.  It is never linked into libg753.so and nothing
//     in the product loads it - the product has no CPU path.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/g753.h"

#if defined(G753_HOST_EMUL)
// ---------------------------------------------------------------- host emulation shims
struct EmulDim {
  unsigned x, y, z;
};
static thread_local EmulDim blockIdx, threadIdx, blockDim, gridDim;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
typedef void* cudaStream_t;
struct alignas(16) uint4 {
  unsigned x, y, z, w;
};
template <class T>
static inline T atomicAdd(T* p, T v) {
  T old = *p;
  *p = old + v;
  return old;
}
template <class T>
static inline T atomicOr(T* p, T v) {
  T old = *p;
  *p = old | v;
  return old;
}
template <class T>
static inline T atomicMax(T* p, T v) {
  T old = *p;
  if (v > old) *p = v;
  return old;
}
static inline unsigned __brev(unsigned x) {
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
  return r;
}
#define G753_LAUNCH(kernel, grid, block, stream, ...)                       \
  do {                                                                      \
    gridDim = EmulDim{(unsigned)(grid), 1, 1};                              \
    blockDim = EmulDim{(unsigned)(block), 1, 1};                            \
    for (unsigned _b = 0; _b < (unsigned)(grid); _b++)                      \
      for (unsigned _t = 0; _t < (unsigned)(block); _t++) {                 \
        blockIdx = EmulDim{_b, 0, 0};                                       \
        threadIdx = EmulDim{_t, 0, 0};                                      \
        kernel(__VA_ARGS__);                                                \
      }                                                                     \
  } while (0)
#define G753_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) G753_LAUNCH(kernel, grid, block, stream, __VA_ARGS__)
namespace g753 {
static thread_local char g_last_error[512] = "";
static inline int fail(int code, const char* msg) {
  snprintf(g_last_error, sizeof(g_last_error), "%s", msg);
  return code;
}
static inline int dev_alloc(void** p, size_t bytes) {
  *p = malloc(bytes ? bytes : 1);
  return *p ? G753_OK : G753_ERR_OOM;
}
static inline void dev_free(void* p) { free(p); }
static inline size_t dev_mem_free() { return (size_t)1 << 40; }
static inline int dev_memset(void* p, int v, size_t bytes, cudaStream_t) {
  memset(p, v, bytes);
  return G753_OK;
}
static inline int h2d(void* d, const void* h, size_t bytes, cudaStream_t) {
  memcpy(d, h, bytes);
  return G753_OK;
}
static inline int d2h(void* h, const void* d, size_t bytes, cudaStream_t) {
  memcpy(h, d, bytes);
  return G753_OK;
}
static inline int d2d(void* d, const void* s, size_t bytes, cudaStream_t) {
  memmove(d, s, bytes);
  return G753_OK;
}
static inline int stream_sync(cudaStream_t) { return G753_OK; }
static inline int launch_check(const char*) { return G753_OK; }
}  // namespace g753
#else
// ---------------------------------------------------------------- real device
#include <cuda_runtime.h>
#define G753_LAUNCH(kernel, grid, block, stream, ...) \
  kernel<<<(unsigned)(grid), (unsigned)(block), 0, (stream)>>>(__VA_ARGS__)
// launch with `smem` bytes of dynamic shared memory (opt-in above 48 KB).  The driver's own carveout
// choice is kept: asking for cudaSharedmemCarveoutMaxShared changed nothing (226.1 vs 224.3 ms at 2^22)
#define G753_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...)                                        \
  do {                                                                                                  \
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem));             \
    kernel<<<(unsigned)(grid), (unsigned)(block), (size_t)(smem), (stream)>>>(__VA_ARGS__);             \
  } while (0)
namespace g753 {
extern thread_local char g_last_error[512];
static inline int fail(int code, const char* msg) {
  snprintf(g_last_error, sizeof(g_last_error), "%s", msg);
  return code;
}
static inline int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
  return e == cudaErrorMemoryAllocation ? G753_ERR_OOM : G753_ERR_CUDA;
}
static inline int dev_alloc(void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes ? bytes : 1);
  return e == cudaSuccess ? G753_OK : cuda_fail(e, "cudaMalloc");
}
static inline void dev_free(void* p) {
  if (p) cudaFree(p);
}
static inline size_t dev_mem_free() {
  size_t free_b = 0, total_b = 0;
  return cudaMemGetInfo(&free_b, &total_b) == cudaSuccess ? free_b : 0;
}
static inline int dev_memset(void* p, int v, size_t bytes, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(p, v, bytes, s);
  return e == cudaSuccess ? G753_OK : cuda_fail(e, "cudaMemsetAsync");
}
static inline int h2d(void* d, const void* h, size_t bytes, cudaStream_t s) {
  cudaError_t e = cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s);
  return e == cudaSuccess ? G753_OK : cuda_fail(e, "cudaMemcpyAsync H2D");
}
static inline int d2h(void* h, const void* d, size_t bytes, cudaStream_t s) {
  cudaError_t e = cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s);
  return e == cudaSuccess ? G753_OK : cuda_fail(e, "cudaMemcpyAsync D2H");
}
static inline int d2d(void* d, const void* s_, size_t bytes, cudaStream_t s) {
  cudaError_t e = cudaMemcpyAsync(d, s_, bytes, cudaMemcpyDeviceToDevice, s);
  return e == cudaSuccess ? G753_OK : cuda_fail(e, "cudaMemcpyAsync D2D");
}
static inline int stream_sync(cudaStream_t s) {
  cudaError_t e = cudaStreamSynchronize(s);
  return e == cudaSuccess ? G753_OK : cuda_fail(e, "cudaStreamSynchronize");
}
static inline int launch_check(const char* what) {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? G753_OK : cuda_fail(e, what);
}
}  // namespace g753
#endif

#define G753_TRY(expr)            \
  do {                            \
    int _rc = (expr);             \
    if (_rc != G753_OK) return _rc; \
  } while (0)

namespace g753 {

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// A grow-only device scratch buffer owned by the context (no per-call cudaMalloc once warm).
struct Scratch {
  void* ptr = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return G753_OK;
    dev_free(ptr);
    ptr = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8;
    int rc = dev_alloc(&ptr, want);
    if (rc != G753_OK) return rc;
    cap = want;
    return G753_OK;
  }
  // would reserve(bytes) succeed?  (growing frees the current block first)
  bool can_hold(size_t bytes) const {
    if (bytes <= cap) return true;
    return dev_mem_free() + cap > bytes + bytes / 8 + ((size_t)256 << 20);
  }
  void release() {
    dev_free(ptr);
    ptr = nullptr;
    cap = 0;
  }
};

// carve typed, 256-byte aligned sub-buffers out of one Scratch
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base((char*)b) {}
  template <class T>
  T* take(size_t count) {
    off = (off + 255) & ~(size_t)255;
    T* p = (T*)(base + off);
    off += count * sizeof(T);
    return p;
  }
  static size_t pad(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
};

}  // namespace g753
