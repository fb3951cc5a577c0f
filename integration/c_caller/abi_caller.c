/* A compiled (C99) caller of libg753.so: the C ABI exercised without Python or ctypes.
 *
 *   gcc -std=c99 -O2 -I include integration/c_caller/abi_caller.c -L ginger-lib_b200 -lg753 \
 *       -Wl,-rpath,$PWD/ginger-lib_b200 -o abi_caller
 *   ./abi_caller gen_g1.bin        (gen_g1.bin: the 24 Montgomery limbs x || y of the MNT4-753 G1 generator)
 *
 * Exit status: 0 all checks passed on a GPU; 3 no CUDA device (and the library refused to run, as it must:
 * there is no CPU fallback); anything else is a failure.
 *
 * Checks, all through the header's entry points only:
 *   MSM   sum_i (i + 1) * (a_i G) over a generated 4096-point key == (sum_i (i + 1) a_i) * G
 *         (g753_bases_generate / g753_msm_host-free resident path / g753_point_op / g753_batch_normalize);
 *   NTT   ifft(fft(x)) == x and fft(delta_0) == (1, 1, ..., 1) on 2^12 elements of mnt4753::Fr
 *         (g753_field_op for the Montgomery one, g753_ntt in place on host memory).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "g753.h"

#define CHECK(call)                                                                   \
  do {                                                                                \
    int rc_ = (call);                                                                 \
    if (rc_ != G753_OK) {                                                             \
      fprintf(stderr, "%s failed: %d (%s)\n", #call, rc_, g753_last_error());         \
      return 10;                                                                      \
    }                                                                                 \
  } while (0)

static uint64_t splitmix64_at(uint64_t seed, uint64_t i) { /* g753_bases_generate's a_i, include/g753.h */
  uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (z ^ (z >> 31)) | 1ull;
}

int main(int argc, char** argv) {
  int count = 0;
  printf("%s, sources %s\n", g753_version(), g753_source_hash());
  if (g753_device_count(&count) != G753_OK || count == 0) {
    g753_ctx* none = NULL;
    int rc = g753_ctx_create(0, &none);
    printf("no CUDA device: g753_ctx_create -> %d (%s)\n", rc, g753_last_error());
    return (rc == G753_ERR_NO_DEVICE && none == NULL) ? 3 : 11;
  }
  if (argc < 2) {
    fprintf(stderr, "usage: %s gen_g1.bin\n", argv[0]);
    return 12;
  }
  uint64_t gen[2 * G753_LIMBS];
  FILE* fh = fopen(argv[1], "rb");
  if (!fh || fread(gen, sizeof(gen), 1, fh) != 1) {
    fprintf(stderr, "cannot read the generator from %s\n", argv[1]);
    return 13;
  }
  fclose(fh);

  g753_ctx* ctx = NULL;
  CHECK(g753_ctx_create(0, &ctx));

  /* ---- MSM ---------------------------------------------------------------------------------- */
  enum { N = 4096 };
  const uint64_t seed = 0xC0DEull;
  g753_bases* key = NULL;
  CHECK(g753_bases_generate(ctx, G753_MNT4_G1, gen, seed, N, &key));
  if (g753_bases_len(key) != N) return 14;
  uint64_t* scalars = (uint64_t*)calloc((size_t)N * G753_LIMBS, sizeof(uint64_t));
  unsigned __int128 dot = 0; /* sum (i + 1) a_i < 2^12 * 2^12 * 2^64: no reduction mod r needed */
  for (uint64_t i = 0; i < N; i++) {
    scalars[i * G753_LIMBS] = i + 1;
    dot += (unsigned __int128)(i + 1) * splitmix64_at(seed, i);
  }
  uint64_t got[3 * G753_LIMBS], want[3 * G753_LIMBS], k[G753_LIMBS] = {0};
  k[0] = (uint64_t)dot;
  k[1] = (uint64_t)(dot >> 64);
  CHECK(g753_msm(ctx, key, 0, N, scalars, got));
  CHECK(g753_point_op(ctx, G753_MNT4_G1, 2, gen, k, want));
  uint64_t both[2 * 3 * G753_LIMBS], xy[2 * 2 * G753_LIMBS];
  uint8_t inf[2];
  memcpy(both, got, sizeof(got));
  memcpy(both + 3 * G753_LIMBS, want, sizeof(want));
  CHECK(g753_batch_normalize(ctx, G753_MNT4_G1, both, 2, xy, inf));
  if (inf[0] || inf[1] || memcmp(xy, xy + 2 * G753_LIMBS, 2 * G753_LIMBS * sizeof(uint64_t)) != 0) {
    fprintf(stderr, "MSM result differs from (sum s_i a_i) * G\n");
    return 20;
  }
  /* the same through precomputed key copies and on a slice view */
  CHECK(g753_bases_precompute(ctx, key, 4));
  CHECK(g753_msm(ctx, key, 0, N, scalars, got));
  memcpy(both, got, sizeof(got));
  CHECK(g753_batch_normalize(ctx, G753_MNT4_G1, both, 2, xy, inf));
  if (memcmp(xy, xy + 2 * G753_LIMBS, 2 * G753_LIMBS * sizeof(uint64_t)) != 0) return 21;
  CHECK(g753_msm(ctx, key, 0, 0, scalars, got)); /* empty input: (0 : 1 : 0) */
  for (int i = 0; i < G753_LIMBS; i++)
    if (got[i] != 0 || got[2 * G753_LIMBS + i] != 0) return 22;
  CHECK(g753_bases_free(ctx, key));
  printf("MSM ok (%d points, %llu kernel launches so far)\n", (int)N, (unsigned long long)g753_launch_count(ctx));

  /* ---- NTT ---------------------------------------------------------------------------------- */
  const unsigned LOG_N = 12;
  const size_t n = (size_t)1 << LOG_N;
  if (g753_domain_check(G753_FIELD_MNT6_FR, 15) != G753_ERR_DOMAIN) return 30; /* EvaluationDomain::new -> None */
  uint64_t one_raw[G753_LIMBS] = {1}, one[G753_LIMBS];
  CHECK(g753_field_op(ctx, G753_FIELD_MNT4_FR, G753_OP_TO_MONT, one_raw, NULL, one, 1));
  uint64_t* x = (uint64_t*)malloc(n * G753_LIMBS * sizeof(uint64_t));
  uint64_t* y = (uint64_t*)malloc(n * G753_LIMBS * sizeof(uint64_t));
  uint64_t s = 0x1234567ull;
  for (size_t i = 0; i < n * G753_LIMBS; i++) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    x[i] = (i % G753_LIMBS == G753_LIMBS - 1) ? (s >> 50) : s; /* < 2^752 < p: valid Montgomery forms */
  }
  memcpy(y, x, n * G753_LIMBS * sizeof(uint64_t));
  CHECK(g753_ntt(ctx, G753_FIELD_MNT4_FR, y, LOG_N, G753_FFT));
  if (memcmp(x, y, n * G753_LIMBS * sizeof(uint64_t)) == 0) return 31;
  CHECK(g753_ntt(ctx, G753_FIELD_MNT4_FR, y, LOG_N, G753_IFFT));
  if (memcmp(x, y, n * G753_LIMBS * sizeof(uint64_t)) != 0) {
    fprintf(stderr, "ifft(fft(x)) != x\n");
    return 32;
  }
  memset(y, 0, n * G753_LIMBS * sizeof(uint64_t));
  memcpy(y, one, sizeof(one));
  CHECK(g753_ntt(ctx, G753_FIELD_MNT4_FR, y, LOG_N, G753_COSET_IFFT));
  CHECK(g753_ntt(ctx, G753_FIELD_MNT4_FR, y, LOG_N, G753_COSET_FFT));
  CHECK(g753_ntt(ctx, G753_FIELD_MNT4_FR, y, LOG_N, G753_FFT));
  for (size_t i = 0; i < n; i++)
    if (memcmp(y + i * G753_LIMBS, one, sizeof(one)) != 0) {
      fprintf(stderr, "fft(delta_0)[%zu] != 1\n", i);
      return 33;
    }
  printf("NTT ok (2^%u elements)\n", LOG_N);
  free(x);
  free(y);
  free(scalars);
  CHECK(g753_ctx_destroy(ctx));
  printf("abi_caller: all checks passed\n");
  return 0;
}
