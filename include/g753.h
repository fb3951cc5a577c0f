/* g753 - C ABI of the B200-native Groth16 prover hot path for MNT4-753 / MNT6-753.
 *
 * This is the drop-in boundary: exactly the entry points a thin FFI layer
 * under ginger-lib's `algebra::msm` / `algebra::fft` would bind (INTEGRATION.md shows the
 * Rust `-sys` crate).  Plain pointers and sizes only; no C++ or torch types.
 *
 * Data formats (identical to the reference's in-memory forms, so buffers cross unconverted):
 *   - a base-field element Fq is 12 little-endian uint64 limbs = BigInteger768.0
 *     (algebra/src/biginteger/mod.rs:20), holding the MONTGOMERY representation
 *     (value * 2^768 mod p), as Fp768 stores it (algebra/src/fields/models/fp_768.rs:24-30);
 *   - Fq2 = c0,c1 and Fq3 = c0,c1,c2, consecutive (fields/models/fp2.rs:51-57, fp3.rs:60-67);
 *   - an affine point is x then y (2*k*12 limbs, k = 1/2/3) plus one byte in a separate
 *     `infinity` array (GroupAffine{x,y,infinity}, curves/models/short_weierstrass_projective.rs:26-32);
 *   - an MSM scalar is the CANONICAL integer < r, 12 limbs, i.e. `Fr::into_repr()`
 *     (proof-systems/src/groth16/prover.rs:241-267);
 *   - an MSM result is a homogeneous projective point X,Y,Z (3*k*12 limbs, Montgomery form,
 *     every coordinate fully reduced), any representative - what
 *     VariableBaseMSM::multi_scalar_mul returns (algebra/src/msm/variable_base.rs:85-90);
 *     the point at infinity is (0, 1, 0) as GroupProjective::zero().
 *
 * Every function returns 0 on success or a G753_ERR_* code; nothing aborts or unwinds across
 * the boundary.  g753_last_error() gives a thread-local description of the last failure.
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef G753_H
#define G753_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G753_OK 0
#define G753_ERR_BAD_ARG 1      /* null pointer, unknown id, size out of range            */
#define G753_ERR_CUDA 2         /* a CUDA runtime call or kernel failed                   */
#define G753_ERR_OOM 3          /* device memory exhausted                                */
#define G753_ERR_DOMAIN 4       /* log_n >= two-adicity: EvaluationDomain::new -> None    */
#define G753_ERR_NO_DEVICE 5    /* no usable CUDA device                                  */

#define G753_LIMBS 12           /* uint64 limbs per Fq element                            */

/* groups: AffineCurve instantiations on the prover path
 *   curves/mnt4753/mod.rs:105-109, curves/mnt6753/mod.rs:106-110 */
#define G753_MNT4_G1 0          /* over Fq   (k = 1) */
#define G753_MNT4_G2 1          /* over Fq2  (k = 2) */
#define G753_MNT6_G1 2          /* over Fq   (k = 1) */
#define G753_MNT6_G2 3          /* over Fq3  (k = 3) */

/* prime fields.  mnt4753::Fq == mnt6753::Fr and mnt6753::Fq == mnt4753::Fr
 *   (fields/mnt4753/fr.rs:1, fields/mnt6753/fr.rs:1) */
#define G753_FIELD_MNT4_FQ 0
#define G753_FIELD_MNT6_FQ 1
#define G753_FIELD_MNT6_FR 0    /* two-adicity 15: domains up to 2^14 */
#define G753_FIELD_MNT4_FR 1    /* two-adicity 30: domains up to 2^29 */

/* EvaluationDomain transforms (algebra/src/fft/domain.rs:113-179) */
#define G753_FFT 0              /* fft_in_place        :120-123 */
#define G753_IFFT 1             /* ifft_in_place       :134-138 */
#define G753_COSET_FFT 2        /* coset_fft_in_place  :163-166 */
#define G753_COSET_IFFT 3       /* coset_ifft_in_place :176-179 */

/* element-wise field operations (test / witness_map helpers) */
#define G753_OP_MUL 0           /* Fp768::mul_assign  fp_768.rs:1009-1185 */
#define G753_OP_ADD 1           /* add_assign         fp_768.rs:929-937   */
#define G753_OP_SUB 2           /* sub_assign         fp_768.rs:939-949   */
#define G753_OP_SQR 3           /* square_in_place    fp_768.rs:339-548   */
#define G753_OP_INV 5           /* inverse            fp_768.rs:551-605   */
#define G753_OP_TO_MONT 6       /* from_repr          fp_768.rs:627-635   */
#define G753_OP_FROM_MONT 7     /* into_repr          fp_768.rs:637-667   */

typedef struct g753_ctx g753_ctx;       /* one CUDA device + stream + scratch + tables */
typedef struct g753_bases g753_bases;   /* a device-resident slice of proving-key bases */
typedef struct g753_ntt_shard g753_ntt_shard; /* one rank's plan of a transform sharded over several GPUs */

/* ---- context ------------------------------------------------------------------------- */
int g753_device_count(int* count);
int g753_ctx_create(int device, g753_ctx** out);
int g753_ctx_destroy(g753_ctx* ctx);
/* run every later call of this context on the caller's CUDA stream (a cudaStream_t, e.g. the
 * one torch.distributed's NCCL collectives are ordered on); the context stops owning a stream */
int g753_ctx_set_stream(g753_ctx* ctx, void* cuda_stream);
const char* g753_last_error(void);
const char* g753_version(void);
/* hash of the sources (every file under csrc, this header, compiler flags) the loaded library was built from; the
 * build recipe (ginger-lib_b200/build.py) rebuilds when it differs from the tree's and bench.py prints
 * both, so a stale prebuilt binary cannot be measured unnoticed */
const char* g753_source_hash(void);

/* ---- MSM: VariableBaseMSM::multi_scalar_mul (algebra/src/msm/variable_base.rs:85-90) -- */
/* Upload n affine bases once per proving key (Parameters::{a,b_g1,b_g2,h,l}_query,
 * proof-systems/src/groth16/mod.rs:313-371); they stay resident in HBM. */
int g753_bases_upload(g753_ctx* ctx, int group, const uint64_t* coords, const uint8_t* infinity,
                      size_t n, g753_bases** out);
/* same from the reference's WIRE format (Parameters::read / GroupAffine::read, proof-systems/src/
 * groth16/mod.rs:211-239, curves/models/short_weierstrass_projective.rs:185-202): n records of
 * x || y || infinity, every coordinate element 96 bytes canonical little-endian (fp_768.rs:784-789),
 * the flag one byte; the conversion to Montgomery form runs on the device.  No curve check, as
 * `read_affine_vec(len, false, ..)` does none. */
int g753_bases_upload_wire(g753_ctx* ctx, int group, const uint8_t* wire, size_t n, g753_bases** out);
int g753_bases_free(g753_ctx* ctx, g753_bases* bases);
/* synthetic key for benchmarks / full-size parity checks: bases[i] = a_i * G on the
 * device, a_i = splitmix64(seed + (i+1) * 0x9E3779B97F4A7C15) | 1 (64 bits), G = gen_xy (affine,
 * Montgomery limbs, e.g. AffineCurve::prime_subgroup_generator(), curves/mnt4753/g1.rs:79-109),
 * normalised to affine.  sum_i s_i bases[i] = (sum_i s_i a_i mod r) * G at any size. */
int g753_bases_generate(g753_ctx* ctx, int group, const uint64_t* gen_xy, uint64_t seed, size_t n,
                        g753_bases** out);
/* Optional, once per resident key: build `copies` (0 = as many as fit a 6 GiB per-key budget, at most
 * 64) tables 2^(j*shift) * P_i next to
 * the key so that an MSM over (most of) it needs W/copies bucket rows instead of W: the bucket
 * reduction and the serial window fold of variable_base.rs:60-82 shrink accordingly.  Costs
 * copies x the key's memory and ~(copies-1) * 760 doublings per point, once; results of later
 * g753_msm* calls are unchanged (same group element).  Calls over short slices (count < n/4)
 * keep using the plain pipeline. */
int g753_bases_precompute(g753_ctx* ctx, g753_bases* bases, unsigned copies);
/* overwrite `count` bases of a resident key (not one with precomputed copies) starting at `first`:
 * lets a caller keep one small pre-allocated key for per-proof points (prover.rs:322-329:
 * s*g_a + r*g1_b - rs*delta_g1 is an MSM over fresh bases) without allocating in the hot path */
int g753_bases_update(g753_ctx* ctx, g753_bases* bases, size_t first, size_t count, const uint64_t* coords,
                      const uint8_t* infinity);
/* copy `count` resident bases starting at `first` back to the host (2*k*12 limbs each) */
int g753_bases_download(g753_ctx* ctx, const g753_bases* bases, size_t first, size_t count,
                        uint64_t* coords);
size_t g753_bases_len(const g753_bases* bases);

/* sum_{i<count} scalars[i] * bases[first+i]  ->  out_xyz (host, 3*k*12 limbs).
 * `first`/`count` give the sub-slice views the prover takes (groth16/mod.rs:318-350); the
 * reference's zip-truncation (variable_base.rs:36) is `count = min(len(bases)-first, len(scalars))`,
 * applied by the caller-side shim.  count == 0 returns the point at infinity. */
int g753_msm(g753_ctx* ctx, const g753_bases* bases, size_t first, size_t count,
             const uint64_t* scalars, uint64_t* out_xyz);
/* same, scalars already in device memory (count*12 limbs) and result left in device memory */
int g753_msm_dev(g753_ctx* ctx, const g753_bases* bases, size_t first, size_t count,
                 const void* d_scalars, void* d_out_xyz);
/* one-shot form with host bases: exactly multi_scalar_mul(&bases[..n_bases], &scalars[..n_scalars]) */
int g753_msm_host(g753_ctx* ctx, int group, const uint64_t* coords, const uint8_t* infinity,
                  size_t n_bases, const uint64_t* scalars, size_t n_scalars, uint64_t* out_xyz);
/* sum of `count` projective points (device memory, 3*k*12 limbs each) -> one projective point;
 * the fold that follows the multi-GPU gather of per-shard partial sums */
int g753_points_sum_dev(g753_ctx* ctx, int group, const void* d_points_xyz, size_t count,
                        void* d_out_xyz);
/* GroupProjective::batch_normalization / into_affine (curves/models/short_weierstrass_projective.rs:
 * 402-442, 663-678): `count` homogeneous points (host, 3*k*12 limbs each) -> affine x,y (2*k*12
 * limbs each, fully reduced) + infinity flags; Z == 0 gives GroupAffine::zero() = (0, 1, true).
 * This is the normalisation that defines bit-exactness of a proof (prover.rs:340-345). */
int g753_batch_normalize(g753_ctx* ctx, int group, const uint64_t* xyz, size_t count, uint64_t* xy,
                         uint8_t* infinity);
/* FixedBaseMSM::multi_scalar_mul followed by batch_normalization / into_affine (algebra/src/msm/
 * fixed_base.rs:66-79; proof-systems/src/groth16/generator.rs:225-319 builds every query of a key
 * this way): out[i] = scalars[i] * base, affine (2*k*12 limbs, fully reduced) + infinity flag
 * ((0, 1, true) for a zero scalar).  base_xy: one affine point; scalars: n x 12 canonical limbs.
 * The window table is internal (the reference's window size, fixed_base.rs:7-13, does not change
 * the points). */
int g753_fixed_base_msm(g753_ctx* ctx, int group, const uint64_t* base_xy, const uint64_t* scalars, size_t n,
                        uint64_t* out_xy, uint8_t* out_infinity);
/* limbs per coordinate element (12 * k) of a group */
int g753_group_coord_limbs(int group);

/* ---- NTT: EvaluationDomain (algebra/src/fft/domain.rs:65-179) -------------------------- */
/* 0 if a radix-2 domain of size 2^log_n exists for the field (domain.rs:70-72), else
 * G753_ERR_DOMAIN - the `None` of EvaluationDomain::new */
int g753_domain_check(int field, unsigned log_n);
/* the public fields of EvaluationDomain (domain.rs:24-39) and the coset constant, 12 Montgomery limbs:
 * which = 0 size_inv, 1 group_gen, 2 group_gen_inv, 3 generator_inv (17^-1),
 * 4 (17^n - 1)^-1, the factor of divide_by_vanishing_poly_on_coset_in_place (domain.rs:245-256) */
int g753_domain_constant(g753_ctx* ctx, int field, unsigned log_n, int which, uint64_t* out);
/* in place on host memory: data = n = 2^log_n elements, already resized by the caller
 * (domain.rs:121 zero-pads or truncates).  mode is one of G753_FFT..G753_COSET_IFFT */
int g753_ntt(g753_ctx* ctx, int field, uint64_t* data, unsigned log_n, int mode);
/* same on a device pointer, for chaining the 7 transforms of R1CStoQAP::witness_map
 * (proof-systems/src/groth16/r1cs_to_qap.rs:121-161) without host round trips */
int g753_ntt_dev(g753_ctx* ctx, int field, void* d_data, unsigned log_n, int mode);
/* element-wise helpers on device vectors for witness_map: a[i] = a[i] op b[i]
 * (mul_polynomials_in_evaluation_domain domain.rs:289-302; ab - c r1cs_to_qap.rs:156-158) and
 * a[i] *= k (divide_by_vanishing_poly_on_coset_in_place domain.rs:245-256) */
int g753_vec_op_dev(g753_ctx* ctx, int field, int op, void* d_a, const void* d_b, size_t n);
int g753_vec_scale_dev(g753_ctx* ctx, int field, void* d_a, const uint64_t* k_mont, size_t n);

/* ---- mixed-radix domains (BASELINE config 4; no counterpart in the reference snapshot) ----------
 * n = 2^a * m, m odd <= 1024, n | p - 1, a <= two-adicity: omega = GENERATOR^((p-1)/n),
 * out[k] = sum_i in[i] omega^(i k) in natural order; the four modes scale as EvaluationDomain's do
 * (1/n for the inverse, coset generator 17).  G753_ERR_DOMAIN when no such domain exists. */
int g753_domain_check_mixed(int field, uint64_t n);
int g753_ntt_mixed(g753_ctx* ctx, int field, uint64_t* data, uint64_t n, int mode);
int g753_ntt_mixed_dev(g753_ctx* ctx, int field, void* d_data, uint64_t n, int mode);

/* ---- NTT sharded over `world` GPUs (one process per GPU), four-step ------------
 * n = n1 * n2 (n1 = 2^ceil(log_n/2) >= n2); rank g owns cols = n2/world columns of the n1 x n2 view:
 *   input  shard  local[i2l][i1] = x[i1*n2 + g*cols + i2l]            (cols x n1 elements)
 *   output shard  local[k1l][k2] = X[(g*rows + k1l) + n1*k2]          (rows x n2 elements, rows = n1/world)
 * One transform = step1 (local column transforms + twiddles, packed into d_send as `world` blocks of
 * rows*cols elements, block h for rank h), an all-to-all of those blocks done by the caller
 * (NCCL over NVLink: the transpose IS the exchange), step2 (unpack d_recv + local row transforms
 * into d_data).  When n1 == n2 an output shard is directly the input shard of the next transform.
 * Results equal g753_ntt's on the gathered vector, for all four modes. */
int g753_ntt_shard_create(g753_ctx* ctx, int field, unsigned log_n, unsigned world, unsigned rank,
                          g753_ntt_shard** out);
int g753_ntt_shard_destroy(g753_ctx* ctx, g753_ntt_shard* plan);
int g753_ntt_shard_shape(const g753_ntt_shard* plan, size_t* n1, size_t* n2, size_t* cols, size_t* rows);
int g753_ntt_shard_step1(g753_ctx* ctx, const g753_ntt_shard* plan, void* d_data, void* d_send, int mode);
int g753_ntt_shard_step2(g753_ctx* ctx, const g753_ntt_shard* plan, const void* d_recv, void* d_data, int mode);
/* Fused compute + exchange over peer memory (NVLink / NVSwitch P2P): peer_z[h] (host array of `world`
 * device pointers) is rank h's row buffer of rows*n2 elements mapped into this process (CUDA IPC /
 * symmetric memory).  step1_fused runs the column transforms and its LAST butterfly pass stores every
 * output, twiddled, straight into its final place Z[k1l][rank*cols + i2l] on the destination GPU: no
 * pack kernel, no collective call, no unpack.  After a cross-rank barrier step2_local transforms the
 * rows of this rank's Z in place (the output shard).  2..8 ranks. */
int g753_ntt_shard_step1_fused(g753_ctx* ctx, const g753_ntt_shard* plan, void* d_data, void* const* peer_z,
                               int mode);
int g753_ntt_shard_step2_local(g753_ctx* ctx, const g753_ntt_shard* plan, void* d_z, int mode);

/* R1CStoQAP::witness_map from the evaluated constraints onwards (proof-systems/src/groth16/
 * r1cs_to_qap.rs:121-166): a, b, c = the n = 2^log_n evaluations <A_i,z>, <B_i,z>, <C_i,z> padded as
 * :111-119 does (Montgomery form); d123 = d1, d2, d3 (3 x 12 limbs, Montgomery form); h receives
 * n + 1 coefficients (Montgomery form).  Seven transforms and the element-wise steps run chained
 * on the device.  The _dev form takes device pointers (a, b, c are clobbered; h holds n + 1). */
int g753_witness_map(g753_ctx* ctx, int field, const uint64_t* a, const uint64_t* b, const uint64_t* c,
                     unsigned log_n, const uint64_t* d123_mont, uint64_t* h);
int g753_witness_map_dev(g753_ctx* ctx, int field, void* d_a, void* d_b, void* d_c, unsigned log_n,
                         const uint64_t* d123_mont, void* d_h);
/* the second half alone, for provers that run the three independent chains ifft -> coset_fft of a, b, c
 * (g753_ntt_dev) on different GPUs: d_a, d_b, d_c hold the COSET EVALUATIONS; d_h receives
 * coset_ifft((a b - c) / Z) with the d1, d2, d3 terms (r1cs_to_qap.rs:137-166); d_a is clobbered.
 * Both _dev forms only queue work on the context's stream. */
int g753_witness_map_tail_dev(g753_ctx* ctx, int field, void* d_a, const void* d_b, const void* d_c, unsigned log_n,
                              const uint64_t* d123_mont, void* d_h);

/* ---- device memory / stream plumbing (so callers can chain without a CUDA toolchain) --- */
int g753_dev_alloc(g753_ctx* ctx, size_t bytes, void** d_ptr);
int g753_dev_free(g753_ctx* ctx, void* d_ptr);
int g753_h2d(g753_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
int g753_d2h(g753_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);
int g753_d2d(g753_ctx* ctx, void* d_dst, const void* d_src, size_t bytes);  /* ordered on the context's stream */
int g753_sync(g753_ctx* ctx);
/* make every later call on `ctx` wait (on the device, not the host) for all work queued so far on
 * `other` - two contexts of one device overlap independent chains of calls, e.g. the latency-bound
 * small MSMs of a proof with its throughput-bound large ones */
int g753_ctx_wait(g753_ctx* ctx, g753_ctx* other);
/* the context's cudaStream_t, as an opaque pointer (for event timing by the harness) */
void* g753_stream(g753_ctx* ctx);

/* ---- introspection / test entry points -------------------------------------------------- */
/* out[i] = a[i] op b[i] over n host elements (b ignored for unary ops) */
int g753_field_op(g753_ctx* ctx, int field, int op, const uint64_t* a, const uint64_t* b,
                  uint64_t* out, size_t n);
/* group-law test hook: op 0: out = a + b (affine + affine), 1: 2a, 2: scalar * a,
 * 3: 2a + 2b through the full projective addition;
 * a, b affine (2*k*12 limbs, (0,0) = infinity), scalar 12 limbs canonical, out projective */
int g753_point_op(g753_ctx* ctx, int group, int op, const uint64_t* a, const uint64_t* b,
                  uint64_t* out_xyz);
/* coordinate-field test hook: out[i] = a[i] op b[i] over n elements of the group's coordinate field
 * (Fq, Fq2 = c0,c1 or Fq3 = c0,c1,c2; k*12 Montgomery limbs each) THROUGH THE TOWER CODE THE MSM KERNELS
 * RUN (fields/models/fp2.rs, fp3.rs on the lane-cooperative device towers); lanes = 0: the lane split
 * of the accumulation kernels, 1: that of the reduction / set-up kernels.
 * op: 0 mul, 1 add, 2 sub, 3 square, 4 neg, 5 inverse, 12 double, 20 / 21 mul with the result
 * aliasing a / b, 23 square in place.  The reference's Fq2 / Fq3 KATs are replayed through it. */
int g753_ext_op(g753_ctx* ctx, int group, int lanes, int op, const uint64_t* a, const uint64_t* b,
                uint64_t* out, size_t n);
/* test hook of the warp-cooperative field arithmetic the latency-bound kernels use (window fold, point folds,
 * reduction tail; csrc/coop.cuh): out[i] = op(a[i], b[i]) on 24-limb values of the BASE field of MNT4 (field 0) /
 * MNT6 (field 1), raw results (lazily reduced, NOT canonical) so that they compare limb for limb with the model
 * in tools/gen_coop.py.  op 0: Montgomery product a b / 2^768 (+ a multiple of p); 1: a + b; 2: a + 2^k p - b;
 * 3: a -> a mod p for a < 2 p; 4: out = 1 if a in {0, p} else 0. */
int g753_coop_op(g753_ctx* ctx, int field, int op, unsigned k, const uint64_t* a, const uint64_t* b, uint64_t* out,
                 size_t n);
/* Run `iters` dependent Montgomery multiplications per thread on `blocks` x `threads`
 * threads and report the kernel time in ms (CUDA events): the integer-pipe roofline probe
 * of SURVEY.md 8d.  variant 0 = fq_mul, 1 = fq_sqr, 2 = raw independent IMAD.WIDE stream,
 * 3 = fq_inv (safegcd), 4 = the instruction mix of one 753-bit Montgomery product on 15 x 52-bit limbs through the
 * FP64 pipe (2 DFMA + 1 DADD + 2 integer adds per limb product; a measurement of the other multiplier pipe, not a
 * multiplier) */
int g753_mac_probe(g753_ctx* ctx, int variant, int blocks, int threads, int iters, float* ms);
/* debugging aid: copy the first `bytes` of the MSM workspace to the host; *cap = its size */
int g753_debug_scratch(g753_ctx* ctx, void* h_dst, size_t bytes, size_t* cap);
/* number of kernel launches issued by this context since creation (bench.py gpu_launches) */
uint64_t g753_launch_count(const g753_ctx* ctx);
/* per-phase device times (ms) of the last g753_msm* call: digits, sort, accumulate, reduce,
 * combine - CUDA events on the context stream; returns the number of phases written */
int g753_last_msm_phases(g753_ctx* ctx, float* ms, int cap);
/* shape of the last g753_msm* call on this context: plan5 = {window bits c, windows W, bucket rows,
 * key copies used, accumulation form (0 = XYZZ running sums, 1 = pairwise tree of affine additions
 * with shared inversions)}.  The executed work of its accumulation phase is count * W additions
 * (bench.py reports it beside the reference op count the roofline is scored on). */
int g753_last_msm_plan(const g753_ctx* ctx, unsigned* plan5);

#ifdef __cplusplus
}
#endif
#endif /* G753_H */
