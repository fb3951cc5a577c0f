//! Raw bindings of `include/g753.h` plus the safe wrappers `algebra` calls (the `algebra` crate is
//! `#![forbid(unsafe_code)]`, algebra/src/lib.rs:34, so every `unsafe` of the integration lives here).
//!
//! NOT COMPILED IN THE CUDA REPOSITORY'S IMAGE (no Rust toolchain there); the same symbols, argument
//! orders and types are exercised from Python (ginger-lib_b200/ffi.py) and from C
//! (integration/c_caller/abi_caller.c), and a test asserts that this file declares every function of
//! the header (tests/test_abi.py).
use std::ffi::CStr;
use std::os::raw::{c_char, c_float, c_int, c_uint, c_void};

#[repr(C)]
pub struct G753Ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct G753Bases {
    _private: [u8; 0],
}
#[repr(C)]
pub struct G753NttShard {
    _private: [u8; 0],
}

pub const G753_OK: c_int = 0;
pub const G753_ERR_BAD_ARG: c_int = 1;
pub const G753_ERR_CUDA: c_int = 2;
pub const G753_ERR_OOM: c_int = 3;
pub const G753_ERR_DOMAIN: c_int = 4;
pub const G753_ERR_NO_DEVICE: c_int = 5;
pub const G753_LIMBS: usize = 12;
pub const G753_MNT4_G1: c_int = 0;
pub const G753_MNT4_G2: c_int = 1;
pub const G753_MNT6_G1: c_int = 2;
pub const G753_MNT6_G2: c_int = 3;
pub const G753_FIELD_MNT6_FR: c_int = 0;
pub const G753_FIELD_MNT4_FR: c_int = 1;
pub const G753_FFT: c_int = 0;
pub const G753_IFFT: c_int = 1;
pub const G753_COSET_FFT: c_int = 2;
pub const G753_COSET_IFFT: c_int = 3;
pub const G753_OP_FROM_MONT: c_int = 7;

extern "C" {
    pub fn g753_device_count(count: *mut c_int) -> c_int;
    pub fn g753_ctx_create(device: c_int, out: *mut *mut G753Ctx) -> c_int;
    pub fn g753_ctx_destroy(ctx: *mut G753Ctx) -> c_int;
    pub fn g753_ctx_set_stream(ctx: *mut G753Ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn g753_last_error() -> *const c_char;
    pub fn g753_version() -> *const c_char;
    pub fn g753_source_hash() -> *const c_char;
    pub fn g753_bases_upload(ctx: *mut G753Ctx, group: c_int, coords: *const u64, infinity: *const u8, n: usize,
                             out: *mut *mut G753Bases) -> c_int;
    pub fn g753_bases_upload_wire(ctx: *mut G753Ctx, group: c_int, wire: *const u8, n: usize,
                                  out: *mut *mut G753Bases) -> c_int;
    pub fn g753_bases_free(ctx: *mut G753Ctx, bases: *mut G753Bases) -> c_int;
    pub fn g753_bases_generate(ctx: *mut G753Ctx, group: c_int, gen_xy: *const u64, seed: u64, n: usize,
                               out: *mut *mut G753Bases) -> c_int;
    pub fn g753_bases_precompute(ctx: *mut G753Ctx, bases: *mut G753Bases, copies: c_uint) -> c_int;
    pub fn g753_bases_update(ctx: *mut G753Ctx, bases: *mut G753Bases, first: usize, count: usize, coords: *const u64,
                             infinity: *const u8) -> c_int;
    pub fn g753_bases_download(ctx: *mut G753Ctx, bases: *const G753Bases, first: usize, count: usize,
                               coords: *mut u64) -> c_int;
    pub fn g753_bases_len(bases: *const G753Bases) -> usize;
    pub fn g753_msm(ctx: *mut G753Ctx, bases: *const G753Bases, first: usize, count: usize, scalars: *const u64,
                    out_xyz: *mut u64) -> c_int;
    pub fn g753_msm_dev(ctx: *mut G753Ctx, bases: *const G753Bases, first: usize, count: usize,
                        d_scalars: *const c_void, d_out_xyz: *mut c_void) -> c_int;
    pub fn g753_msm_host(ctx: *mut G753Ctx, group: c_int, coords: *const u64, infinity: *const u8, n_bases: usize,
                         scalars: *const u64, n_scalars: usize, out_xyz: *mut u64) -> c_int;
    pub fn g753_points_sum_dev(ctx: *mut G753Ctx, group: c_int, d_points_xyz: *const c_void, count: usize,
                               d_out_xyz: *mut c_void) -> c_int;
    pub fn g753_batch_normalize(ctx: *mut G753Ctx, group: c_int, xyz: *const u64, count: usize, xy: *mut u64,
                                infinity: *mut u8) -> c_int;
    pub fn g753_fixed_base_msm(ctx: *mut G753Ctx, group: c_int, base_xy: *const u64, scalars: *const u64, n: usize,
                               out_xy: *mut u64, out_infinity: *mut u8) -> c_int;
    pub fn g753_group_coord_limbs(group: c_int) -> c_int;
    pub fn g753_domain_check(field: c_int, log_n: c_uint) -> c_int;
    pub fn g753_domain_constant(ctx: *mut G753Ctx, field: c_int, log_n: c_uint, which: c_int, out: *mut u64) -> c_int;
    pub fn g753_ntt(ctx: *mut G753Ctx, field: c_int, data: *mut u64, log_n: c_uint, mode: c_int) -> c_int;
    pub fn g753_ntt_dev(ctx: *mut G753Ctx, field: c_int, d_data: *mut c_void, log_n: c_uint, mode: c_int) -> c_int;
    pub fn g753_vec_op_dev(ctx: *mut G753Ctx, field: c_int, op: c_int, d_a: *mut c_void, d_b: *const c_void,
                           n: usize) -> c_int;
    pub fn g753_vec_scale_dev(ctx: *mut G753Ctx, field: c_int, d_a: *mut c_void, k_mont: *const u64, n: usize) -> c_int;
    pub fn g753_domain_check_mixed(field: c_int, n: u64) -> c_int;
    pub fn g753_ntt_mixed(ctx: *mut G753Ctx, field: c_int, data: *mut u64, n: u64, mode: c_int) -> c_int;
    pub fn g753_ntt_mixed_dev(ctx: *mut G753Ctx, field: c_int, d_data: *mut c_void, n: u64, mode: c_int) -> c_int;
    pub fn g753_ntt_shard_create(ctx: *mut G753Ctx, field: c_int, log_n: c_uint, world: c_uint, rank: c_uint,
                                 out: *mut *mut G753NttShard) -> c_int;
    pub fn g753_ntt_shard_destroy(ctx: *mut G753Ctx, plan: *mut G753NttShard) -> c_int;
    pub fn g753_ntt_shard_shape(plan: *const G753NttShard, n1: *mut usize, n2: *mut usize, cols: *mut usize,
                                rows: *mut usize) -> c_int;
    pub fn g753_ntt_shard_step1(ctx: *mut G753Ctx, plan: *const G753NttShard, d_data: *mut c_void, d_send: *mut c_void,
                                mode: c_int) -> c_int;
    pub fn g753_ntt_shard_step2(ctx: *mut G753Ctx, plan: *const G753NttShard, d_recv: *const c_void,
                                d_data: *mut c_void, mode: c_int) -> c_int;
    pub fn g753_ntt_shard_step1_fused(ctx: *mut G753Ctx, plan: *const G753NttShard, d_data: *mut c_void,
                                      peer_z: *const *mut c_void, mode: c_int) -> c_int;
    pub fn g753_ntt_shard_step2_local(ctx: *mut G753Ctx, plan: *const G753NttShard, d_z: *mut c_void, mode: c_int) -> c_int;
    pub fn g753_witness_map(ctx: *mut G753Ctx, field: c_int, a: *const u64, b: *const u64, c: *const u64,
                            log_n: c_uint, d123_mont: *const u64, h: *mut u64) -> c_int;
    pub fn g753_witness_map_dev(ctx: *mut G753Ctx, field: c_int, d_a: *mut c_void, d_b: *mut c_void, d_c: *mut c_void,
                                log_n: c_uint, d123_mont: *const u64, d_h: *mut c_void) -> c_int;
    pub fn g753_witness_map_tail_dev(ctx: *mut G753Ctx, field: c_int, d_a: *mut c_void, d_b: *const c_void,
                                     d_c: *const c_void, log_n: c_uint, d123_mont: *const u64, d_h: *mut c_void) -> c_int;
    pub fn g753_dev_alloc(ctx: *mut G753Ctx, bytes: usize, d_ptr: *mut *mut c_void) -> c_int;
    pub fn g753_dev_free(ctx: *mut G753Ctx, d_ptr: *mut c_void) -> c_int;
    pub fn g753_h2d(ctx: *mut G753Ctx, d_dst: *mut c_void, h_src: *const c_void, bytes: usize) -> c_int;
    pub fn g753_d2h(ctx: *mut G753Ctx, h_dst: *mut c_void, d_src: *const c_void, bytes: usize) -> c_int;
    pub fn g753_d2d(ctx: *mut G753Ctx, d_dst: *mut c_void, d_src: *const c_void, bytes: usize) -> c_int;
    pub fn g753_sync(ctx: *mut G753Ctx) -> c_int;
    pub fn g753_ctx_wait(ctx: *mut G753Ctx, other: *mut G753Ctx) -> c_int;
    pub fn g753_stream(ctx: *mut G753Ctx) -> *mut c_void;
    pub fn g753_field_op(ctx: *mut G753Ctx, field: c_int, op: c_int, a: *const u64, b: *const u64, out: *mut u64,
                         n: usize) -> c_int;
    pub fn g753_point_op(ctx: *mut G753Ctx, group: c_int, op: c_int, a: *const u64, b: *const u64,
                         out_xyz: *mut u64) -> c_int;
    pub fn g753_ext_op(ctx: *mut G753Ctx, group: c_int, lanes: c_int, op: c_int, a: *const u64, b: *const u64,
                       out: *mut u64, n: usize) -> c_int;
    pub fn g753_coop_op(ctx: *mut G753Ctx, field: c_int, op: c_int, k: c_uint, a: *const u64, b: *const u64, out: *mut u64,
                        n: usize) -> c_int;
    pub fn g753_mac_probe(ctx: *mut G753Ctx, variant: c_int, blocks: c_int, threads: c_int, iters: c_int,
                          ms: *mut c_float) -> c_int;
    pub fn g753_debug_scratch(ctx: *mut G753Ctx, h_dst: *mut c_void, bytes: usize, cap: *mut usize) -> c_int;
    pub fn g753_launch_count(ctx: *const G753Ctx) -> u64;
    pub fn g753_last_msm_phases(ctx: *mut G753Ctx, ms: *mut c_float, cap: c_int) -> c_int;
    pub fn g753_last_msm_plan(ctx: *const G753Ctx, plan5: *mut c_uint) -> c_int;
}

/// An error code of the C ABI with the library's thread-local description of it.
#[derive(Debug, Clone)]
pub struct Error {
    pub code: i32,
    pub message: String,
}

fn check(rc: c_int) -> Result<(), Error> {
    if rc == G753_OK {
        return Ok(());
    }
    let message = unsafe {
        let p = g753_last_error();
        if p.is_null() { String::new() } else { CStr::from_ptr(p).to_string_lossy().into_owned() }
    };
    Err(Error { code: rc, message })
}

/// `(3 - limbs per coordinate element)` helper: k = 1, 2, 3 for G1, G2/Fq2, G2/Fq3.
pub fn group_k(group: i32) -> usize {
    (unsafe { g753_group_coord_limbs(group) } as usize) / G753_LIMBS
}

/// One GPU: stream, scratch memory, twiddle tables.  Calls on one context are serialised by a mutex on
/// the C side, so a `Context` may be shared between threads.
pub struct Context(*mut G753Ctx);
unsafe impl Send for Context {}
unsafe impl Sync for Context {}

impl Context {
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut p = std::ptr::null_mut();
        check(unsafe { g753_ctx_create(device, &mut p) })?;
        Ok(Context(p))
    }

    /// `VariableBaseMSM::multi_scalar_mul(&bases[..n], &scalars[..m])` with host bases:
    /// coords = x || y limbs per point (2 k 12 u64, Montgomery form), infinity = one byte per point,
    /// scalars = canonical 12-limb integers.  Returns X, Y, Z (3 k 12 limbs, Montgomery form).
    pub fn msm_host(&self, group: i32, coords: &[u64], infinity: &[u8], scalars: &[u64]) -> Result<Vec<u64>, Error> {
        let k = group_k(group);
        assert_eq!(coords.len(), infinity.len() * 2 * k * G753_LIMBS);
        assert_eq!(scalars.len() % G753_LIMBS, 0);
        let mut out = vec![0u64; 3 * k * G753_LIMBS];
        check(unsafe {
            g753_msm_host(self.0, group, coords.as_ptr(), infinity.as_ptr(), infinity.len(), scalars.as_ptr(),
                          scalars.len() / G753_LIMBS, out.as_mut_ptr())
        })?;
        Ok(out)
    }

    /// Upload a proving-key query once; it stays resident in HBM behind the returned handle.
    pub fn upload_bases(&self, group: i32, coords: &[u64], infinity: &[u8]) -> Result<Bases, Error> {
        let k = group_k(group);
        assert_eq!(coords.len(), infinity.len() * 2 * k * G753_LIMBS);
        let mut h = std::ptr::null_mut();
        check(unsafe { g753_bases_upload(self.0, group, coords.as_ptr(), infinity.as_ptr(), infinity.len(), &mut h) })?;
        Ok(Bases { ctx: self.0, handle: h, group, len: infinity.len() })
    }

    /// The same from the reference's serialisation (`GroupAffine::write` records).
    pub fn upload_bases_wire(&self, group: i32, wire: &[u8]) -> Result<Bases, Error> {
        let rec = 2 * group_k(group) * 96 + 1;
        assert_eq!(wire.len() % rec, 0);
        let n = wire.len() / rec;
        let mut h = std::ptr::null_mut();
        check(unsafe { g753_bases_upload_wire(self.0, group, wire.as_ptr(), n, &mut h) })?;
        Ok(Bases { ctx: self.0, handle: h, group, len: n })
    }

    /// sum_i scalars[i] * bases[first + i] over `min(bases.len() - first, scalars.len() / 12)` terms.
    pub fn msm(&self, bases: &Bases, first: usize, scalars: &[u64]) -> Result<Vec<u64>, Error> {
        let k = group_k(bases.group);
        let count = std::cmp::min(bases.len.saturating_sub(first), scalars.len() / G753_LIMBS);
        let mut out = vec![0u64; 3 * k * G753_LIMBS];
        check(unsafe { g753_msm(self.0, bases.handle, first, count, scalars.as_ptr(), out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// `EvaluationDomain::new` would return `None` for this size.
    pub fn domain_exists(field: i32, log_n: u32) -> bool {
        unsafe { g753_domain_check(field, log_n) == G753_OK }
    }

    /// In-place transform of `2^log_n` elements (12 Montgomery limbs each), `mode` = G753_FFT ..
    pub fn ntt(&self, field: i32, data: &mut [u64], log_n: u32, mode: i32) -> Result<(), Error> {
        assert_eq!(data.len(), G753_LIMBS << log_n);
        check(unsafe { g753_ntt(self.0, field, data.as_mut_ptr(), log_n, mode) })
    }

    /// `R1CStoQAP::witness_map` from the evaluated constraints onwards (r1cs_to_qap.rs:121-166).
    pub fn witness_map(&self, field: i32, a: &[u64], b: &[u64], c: &[u64], log_n: u32, d123_mont: &[u64; 36])
                       -> Result<Vec<u64>, Error> {
        let n = 1usize << log_n;
        assert!(a.len() == n * G753_LIMBS && b.len() == a.len() && c.len() == a.len());
        let mut h = vec![0u64; (n + 1) * G753_LIMBS];
        check(unsafe {
            g753_witness_map(self.0, field, a.as_ptr(), b.as_ptr(), c.as_ptr(), log_n, d123_mont.as_ptr(), h.as_mut_ptr())
        })?;
        Ok(h)
    }

    /// `batch_normalization` + `into_affine`: X, Y, Z -> x, y, infinity.
    pub fn batch_normalize(&self, group: i32, xyz: &[u64]) -> Result<(Vec<u64>, Vec<u8>), Error> {
        let k = group_k(group);
        let count = xyz.len() / (3 * k * G753_LIMBS);
        let mut xy = vec![0u64; count * 2 * k * G753_LIMBS];
        let mut inf = vec![0u8; count];
        check(unsafe { g753_batch_normalize(self.0, group, xyz.as_ptr(), count, xy.as_mut_ptr(), inf.as_mut_ptr()) })?;
        Ok((xy, inf))
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe {
            g753_ctx_destroy(self.0);
        }
    }
}

/// A device-resident slice of proving-key bases.  Must not outlive its `Context`.
pub struct Bases {
    ctx: *mut G753Ctx,
    handle: *mut G753Bases,
    pub group: i32,
    pub len: usize,
}
unsafe impl Send for Bases {}
unsafe impl Sync for Bases {}

impl Bases {
    /// Once per key: shifted copies 2^(j rows c) P_i (0 = as many as fit the 6 GiB per-key budget).
    pub fn precompute(&mut self, copies: u32) -> Result<(), Error> {
        check(unsafe { g753_bases_precompute(self.ctx, self.handle, copies) })
    }
}

impl Drop for Bases {
    fn drop(&mut self) {
        unsafe {
            g753_bases_free(self.ctx, self.handle);
        }
    }
}
