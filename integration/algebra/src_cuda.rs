//! `algebra/src/cuda.rs` (new file, feature `cuda`): the safe glue between ginger-lib's generic
//! `VariableBaseMSM` / `EvaluationDomain` and the C ABI of libg753.so (crate `algebra-cuda-sys`).
//!
//! Everything crosses the boundary in the reference's own limb layout - `BigInteger768.0` is
//! `[u64; 12]`, `Fp768.0` holds the Montgomery representation (fp_768.rs:24-30), MSM scalars are the
//! canonical `into_repr()` integers - so this file only FLATTENS structs; no number is converted.
//! Generic code reaches the four concrete curve types through `Any` (all curve and field types are
//! `'static`, curves/mod.rs:198-213), which keeps the crate `#![forbid(unsafe_code)]`.
//!
//! NOT COMPILED IN THE CUDA REPOSITORY'S IMAGE (no Rust toolchain there).
use crate::{
    biginteger::BigInteger768,
    curves::{
        mnt4753, mnt6753,
        models::{short_weierstrass_projective::{GroupAffine, GroupProjective}, SWModelParameters},
        AffineCurve,
    },
    fields::{models::{Fp2, Fp2Parameters, Fp3, Fp3Parameters, Fp768, Fp768Parameters}, PrimeField},
};
use algebra_cuda_sys as sys;
use std::any::{Any, TypeId};
use std::collections::HashMap;
use std::marker::PhantomData;
use std::sync::{Arc, Mutex};

const LIMBS: usize = sys::G753_LIMBS;

lazy_static::lazy_static! {
    /// One context per process (device `G753_DEVICE`, default 0), created on first use.  There is no
    /// CPU fallback inside the library: if no B200 is present this panics with the library's message,
    /// exactly as the feature `cuda` promises.
    pub static ref CTX: sys::Context = {
        let dev = std::env::var("G753_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        sys::Context::new(dev).unwrap_or_else(|e| panic!("g753_ctx_create failed: {:?}", e))
    };
    /// Resident proving-key queries, keyed by (group, address, length) of the bases slice with a
    /// fingerprint of its first and last point: `create_proof` passes sub-slices of the same
    /// `Parameters` vectors for every proof (prover.rs:273-325), so each query is uploaded (and its
    /// shifted copies built) once and found again by address.  `forget_keys()` drops them all; a caller
    /// that mutates a key in place must call it.
    static ref KEYS: Mutex<HashMap<(i32, usize, usize), (u64, Arc<sys::Bases>)>> = Mutex::new(HashMap::new());
}

/// Slices at least this long are treated as proving-key queries (kept resident, copies precomputed);
/// shorter ones go through the one-shot call.
pub const RESIDENT_MIN_LEN: usize = 1 << 14;

pub fn forget_keys() {
    KEYS.lock().unwrap().clear();
}

// ------------------------------------------------------------------------------------------------
// flattening: coordinate-field elements <-> limbs
// ------------------------------------------------------------------------------------------------
/// A coordinate field of one of the four groups: Fq, Fq2 (c0, c1) or Fq3 (c0, c1, c2) over an Fp768.
pub trait CoordLimbs: Sized {
    const K: usize;
    fn push_limbs(&self, out: &mut Vec<u64>);
    fn from_limbs(limbs: &[u64]) -> Self;
}

impl<P: Fp768Parameters> CoordLimbs for Fp768<P> {
    const K: usize = 1;
    fn push_limbs(&self, out: &mut Vec<u64>) {
        out.extend_from_slice(&(self.0).0);            // raw Montgomery limbs, fp_768.rs:24-30
    }
    fn from_limbs(limbs: &[u64]) -> Self {
        let mut a = [0u64; LIMBS];
        a.copy_from_slice(&limbs[..LIMBS]);
        Fp768::<P>(BigInteger768(a), PhantomData)      // already fully reduced Montgomery form
    }
}

impl<P: Fp2Parameters> CoordLimbs for Fp2<P> where P::Fp: CoordLimbs {
    const K: usize = 2;
    fn push_limbs(&self, out: &mut Vec<u64>) {
        self.c0.push_limbs(out);
        self.c1.push_limbs(out);
    }
    fn from_limbs(limbs: &[u64]) -> Self {
        Fp2::new(P::Fp::from_limbs(&limbs[..LIMBS]), P::Fp::from_limbs(&limbs[LIMBS..]))
    }
}

impl<P: Fp3Parameters> CoordLimbs for Fp3<P> where P::Fp: CoordLimbs {
    const K: usize = 3;
    fn push_limbs(&self, out: &mut Vec<u64>) {
        self.c0.push_limbs(out);
        self.c1.push_limbs(out);
        self.c2.push_limbs(out);
    }
    fn from_limbs(limbs: &[u64]) -> Self {
        Fp3::new(P::Fp::from_limbs(&limbs[..LIMBS]), P::Fp::from_limbs(&limbs[LIMBS..2 * LIMBS]),
                 P::Fp::from_limbs(&limbs[2 * LIMBS..]))
    }
}

/// The group id of `G` when it is one of the four groups libg753 implements (include/g753.h).
pub fn group_id<G: AffineCurve>() -> Option<i32> {
    let t = TypeId::of::<G>();
    if t == TypeId::of::<mnt4753::G1Affine>() { Some(sys::G753_MNT4_G1) }
    else if t == TypeId::of::<mnt4753::G2Affine>() { Some(sys::G753_MNT4_G2) }
    else if t == TypeId::of::<mnt6753::G1Affine>() { Some(sys::G753_MNT6_G1) }
    else if t == TypeId::of::<mnt6753::G2Affine>() { Some(sys::G753_MNT6_G2) }
    else { None }
}

/// The field id of `F` when it is one of the two 753-bit scalar fields.
pub fn field_id<F: PrimeField>() -> Option<i32> {
    let t = TypeId::of::<F>();
    if t == TypeId::of::<crate::fields::mnt4753::Fr>() { Some(sys::G753_FIELD_MNT4_FR) }
    else if t == TypeId::of::<crate::fields::mnt6753::Fr>() { Some(sys::G753_FIELD_MNT6_FR) }
    else { None }
}

fn flatten_concrete<P: SWModelParameters>(bases: &[GroupAffine<P>]) -> (Vec<u64>, Vec<u8>)
where P::BaseField: CoordLimbs {
    let k = <P::BaseField as CoordLimbs>::K;
    let mut coords = Vec::with_capacity(bases.len() * 2 * k * LIMBS);
    let mut inf = Vec::with_capacity(bases.len());
    for b in bases {                                   // rayon-parallel in a real build; O(n) copies
        b.x.push_limbs(&mut coords);
        b.y.push_limbs(&mut coords);
        inf.push(b.infinity as u8);
    }
    (coords, inf)
}

/// x || y limbs and infinity flags of `bases` (`G` is one of the four supported affine types).
pub fn flatten_affine<G: AffineCurve>(bases: &[G]) -> (Vec<u64>, Vec<u8>) {
    fn via<G: AffineCurve, P: SWModelParameters>(bases: &[G]) -> (Vec<u64>, Vec<u8>)
    where P::BaseField: CoordLimbs {
        // element-wise downcast: `G` IS `GroupAffine<P>` (checked by the caller through `group_id`)
        let concrete: Vec<GroupAffine<P>> =
            bases.iter().map(|b| *(b as &dyn Any).downcast_ref::<GroupAffine<P>>().expect("group id / type mismatch")).collect();
        flatten_concrete(&concrete)
    }
    match group_id::<G>().expect("unsupported curve") {
        sys::G753_MNT4_G1 => via::<G, mnt4753::g1::MNT4G1Parameters>(bases),
        sys::G753_MNT4_G2 => via::<G, mnt4753::g2::MNT4G2Parameters>(bases),
        sys::G753_MNT6_G1 => via::<G, mnt6753::g1::MNT6G1Parameters>(bases),
        _ => via::<G, mnt6753::g2::MNT6G2Parameters>(bases),
    }
}

/// `GroupProjective::new(x, y, z)` from the 3 k 12 limbs an MSM returns.
pub fn projective_from_limbs<G: AffineCurve>(xyz: &[u64]) -> G::Projective {
    fn via<G: AffineCurve, P: SWModelParameters>(xyz: &[u64]) -> G::Projective
    where P::BaseField: CoordLimbs {
        let w = <P::BaseField as CoordLimbs>::K * LIMBS;
        let p = GroupProjective::<P>::new(P::BaseField::from_limbs(&xyz[..w]), P::BaseField::from_limbs(&xyz[w..2 * w]),
                                          P::BaseField::from_limbs(&xyz[2 * w..]));
        let boxed: Box<dyn Any> = Box::new(p);
        *boxed.downcast::<G::Projective>().expect("group id / type mismatch")
    }
    match group_id::<G>().expect("unsupported curve") {
        sys::G753_MNT4_G1 => via::<G, mnt4753::g1::MNT4G1Parameters>(xyz),
        sys::G753_MNT4_G2 => via::<G, mnt4753::g2::MNT4G2Parameters>(xyz),
        sys::G753_MNT6_G1 => via::<G, mnt6753::g1::MNT6G1Parameters>(xyz),
        _ => via::<G, mnt6753::g2::MNT6G2Parameters>(xyz),
    }
}

/// Canonical scalars (`BigInteger768`) -> limbs.
pub fn flatten_scalars<B: AsRef<[u64]>>(scalars: &[B]) -> Vec<u64> {
    let mut out = Vec::with_capacity(scalars.len() * LIMBS);
    for s in scalars {
        out.extend_from_slice(s.as_ref());
    }
    out
}

/// Field elements (Montgomery form) <-> limbs, for the transforms.
pub fn flatten_field<F: PrimeField>(v: &[F]) -> Vec<u64> {
    let mut out = Vec::with_capacity(v.len() * LIMBS);
    for e in v {
        // F is mnt4753::Fr or mnt6753::Fr = Fp768<_> (checked by the caller through `field_id`)
        if let Some(x) = (e as &dyn Any).downcast_ref::<crate::fields::mnt4753::Fr>() { out.extend_from_slice(&(x.0).0); }
        else if let Some(x) = (e as &dyn Any).downcast_ref::<crate::fields::mnt6753::Fr>() { out.extend_from_slice(&(x.0).0); }
        else { panic!("unsupported field"); }
    }
    out
}

pub fn unflatten_field<F: PrimeField>(limbs: &[u64], out: &mut [F]) {
    for (i, e) in out.iter_mut().enumerate() {
        let mut a = [0u64; LIMBS];
        a.copy_from_slice(&limbs[i * LIMBS..(i + 1) * LIMBS]);
        let any: &mut dyn Any = e;
        if let Some(x) = any.downcast_mut::<crate::fields::mnt4753::Fr>() { x.0 = BigInteger768(a); continue; }
        let any: &mut dyn Any = e;
        if let Some(x) = any.downcast_mut::<crate::fields::mnt6753::Fr>() { x.0 = BigInteger768(a); continue; }
        panic!("unsupported field");
    }
}

fn fingerprint(coords: &[u64], k: usize) -> u64 {
    let w = 2 * k * LIMBS;
    let mut h = 0xcbf29ce484222325u64;
    for v in coords[..w.min(coords.len())].iter().chain(coords[coords.len().saturating_sub(w)..].iter()) {
        h = (h ^ v).wrapping_mul(0x100000001b3);
    }
    h
}

/// `VariableBaseMSM::multi_scalar_mul` on the GPU (variable_base.rs:85-90): zip semantics, a resident
/// handle for long (proving-key) slices, the one-shot call otherwise.  `None`: `G` is not one of the
/// four supported groups and the caller keeps the CPU path.
pub fn multi_scalar_mul<G: AffineCurve>(bases: &[G], scalars: &[<G::ScalarField as PrimeField>::BigInt])
                                        -> Option<G::Projective> {
    let group = group_id::<G>()?;
    let n = core::cmp::min(bases.len(), scalars.len());       // scalars.iter().zip(bases), variable_base.rs:36
    let sc = flatten_scalars(&scalars[..n]);
    let xyz = if n >= RESIDENT_MIN_LEN {
        let key = (group, bases.as_ptr() as usize, bases.len());
        let cached = KEYS.lock().unwrap().get(&key).cloned();
        let (coords, inf) = if cached.is_none() { flatten_affine(bases) } else { (Vec::new(), Vec::new()) };
        let handle = match cached {
            // the fingerprint of a cached key is re-checked against the live slice's first / last point
            Some((fp, h)) if fp == fingerprint(&flatten_affine(&[bases[0], bases[bases.len() - 1]]).0, sys::group_k(group)) => h,
            _ => {
                let (coords, inf) = if coords.is_empty() { flatten_affine(bases) } else { (coords, inf) };
                let mut h = CTX.upload_bases(group, &coords, &inf).unwrap_or_else(|e| panic!("g753_bases_upload: {:?}", e));
                h.precompute(0).unwrap_or_else(|e| panic!("g753_bases_precompute: {:?}", e));
                let h = Arc::new(h);
                let fp = fingerprint(&flatten_affine(&[bases[0], bases[bases.len() - 1]]).0, sys::group_k(group));
                KEYS.lock().unwrap().insert(key, (fp, h.clone()));
                h
            }
        };
        CTX.msm(&handle, 0, &sc)
    } else {
        let (coords, inf) = flatten_affine(&bases[..n]);
        CTX.msm_host(group, &coords, &inf, &sc)
    }
    .unwrap_or_else(|e| panic!("g753 MSM failed: {:?}", e));  // the reference's signature is infallible
    Some(projective_from_limbs::<G>(&xyz))
}

/// One of the four `EvaluationDomain` transforms in place (domain.rs:120-179); `coeffs` already resized.
/// `false`: `F` is not one of the two 753-bit scalar fields and the caller keeps the CPU path.
pub fn transform<F: PrimeField>(coeffs: &mut [F], log_n: u32, mode: i32) -> bool {
    let field = match field_id::<F>() { Some(f) => f, None => return false };
    let mut limbs = flatten_field(coeffs);
    CTX.ntt(field, &mut limbs, log_n, mode).unwrap_or_else(|e| panic!("g753_ntt failed: {:?}", e));
    unflatten_field(&limbs, coeffs);
    true
}
