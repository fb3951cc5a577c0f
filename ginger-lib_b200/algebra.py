"""Host-side mirror of the reference's operator interface for the hot path.

Same names, argument meaning and error behaviour as ginger-lib's
  * `VariableBaseMSM::multi_scalar_mul(bases, scalars) -> G::Projective`
    (algebra/src/msm/variable_base.rs:85-90), and
  * `EvaluationDomain<F>` with `new`, `size`, `fft[_in_place]`, `ifft[_in_place]`,
    `coset_fft[_in_place]`, `coset_ifft[_in_place]` (algebra/src/fft/domain.rs:65-179),
so the parity tests read like the reference's own tests.  All arithmetic happens in
libg753.so on the GPU; this file only shapes buffers (numpy uint64, the reference's raw limb
layout: 12 little-endian u64 per Fq, Montgomery form for field/point coordinates, canonical
integers for MSM scalars).
"""
import ctypes

import numpy as np

from . import ffi

LIMBS = 12


class Context:
    """One GPU (device index) with its stream, scratch memory and twiddle tables."""

    def __init__(self, device=0, library=None, stream=None):
        """`stream`: optional raw cudaStream_t (int), e.g. `torch.cuda.Stream().cuda_stream`, so that
        the library's kernels are ordered with the caller's own work (NCCL collectives, events)."""
        self.lib = library or ffi.default_library()
        h = ctypes.c_void_p()
        self.lib.check(self.lib.ctx_create(int(device), ctypes.byref(h)))
        self.handle = h
        self.device = device
        if stream is not None:
            self.lib.check(self.lib.ctx_set_stream(self.handle, ctypes.c_void_p(int(stream))))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- proving-key bases ------------------------------------------------------------
    def upload_bases(self, group, coords, infinity=None):
        return Bases(self, group, coords, infinity)

    def generate_bases(self, group, n, seed):
        """synthetic resident key: bases[i] = a_i * G (see Bases.generated_logs)"""
        return Bases.generate(self, group, n, seed)

    def sync(self):
        self.lib.check(self.lib.sync(self.handle))

    @property
    def launches(self):
        return int(self.lib.launch_count(self.handle))

    def last_msm_plan(self):
        """window bits, windows, bucket rows, key copies and accumulation form of the last MSM on this context"""
        buf = (ctypes.c_uint * 5)()
        self.lib.check(self.lib.last_msm_plan(self.handle, buf))
        return {"c": int(buf[0]), "windows": int(buf[1]), "rows": int(buf[2]), "copies": int(buf[3]),
                "accumulation": "affine_tree" if buf[4] else "xyzz_running_sums"}

    def last_msm_phases(self):
        buf = (ctypes.c_float * 8)()
        k = self.lib.last_msm_phases(self.handle, buf, 8)
        names = ["digits", "sort", "accumulate", "reduce", "combine"]
        return {names[i]: float(buf[i]) for i in range(k)}


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class Bases:
    """Device-resident affine bases (`&[G::Affine]`), uploaded once per proving key."""

    def __init__(self, ctx, group, coords, infinity=None):
        k = ffi.GROUP_K[group]
        coords = ffi.as_u64(coords).reshape(-1, 2 * k * LIMBS)
        n = coords.shape[0]
        inf = None
        if infinity is not None:
            inf = np.ascontiguousarray(infinity, dtype=np.uint8).reshape(-1)
            if inf.shape[0] != n:
                raise ValueError("infinity flags length mismatch")
        self.ctx, self.group, self.k, self.n = ctx, group, k, n
        h = ctypes.c_void_p()
        ctx.lib.check(ctx.lib.bases_upload(ctx.handle, group, ffi.ptr(coords), ffi.ptr(inf), n, ctypes.byref(h)))
        self.handle = h

    def __len__(self):
        return self.n

    @classmethod
    def from_wire(cls, ctx, group, data):
        """bases from the reference's serialisation (`GroupAffine::write`: x || y || infinity byte,
        canonical little-endian coordinates) - the proving-key loader of SURVEY.md 8f-2"""
        k = ffi.GROUP_K[group]
        rec = 2 * k * 96 + 1
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        if buf.size % rec:
            raise ValueError("wire data is not a whole number of %d-byte points" % rec)
        self = cls.__new__(cls)
        self.ctx, self.group, self.k, self.n = ctx, group, k, buf.size // rec
        h = ctypes.c_void_p()
        ctx.lib.check(ctx.lib.bases_upload_wire(ctx.handle, group, ffi.ptr(np.ascontiguousarray(buf)), self.n,
                                                ctypes.byref(h)))
        self.handle = h
        return self

    @classmethod
    def generate(cls, ctx, group, n, seed):
        from . import params
        self = cls.__new__(cls)
        self.ctx, self.group, self.k, self.n = ctx, group, ffi.GROUP_K[group], int(n)
        gen = np.array([[(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(LIMBS)]
                        for v in params.GENERATOR_MONT[group]], dtype=np.uint64)
        h = ctypes.c_void_p()
        ctx.lib.check(ctx.lib.bases_generate(ctx.handle, group, ffi.ptr(gen), ctypes.c_uint64(seed), self.n,
                                             ctypes.byref(h)))
        self.handle = h
        return self

    @staticmethod
    def generated_logs(n, seed, first=0):
        """a_i of `generate`: splitmix64(seed + (i+1) * golden) | 1, as a numpy uint64 array"""
        with np.errstate(over="ignore"):
            i = np.arange(first + 1, first + n + 1, dtype=np.uint64)
            z = np.uint64(seed) + i * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return (z ^ (z >> np.uint64(31))) | np.uint64(1)

    def precompute(self, copies=0):
        """build the shifted key copies (g753_bases_precompute); MSM results are unchanged"""
        self.ctx.lib.check(self.ctx.lib.bases_precompute(self.ctx.handle, self.handle, int(copies)))
        return self

    def download(self, first=0, count=None):
        count = self.n - first if count is None else count
        out = np.zeros((count, 2 * self.k * LIMBS), dtype=np.uint64)
        self.ctx.lib.check(self.ctx.lib.bases_download(self.ctx.handle, self.handle, first, count, ffi.ptr(out)))
        return out

    def free(self):
        if getattr(self, "handle", None):
            self.ctx.lib.bases_free(self.ctx.handle, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class VariableBaseMSM:
    """algebra/src/msm/variable_base.rs:7-90."""

    @staticmethod
    def multi_scalar_mul(bases, scalars, group=None, infinity=None, ctx=None, first=0):
        """sum_i scalars[i] * bases[i] over zip(bases, scalars) -> projective (X, Y, Z) limbs.

        `bases` is either a resident `Bases` handle (the proving-key case; `first` selects
        the sub-slice view `bases[first..]` the prover takes, groth16/mod.rs:318-350) or a
        numpy array of affine coordinates (+ `infinity` flags, + `group`), which is the
        reference's one-shot signature.  `scalars`: (m, 12) uint64 canonical `BigInteger768`.
        Returns a (3, k*12) uint64 array: the reference's GroupProjective {x, y, z}.
        """
        scalars = ffi.as_u64(scalars).reshape(-1, LIMBS)
        if isinstance(bases, Bases):
            ctx = bases.ctx
            k = bases.k
            avail = len(bases) - first
            if avail < 0:
                raise ValueError("slice start beyond the bases")
            count = min(avail, scalars.shape[0])       # zip semantics, variable_base.rs:36
            out = np.zeros((3, k * LIMBS), dtype=np.uint64)
            ctx.lib.check(ctx.lib.msm(ctx.handle, bases.handle, first, count, ffi.ptr(scalars), ffi.ptr(out)))
            return out
        if group is None:
            raise ValueError("group id required with host bases")
        ctx = ctx or default_context()
        k = ffi.GROUP_K[group]
        coords = ffi.as_u64(bases).reshape(-1, 2 * k * LIMBS)[first:]
        inf = None
        if infinity is not None:
            inf = np.ascontiguousarray(infinity, dtype=np.uint8).reshape(-1)[first:]
        out = np.zeros((3, k * LIMBS), dtype=np.uint64)
        ctx.lib.check(ctx.lib.msm_host(ctx.handle, group, ffi.ptr(coords), ffi.ptr(inf), coords.shape[0],
                                       ffi.ptr(scalars), scalars.shape[0], ffi.ptr(out)))
        return out


class FixedBaseMSM:
    """algebra/src/msm/fixed_base.rs:4-80 with the normalisation the key generator applies to its
    results (groth16/generator.rs:225-319): `multi_scalar_mul(group, base, scalars)` returns the affine
    points scalars[i] * base, (n, 2*k*12) Montgomery limbs + infinity flags."""

    @staticmethod
    def multi_scalar_mul(group, base_xy, scalars, ctx=None):
        ctx = ctx or default_context()
        k = ffi.GROUP_K[group]
        base = ffi.as_u64(base_xy).reshape(2 * k * LIMBS)
        scalars = ffi.as_u64(scalars).reshape(-1, LIMBS)
        n = scalars.shape[0]
        out = np.zeros((n, 2 * k * LIMBS), dtype=np.uint64)
        inf = np.zeros(n, dtype=np.uint8)
        ctx.lib.check(ctx.lib.fixed_base_msm(ctx.handle, group, ffi.ptr(base), ffi.ptr(scalars), n, ffi.ptr(out),
                                             ffi.ptr(inf)))
        return out, inf


class EvaluationDomain:
    """algebra/src/fft/domain.rs:24-179 for the two 753-bit scalar fields.

    `field` is ffi.FIELD_MNT4_FR (two-adicity 30) or ffi.FIELD_MNT6_FR (two-adicity 15).
    """

    def __init__(self, field, size, log_size, ctx):
        self.field = field
        self.size_ = size
        self.log_size_of_group = log_size
        self.ctx = ctx

    @staticmethod
    def new(field, num_coeffs, ctx=None):
        """`EvaluationDomain::new`: None when 2^ceil(log2 n) is too large for the field
        (domain.rs:65-72)."""
        size = 1
        while size < num_coeffs:
            size <<= 1
        log_size = size.bit_length() - 1
        lib = (ctx.lib if ctx else ffi.default_library())
        rc = lib.domain_check(field, log_size)
        if rc == ffi.ERR_DOMAIN:
            return None
        lib.check(rc)
        return EvaluationDomain(field, size, log_size, ctx or default_context())

    @staticmethod
    def compute_size_of_domain(field, num_coeffs, library=None):
        size = 1
        while size < num_coeffs:
            size <<= 1
        lib = library or ffi.default_library()
        return size if lib.domain_check(field, size.bit_length() - 1) == ffi.OK else None

    def size(self):
        return self.size_

    def _resized(self, v):
        """Vec::resize(size, zero): truncate or zero-pad (domain.rs:121)."""
        v = ffi.as_u64(v).reshape(-1, LIMBS)
        out = np.zeros((self.size_, LIMBS), dtype=np.uint64)
        m = min(self.size_, v.shape[0])
        out[:m] = v[:m]
        return out

    def _run(self, v, mode):
        buf = self._resized(v)
        self.ctx.lib.check(self.ctx.lib.ntt(self.ctx.handle, self.field, ffi.ptr(buf), self.log_size_of_group, mode))
        return buf

    def fft(self, coeffs):
        return self._run(coeffs, ffi.FFT)

    def ifft(self, evals):
        return self._run(evals, ffi.IFFT)

    def coset_fft(self, coeffs):
        return self._run(coeffs, ffi.COSET_FFT)

    def coset_ifft(self, evals):
        return self._run(evals, ffi.COSET_IFFT)

    # the public fields of the struct (domain.rs:24-39), 12 Montgomery limbs each, computed on the device
    _CONSTANTS = {"size_inv": 0, "group_gen": 1, "group_gen_inv": 2, "generator_inv": 3,
                  "vanishing_on_coset_inv": 4}

    def __getattr__(self, name):
        which = EvaluationDomain._CONSTANTS.get(name)
        if which is None:
            raise AttributeError(name)
        out = np.zeros(LIMBS, dtype=np.uint64)
        self.ctx.lib.check(self.ctx.lib.domain_constant(self.ctx.handle, self.field, self.log_size_of_group,
                                                        which, ffi.ptr(out)))
        setattr(self, name, out)
        return out

    def mul_polynomials_in_evaluation_domain(self, self_evals, other_evals):
        """Point-wise product of two evaluation vectors over the domain (domain.rs:289-302)."""
        a = ffi.as_u64(self_evals).reshape(-1, LIMBS)
        b = ffi.as_u64(other_evals).reshape(-1, LIMBS)
        assert a.shape == b.shape
        da = DeviceVector(self.ctx, self.field, a.shape[0], a)
        db = DeviceVector(self.ctx, self.field, b.shape[0], b)
        try:
            da.op(ffi.OP_MUL, db)
            return da.download()
        finally:
            da.free()
            db.free()

    def divide_by_vanishing_poly_on_coset_in_place(self, evals):
        """evals *= (17^n - 1)^-1: Z is constant on the coset (domain.rs:245-256)."""
        v = ffi.as_u64(evals).reshape(-1, LIMBS)
        d = DeviceVector(self.ctx, self.field, v.shape[0], v)
        try:
            d.scale(self.vanishing_on_coset_inv)
            return d.download()
        finally:
            d.free()

    # ---- the rest of the struct's methods (domain.rs:183-287); field arithmetic on the device ----
    def _fop(self, op, a, b=None):
        out = np.zeros((1, LIMBS), dtype=np.uint64)
        a = np.ascontiguousarray(ffi.as_u64(a).reshape(1, LIMBS))
        b = np.ascontiguousarray(ffi.as_u64(b).reshape(1, LIMBS)) if b is not None else None
        self.ctx.lib.check(self.ctx.lib.field_op(self.ctx.handle, self.field, op, ffi.ptr(a), ffi.ptr(b), ffi.ptr(out), 1))
        return out

    def _one(self):
        one = np.zeros((1, LIMBS), dtype=np.uint64)
        one[0, 0] = 1
        return self._fop(ffi.OP_TO_MONT, one)

    def _pow_size(self, tau):
        """tau^size by log2(size) squarings (`tau.pow(&[self.size])`)"""
        t = ffi.as_u64(tau).reshape(1, LIMBS)
        for _ in range(self.log_size_of_group):
            t = self._fop(ffi.OP_SQR, t)
        return t

    def elements(self):
        """the domain's elements 1, g, g^2, ... (domain.rs:234-240, 419-441) as an (n, 12) Montgomery
        array: the transform of the unit vector e_1"""
        if self.size_ == 1:
            return self._one()
        e1 = np.zeros((self.size_, LIMBS), dtype=np.uint64)
        e1[1] = self._one()[0]
        return self.fft(e1)

    def evaluate_vanishing_polynomial(self, tau):
        """z(tau) = tau^size - 1 (domain.rs:229-231); tau: 12 Montgomery limbs"""
        return self._fop(ffi.OP_SUB, self._pow_size(tau), self._one()).reshape(LIMBS)

    def vanishing_polynomial(self):
        """the sparse polynomial X^size - 1 as [(degree, coefficient)] (domain.rs:222-225)"""
        one = self._one()
        zero = np.zeros((1, LIMBS), dtype=np.uint64)
        return [(0, self._fop(ffi.OP_SUB, zero, one).reshape(LIMBS)), (self.size_, one.reshape(LIMBS))]

    def evaluate_all_lagrange_coefficients(self, tau):
        """L_i(tau) for every i (domain.rs:183-220): the indicator of tau's position when tau lies in the
        domain, else (tau^n - 1) / n * g^i / (tau - g^i) - one element-wise inversion pass on the device
        instead of the reference's batch_inversion (the inverses are the same numbers)."""
        tau = np.ascontiguousarray(ffi.as_u64(tau).reshape(1, LIMBS))
        n = self.size_
        one = self._one()
        t_size = self._pow_size(tau)
        elems = self.elements().reshape(n, LIMBS)
        if np.array_equal(t_size, one):
            u = np.zeros((n, LIMBS), dtype=np.uint64)
            hit = np.flatnonzero((elems == tau).all(axis=1))
            if hit.size:
                u[hit[0]] = one[0]
            return u
        l0 = self._fop(ffi.OP_MUL, self._fop(ffi.OP_SUB, t_size, one), self.size_inv)
        u = DeviceVector(self.ctx, self.field, n, np.repeat(tau, n, axis=0))
        e = DeviceVector(self.ctx, self.field, n, elems)
        try:
            u.op(ffi.OP_SUB, e)          # tau - g^i
            u.op(ffi.OP_INV)
            u.op(ffi.OP_MUL, e)
            u.scale(l0)                  # * (tau^n - 1) / n
            return u.download()
        finally:
            u.free()
            e.free()

    def reindex_by_subdomain(self, other, index):
        """index of the `index`-th element in the ordering that lists the subdomain `other` first
        (domain.rs:261-284)"""
        assert self.size_ >= other.size_
        period = self.size_ // other.size_
        if index < other.size_:
            return index * period
        i = index - other.size_
        x = period - 1
        return i + (i // x) + 1

    # numpy arrays cannot be resized in place; the *_in_place forms return the resized vector
    fft_in_place = fft
    ifft_in_place = ifft
    coset_fft_in_place = coset_fft
    coset_ifft_in_place = coset_ifft


class MixedRadixDomain:
    """A multiplicative subgroup of size n = 2^a * m (m odd) - the mixed-radix domain of BASELINE
    config 4.  The reference snapshot implements radix-2 domains only; the interface mirrors
    EvaluationDomain (new -> None when no such subgroup exists, fft / ifft / coset_fft / coset_ifft
    on exactly `size` elements, natural order, omega = GENERATOR^((p-1)/n))."""

    def __init__(self, field, size, ctx):
        self.field, self.size_, self.ctx = field, size, ctx

    @staticmethod
    def new(field, size, ctx=None):
        lib = (ctx.lib if ctx else ffi.default_library())
        rc = lib.domain_check_mixed(field, size)
        if rc == ffi.ERR_DOMAIN:
            return None
        lib.check(rc)
        return MixedRadixDomain(field, size, ctx or default_context())

    def size(self):
        return self.size_

    def _run(self, v, mode):
        buf = np.zeros((self.size_, LIMBS), dtype=np.uint64)
        v = ffi.as_u64(v).reshape(-1, LIMBS)
        m = min(self.size_, v.shape[0])
        buf[:m] = v[:m]
        self.ctx.lib.check(self.ctx.lib.ntt_mixed(self.ctx.handle, self.field, ffi.ptr(buf), self.size_, mode))
        return buf

    def fft(self, coeffs):
        return self._run(coeffs, ffi.FFT)

    def ifft(self, evals):
        return self._run(evals, ffi.IFFT)

    def coset_fft(self, coeffs):
        return self._run(coeffs, ffi.COSET_FFT)

    def coset_ifft(self, evals):
        return self._run(evals, ffi.COSET_IFFT)


class DensePolynomial:
    """`DensePolynomial<F>` for the two 753-bit scalar fields as far as the transforms carry it
    (algebra/src/fft/polynomial/dense.rs): coefficient vector (n, 12) Montgomery, lowest degree first,
    no trailing zeros (`from_coefficients_vec`, dense.rs:64-74)."""

    def __init__(self, field, coeffs, ctx=None):
        c = ffi.as_u64(coeffs).reshape(-1, LIMBS)
        nz = np.flatnonzero(c.any(axis=1))
        self.field, self.ctx = field, ctx or default_context()
        self.coeffs = c[:nz[-1] + 1].copy() if nz.size else c[:0].copy()

    def is_zero(self):
        return self.coeffs.shape[0] == 0

    def degree(self):
        return max(self.coeffs.shape[0] - 1, 0)

    def evaluate_over_domain(self, domain):
        """dense.rs:237-248: the evaluations over the whole domain (zero-padded fft)"""
        return domain.fft(self.coeffs)

    def __mul__(self, other):
        """`Mul for &DensePolynomial` (dense.rs:342-357): evaluate both over a domain of
        len(a) + len(b) points, multiply point-wise, interpolate - chained on the device."""
        if self.is_zero() or other.is_zero():
            return DensePolynomial(self.field, np.zeros((0, LIMBS), dtype=np.uint64), self.ctx)
        domain = EvaluationDomain.new(self.field, self.coeffs.shape[0] + other.coeffs.shape[0], ctx=self.ctx)
        if domain is None:
            raise ValueError("field is not smooth enough to construct domain")
        a = DeviceVector(self.ctx, self.field, domain.size(), domain._resized(self.coeffs))
        b = DeviceVector(self.ctx, self.field, domain.size(), domain._resized(other.coeffs))
        try:
            a.ntt(ffi.FFT)
            b.ntt(ffi.FFT)
            a.op(ffi.OP_MUL, b)
            a.ntt(ffi.IFFT)
            return DensePolynomial(self.field, a.download(), self.ctx)
        finally:
            a.free()
            b.free()


class DeviceVector:
    """A vector of field elements resident in HBM, for chaining transforms the way
    R1CStoQAP::witness_map does (r1cs_to_qap.rs:121-161) without host round trips."""

    def __init__(self, ctx, field, n, host=None):
        self.ctx, self.field, self.n = ctx, field, n
        p = ctypes.c_void_p()
        ctx.lib.check(ctx.lib.dev_alloc(ctx.handle, n * 96, ctypes.byref(p)))
        self.ptr = p
        if host is not None:
            self.upload(host)

    def upload(self, host):
        host = ffi.as_u64(host).reshape(-1, LIMBS)
        assert host.shape[0] == self.n
        self.ctx.lib.check(self.ctx.lib.h2d(self.ctx.handle, self.ptr, ffi.ptr(host), self.n * 96))

    def download(self):
        out = np.empty((self.n, LIMBS), dtype=np.uint64)
        self.ctx.lib.check(self.ctx.lib.d2h(self.ctx.handle, ffi.ptr(out), self.ptr, self.n * 96))
        return out

    def ntt(self, mode):
        log_n = self.n.bit_length() - 1
        assert 1 << log_n == self.n
        self.ctx.lib.check(self.ctx.lib.ntt_dev(self.ctx.handle, self.field, self.ptr, log_n, mode))

    def op(self, op, other=None):
        self.ctx.lib.check(self.ctx.lib.vec_op_dev(self.ctx.handle, self.field, op, self.ptr,
                                                   other.ptr if other is not None else None, self.n))

    def scale(self, k_mont):
        k = ffi.as_u64(k_mont).reshape(LIMBS)
        self.ctx.lib.check(self.ctx.lib.vec_scale_dev(self.ctx.handle, self.field, self.ptr, ffi.ptr(k), self.n))

    def free(self):
        if getattr(self, "ptr", None):
            self.ctx.lib.dev_free(self.ctx.handle, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
