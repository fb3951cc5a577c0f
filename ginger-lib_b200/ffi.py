"""ctypes binding of include/g753.h - the same symbols a Rust `-sys` crate would bind.

No arithmetic happens in Python: this module only marshals numpy buffers across the C ABI.
If the CUDA library is missing the import of the product fails loudly - there is no CPU path.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# G753_LIB: a differently built libg753.so (development A/B runs); the in-tree library otherwise
DEFAULT_LIB = os.environ.get("G753_LIB") or os.path.join(HERE, "libg753.so")

OK, ERR_BAD_ARG, ERR_CUDA, ERR_OOM, ERR_DOMAIN, ERR_NO_DEVICE = range(6)

MNT4_G1, MNT4_G2, MNT6_G1, MNT6_G2 = 0, 1, 2, 3
GROUP_K = {MNT4_G1: 1, MNT4_G2: 2, MNT6_G1: 1, MNT6_G2: 3}
FIELD_MNT4_FQ = FIELD_MNT6_FR = 0
FIELD_MNT6_FQ = FIELD_MNT4_FR = 1
FFT, IFFT, COSET_FFT, COSET_IFFT = 0, 1, 2, 3
OP_MUL, OP_ADD, OP_SUB, OP_SQR, OP_INV, OP_TO_MONT, OP_FROM_MONT = 0, 1, 2, 3, 5, 6, 7

# every symbol include/g753.h declares: (name, restype, argtypes)
_vp, _sz, _i, _u = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint
_pvp = ctypes.POINTER(ctypes.c_void_p)
SYMBOLS = [
    ("g753_device_count", _i, [ctypes.POINTER(_i)]),
    ("g753_ctx_create", _i, [_i, _pvp]),
    ("g753_ctx_destroy", _i, [_vp]),
    ("g753_ctx_set_stream", _i, [_vp, _vp]),
    ("g753_last_error", ctypes.c_char_p, []),
    ("g753_version", ctypes.c_char_p, []),
    ("g753_source_hash", ctypes.c_char_p, []),
    ("g753_bases_upload", _i, [_vp, _i, _vp, _vp, _sz, _pvp]),
    ("g753_bases_upload_wire", _i, [_vp, _i, _vp, _sz, _pvp]),
    ("g753_bases_free", _i, [_vp, _vp]),
    ("g753_bases_len", _sz, [_vp]),
    ("g753_bases_generate", _i, [_vp, _i, _vp, ctypes.c_uint64, _sz, _pvp]),
    ("g753_bases_precompute", _i, [_vp, _vp, _u]),
    ("g753_bases_update", _i, [_vp, _vp, _sz, _sz, _vp, _vp]),
    ("g753_bases_download", _i, [_vp, _vp, _sz, _sz, _vp]),
    ("g753_msm", _i, [_vp, _vp, _sz, _sz, _vp, _vp]),
    ("g753_msm_dev", _i, [_vp, _vp, _sz, _sz, _vp, _vp]),
    ("g753_msm_host", _i, [_vp, _i, _vp, _vp, _sz, _vp, _sz, _vp]),
    ("g753_points_sum_dev", _i, [_vp, _i, _vp, _sz, _vp]),
    ("g753_batch_normalize", _i, [_vp, _i, _vp, _sz, _vp, _vp]),
    ("g753_fixed_base_msm", _i, [_vp, _i, _vp, _vp, _sz, _vp, _vp]),
    ("g753_group_coord_limbs", _i, [_i]),
    ("g753_domain_check", _i, [_i, _u]),
    ("g753_domain_constant", _i, [_vp, _i, _u, _i, _vp]),
    ("g753_ntt", _i, [_vp, _i, _vp, _u, _i]),
    ("g753_ntt_dev", _i, [_vp, _i, _vp, _u, _i]),
    ("g753_domain_check_mixed", _i, [_i, ctypes.c_uint64]),
    ("g753_ntt_mixed", _i, [_vp, _i, _vp, ctypes.c_uint64, _i]),
    ("g753_ntt_mixed_dev", _i, [_vp, _i, _vp, ctypes.c_uint64, _i]),
    ("g753_ntt_shard_create", _i, [_vp, _i, _u, _u, _u, _pvp]),
    ("g753_ntt_shard_destroy", _i, [_vp, _vp]),
    ("g753_ntt_shard_shape", _i, [_vp, ctypes.POINTER(_sz), ctypes.POINTER(_sz), ctypes.POINTER(_sz), ctypes.POINTER(_sz)]),
    ("g753_ntt_shard_step1", _i, [_vp, _vp, _vp, _vp, _i]),
    ("g753_ntt_shard_step2", _i, [_vp, _vp, _vp, _vp, _i]),
    ("g753_ntt_shard_step1_fused", _i, [_vp, _vp, _vp, _vp, _i]),
    ("g753_ntt_shard_step2_local", _i, [_vp, _vp, _vp, _i]),
    ("g753_vec_op_dev", _i, [_vp, _i, _i, _vp, _vp, _sz]),
    ("g753_vec_scale_dev", _i, [_vp, _i, _vp, _vp, _sz]),
    ("g753_witness_map", _i, [_vp, _i, _vp, _vp, _vp, _u, _vp, _vp]),
    ("g753_witness_map_dev", _i, [_vp, _i, _vp, _vp, _vp, _u, _vp, _vp]),
    ("g753_witness_map_tail_dev", _i, [_vp, _i, _vp, _vp, _vp, _u, _vp, _vp]),
    ("g753_dev_alloc", _i, [_vp, _sz, _pvp]),
    ("g753_dev_free", _i, [_vp, _vp]),
    ("g753_h2d", _i, [_vp, _vp, _vp, _sz]),
    ("g753_d2h", _i, [_vp, _vp, _vp, _sz]),
    ("g753_d2d", _i, [_vp, _vp, _vp, _sz]),
    ("g753_sync", _i, [_vp]),
    ("g753_ctx_wait", _i, [_vp, _vp]),
    ("g753_stream", _vp, [_vp]),
    ("g753_field_op", _i, [_vp, _i, _i, _vp, _vp, _vp, _sz]),
    ("g753_point_op", _i, [_vp, _i, _i, _vp, _vp, _vp]),
    ("g753_ext_op", _i, [_vp, _i, _i, _i, _vp, _vp, _vp, _sz]),
    ("g753_coop_op", _i, [_vp, _i, _i, _u, _vp, _vp, _vp, _sz]),
    ("g753_mac_probe", _i, [_vp, _i, _i, _i, _i, ctypes.POINTER(ctypes.c_float)]),
    ("g753_debug_scratch", _i, [_vp, _vp, _sz, ctypes.POINTER(_sz)]),
    ("g753_launch_count", ctypes.c_uint64, [_vp]),
    ("g753_last_msm_phases", _i, [_vp, ctypes.POINTER(ctypes.c_float), _i]),
    ("g753_last_msm_plan", _i, [_vp, ctypes.POINTER(ctypes.c_uint)]),
]


class G753Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("g753 error %d: %s" % (code, msg))
        self.code = code


class Library:
    """A loaded libg753.so with typed entry points."""

    def __init__(self, path=None):
        path = path or DEFAULT_LIB
        if not os.path.exists(path):
            raise ImportError(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the product has no CPU fallback)" % path)
        self.path = path
        self.cdll = ctypes.CDLL(path)
        for name, res, args in SYMBOLS:
            fn = getattr(self.cdll, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
            setattr(self, name[5:], fn)

    def check(self, rc):
        if rc != OK:
            raise G753Error(rc, (self.last_error() or b"").decode("utf-8", "replace"))

    def device_count_safe(self):
        n = ctypes.c_int(0)
        rc = self.device_count(ctypes.byref(n))
        return n.value if rc == OK else 0


_default = None


def default_library():
    global _default
    if _default is None:
        _default = Library()
    return _default


def as_u64(a, shape_tail=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if shape_tail is not None and tuple(a.shape[-len(shape_tail):]) != tuple(shape_tail):
        raise ValueError("expected trailing shape %r, got %r" % (shape_tail, a.shape))
    return a


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
