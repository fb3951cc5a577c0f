"""GM17 prover path and DensePolynomial products on the GPU: the host mirrors over
the C ABI against the Python oracle at tiny sizes (bit-exact proofs), and the SAP witness map at 2^12
against the C++ restatement's transforms (every limb).  Named to run after the other GPU tests."""
import importlib

import numpy as np
import pytest

from oracle import g753 as O
from test_gm17_emul import check_instance
from util753 import FIELDS, G, array_field, ffi, field_array, ints_to_array

gm17 = importlib.import_module("ginger-lib_b200.gm17")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = G.Context(0)
    yield c
    c.close()


def test_gm17_proof_tiny_vs_oracle(ctx):
    F = O.MNT4_FR
    check_instance(ctx, 0x1707, 8, 2, 3, 2, 5, 7, 0x1234567 << 600)
    check_instance(ctx, 0x1708, 16, 3, 4, 4, 0, 0, F.p - 3)
    check_instance(ctx, 0x1709, 4, 1, 3, 1, F.p - 1, 3, 0)
    check_instance(ctx, 0x170a, 64, 3, 40, 20, 11, 13, 12345)
    check_instance(ctx, 0x1761, 8, 2, 3, 2, 1, 2, O.MNT6_FR.p - 5, engine="mnt6")    # G2 over Fq3


def test_sap_witness_map_2e12_vs_cpp_restatement(ctx):
    """every limb of h against the C++ restatement's ifft / coset_fft / coset_ifft chain"""
    import bench
    from oracle import ref753
    F = O.MNT4_FR
    field = ffi.FIELD_MNT4_FR
    n = 1 << 12
    a, c = (bench.random_scalars(n, 0x200 + i) for i in range(2))
    for v in (a, c):
        v[:, 11] &= np.uint64(0xFFFF)      # < p: a valid Montgomery representation
    d1, d2 = 9, 8
    h = gm17.witness_map(ctx, field, a, c, d1, d2)
    ca = ref753.fft(field, a, ffi.IFFT)
    fa = ref753.fft(field, ca, ffi.COSET_FFT)
    fc = ref753.fft(field, ref753.fft(field, c, ffi.IFFT), ffi.COSET_FFT)
    aa = ref753.field_op(field, ffi.OP_SUB, ref753.field_op(field, ffi.OP_MUL, fa, fa), fc)
    zinv = F.to_mont(pow((pow(F.generator, n, F.p) - 1) % F.p, -1, F.p))
    aa = ref753.field_op(field, ffi.OP_MUL, aa, np.tile(ints_to_array([zinv]), (n, 1)))
    q = ref753.fft(field, aa, ffi.COSET_IFFT)
    want = ref753.field_op(field, ffi.OP_MUL, ca, np.tile(ints_to_array([F.to_mont(2 * d1)]), (n, 1)))
    want = np.concatenate([want, ints_to_array([F.to_mont(d1 * d1 % F.p)])])
    want[0] = ints_to_array([F.to_mont((F.from_mont(int(sum(int(want[0, i]) << (64 * i) for i in range(12)))) - d2 - d1 * d1) % F.p)])[0]
    want[:n - 1] = ref753.field_op(field, ffi.OP_ADD, want[:n - 1], q[:n - 1])
    assert np.array_equal(h, want)


@pytest.mark.parametrize("field", sorted(FIELDS))
def test_dense_polynomial_mul(ctx, field):
    F = FIELDS[field]
    rng = O.SplitMix64(0xDE + field)
    for la, lb in ((1, 1), (3, 5), (40, 25)):
        a = [O.random_field_element(rng, F) for _ in range(la)]
        b = [O.random_field_element(rng, F) for _ in range(lb)]
        want = [0] * (la + lb - 1)
        for i, x in enumerate(a):
            for j, y in enumerate(b):
                want[i + j] = (want[i + j] + x * y) % F.p
        pa = G.DensePolynomial(field, field_array(F, a + [0]), ctx)
        pb = G.DensePolynomial(field, field_array(F, b), ctx)
        assert array_field(F, (pa * pb).coeffs) == want


@pytest.mark.parametrize("field", sorted(FIELDS))
def test_domain_elements_vanishing_lagrange(ctx, field):
    """EvaluationDomain::{elements, evaluate_vanishing_polynomial, evaluate_all_lagrange_coefficients,
    reindex_by_subdomain} (domain.rs:183-284) on the device"""
    from test_pipeline_emul import check_domain_methods
    check_domain_methods(ctx, field, (1, 2, 64, 1 << 10))
