"""B200-native drop-in for ginger-lib's Groth16 prover hot path (MSM + radix-2 FFT over the
MNT4-753 / MNT6-753 cycle).  The compute lives in libg753.so (hand-written CUDA, sm_100a)
behind the C ABI of include/g753.h; this package is the host-side mirror of the reference's
operator interface.  Importing the compute classes requires the built library - there is no
CPU fallback."""
from . import ffi  # noqa: F401
from .algebra import (Bases, Context, DensePolynomial, DeviceVector, EvaluationDomain, FixedBaseMSM, MixedRadixDomain, VariableBaseMSM,  # noqa: F401
                      default_context)  # noqa: F401
