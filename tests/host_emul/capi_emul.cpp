// TEST-ONLY host-emulation build of the whole C ABI (see csrc/device.cuh): every kernel
// launch becomes a sequential loop over (block, thread) on the CPU.  Used by
// tests/test_pipeline_emul.py to check digit extraction, sorting, bucket bookkeeping,
// reduction levels and NTT indexing against the oracle in the CPU-only tier.  Never shipped,
// never loaded by the product.
#define G753_HOST_EMUL 1
#include "../../ginger-lib_b200/csrc/capi.cu"
