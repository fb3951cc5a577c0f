#!/bin/bash
# A/B of the accumulation kernels: G1 six-slot interpreter (main, nc64) and the Fq2 lazy-reduction variants
mkdir -p gpurun_out
V=ginger-lib_b200/variants
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "msm_small or accumulate_exceptions" > gpurun_out/pytest_gpu40.log 2>&1
tail -3 gpurun_out/pytest_gpu40.log
out=gpurun_out/ab40.jsonl; : > $out
run() { echo "{\"variant\": \"$1\"}" >> $out; G753_LIB=$2 timeout 200 python tools/gpu_msm_groups.py $3 $4 $5 >> $out 2>> gpurun_out/ab40.err; }
run main "" 22 0 0
run nc64 $V/libg753_nc64.so 22 0 0
run main "" 20 1 0
run lazy0 $V/libg753_lazy0.so 20 1 0
run lazy2 $V/libg753_lazy2.so 20 1 0
cat $out | cut -c1-400
