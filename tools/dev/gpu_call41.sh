#!/bin/bash
# A/B: shared-memory carveout on/off, squarings through the product body; parity of the lazy Fq2 tower
mkdir -p gpurun_out
V=ginger-lib_b200/variants
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "msm_small" > gpurun_out/pytest_gpu41.log 2>&1
tail -3 gpurun_out/pytest_gpu41.log
out=gpurun_out/ab41.jsonl; : > $out
run() { echo "{\"variant\": \"$1\"}" >> $out; G753_LIB=$2 timeout 200 python tools/gpu_msm_groups.py $3 $4 $5 >> $out 2>> gpurun_out/ab41.err; }
run main "" 22 0 0
run nocarve $V/libg753_nocarve.so 22 0 0
run sqrmul $V/libg753_sqrmul.so 22 0 0
run main "" 20 1 0
run nocarve $V/libg753_nocarve.so 20 1 0
cat $out | cut -c1-400
