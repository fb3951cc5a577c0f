#!/bin/bash
# A/B of the opt-in kernel forms against the default build on one B200 (round-1 result for G1 at 2^22:
# default 224.3 ms, rolled6 229.1 ms, rolled 246.3 ms - profiles/r01_ab45_rolled.jsonl; G2 not yet measured).
#   here (CPU):   tools/dev/build_variant.sh rolled  -DG753_ROLLED=1
#                 tools/dev/build_variant.sh rolled6 -DG753_ROLLED=1 -DG753_ACC6=1
#   then:         gpurun --timeout 600 -- bash tools/dev/r2_ab_rolled.sh
# Prints accumulate / reduce phase times of the 2^22 G1 MSM and the 2^20 G2 (Fq2, Fq3) MSMs per build.
mkdir -p gpurun_out
V=ginger-lib_b200/variants
out=gpurun_out/ab_rolled.jsonl; : > $out
run() { echo "{\"variant\": \"$1\"}" >> $out; G753_LIB=$2 timeout 300 python tools/gpu_msm_groups.py $3 $4 $5 >> $out 2>> gpurun_out/ab_rolled.err; }
for v in main rolled rolled6; do
  lib=""; [ $v != main ] && lib=$V/libg753_$v.so
  run $v "$lib" 22 0 0
  run $v "$lib" 20 1,3 0
done
cut -c1-400 $out
