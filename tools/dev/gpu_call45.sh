#!/bin/bash
# last GPU seconds of round 1: one timing of the opt-in build (rolled multiplier + six-slot addition), G1 2^22
mkdir -p gpurun_out
G753_LIB=ginger-lib_b200/variants/libg753_rolled6.so timeout 70 python tools/gpu_msm_groups.py 22 0 0 > gpurun_out/ab45_rolled6.jsonl 2> gpurun_out/ab45.err
cut -c1-400 gpurun_out/ab45_rolled6.jsonl; tail -2 gpurun_out/ab45.err
