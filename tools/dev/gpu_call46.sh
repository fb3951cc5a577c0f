#!/bin/bash
# last GPU seconds of round 1: one timing of the opt-in build (rolled multiplier, eight-slot addition), G1 2^22
mkdir -p gpurun_out
G753_LIB=ginger-lib_b200/variants/libg753_rolled.so timeout 70 python tools/gpu_msm_groups.py 22 0 0 > gpurun_out/ab46_rolled.jsonl 2> gpurun_out/ab46.err
cut -c1-400 gpurun_out/ab46_rolled.jsonl; tail -2 gpurun_out/ab46.err
