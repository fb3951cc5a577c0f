#!/bin/bash
mkdir -p gpurun_out
for c in 13 14 15 16 17; do
G753_MSM_C=$c timeout 300 python tools/gpu_msm_groups.py 20 0 1 2>&1 | tail -1
done > gpurun_out/msm_c_sweep_2p20.jsonl
cat gpurun_out/msm_c_sweep_2p20.jsonl
