#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 1200 python tools/gpu_msm_groups.py 20 0,1,3 1,8 > gpurun_out/msm_groups_2p20_v2.jsonl 2>&1; cat gpurun_out/msm_groups_2p20_v2.jsonl
for c in 15 16; do
G753_MSM_C=$c timeout 300 python tools/gpu_msm_groups.py 20 0 1 2>&1 | tail -1
done
