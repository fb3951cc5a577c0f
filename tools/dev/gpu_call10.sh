#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/gpu_msm_groups.py 20 0,1,3 1,8 > gpurun_out/msm_groups_2p20.jsonl 2>&1; cat gpurun_out/msm_groups_2p20.jsonl
