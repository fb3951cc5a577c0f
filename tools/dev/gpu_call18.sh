#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_groth16.py -m gpu -q -x -k "kats or msm or generated or proof" > gpurun_out/pytest_gpu_tp2.log 2>&1; tail -3 gpurun_out/pytest_gpu_tp2.log
timeout 600 python tools/gpu_msm_groups.py 20 1 1,8 > gpurun_out/msm_2p20_g2_tp2.jsonl 2>&1; cat gpurun_out/msm_2p20_g2_tp2.jsonl
