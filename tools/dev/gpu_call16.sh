#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_inv_probe.py > gpurun_out/inv_probe.jsonl 2>&1; cat gpurun_out/inv_probe.jsonl
G753_MSM_AFFINE=1 timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "msm or kats or generated" > gpurun_out/pytest_gpu_affine.log 2>&1; tail -3 gpurun_out/pytest_gpu_affine.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-groth16 --no-fft > gpurun_out/bench_v6_affine.json 2> gpurun_out/bench_v6.err; cat gpurun_out/bench_v6_affine.json; tail -3 gpurun_out/bench_v6.err
G753_MSM_AFFINE=0 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-groth16 --no-fft > gpurun_out/bench_v6_xyzz.json 2>&1; cat gpurun_out/bench_v6_xyzz.json
G753_MSM_AFFINE=1 timeout 600 python tools/gpu_msm_groups.py 20 0 8 > gpurun_out/msm_2p20_affine.jsonl 2>&1; cat gpurun_out/msm_2p20_affine.jsonl
