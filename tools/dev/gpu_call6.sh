#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "ntt" > gpurun_out/pytest_gpu_ntt.log 2>&1; tail -5 gpurun_out/pytest_gpu_ntt.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err; cat gpurun_out/bench_v4.json; tail -5 gpurun_out/bench_v4.err
timeout 600 python tools/gpu_probe.py ntt > gpurun_out/probe_ntt_v1.log 2>&1; tail -12 gpurun_out/probe_ntt_v1.log
