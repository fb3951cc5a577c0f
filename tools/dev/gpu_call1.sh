#!/bin/bash
# debug call: full GPU test list, G2 MSM sweep, sanitizer on the smallest failing case
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
timeout 300 python tools/gpu_debug.py sweep > gpurun_out/debug_sweep.log 2>&1
timeout 200 python tools/gpu_debug.py smul > gpurun_out/debug_smul.log 2>&1
timeout 100 python tools/gpu_debug.py dump '' 1 5 3 gpurun_out/dump_g2.npz > gpurun_out/debug_dump.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python tools/gpu_debug.py one 1 5 3 > gpurun_out/sanitizer.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/debug_sweep.log | tail -30; tail -5 gpurun_out/sanitizer.log
