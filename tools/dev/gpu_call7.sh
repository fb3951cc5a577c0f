#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_groth16.py -m gpu -q -x > gpurun_out/pytest_gpu_groth16.log 2>&1; tail -15 gpurun_out/pytest_gpu_groth16.log
