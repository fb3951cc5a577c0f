#!/bin/bash
mkdir -p gpurun_out
timeout 110 python -m pytest tests/test_gpu_zz_gm17.py -x -q -m gpu > gpurun_out/pytest_gpu44.log 2>&1
tail -5 gpurun_out/pytest_gpu44.log
