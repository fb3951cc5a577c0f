#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench_groth16.py --log-n 16 --steps 2 > gpurun_out/groth16_2p16.json 2> gpurun_out/groth16_2p16.err; cat gpurun_out/groth16_2p16.json; tail -5 gpurun_out/groth16_2p16.err
timeout 900 python bench_groth16.py --log-n 20 --steps 3 > gpurun_out/groth16_2p20.json 2> gpurun_out/groth16_2p20.err; cat gpurun_out/groth16_2p20.json; tail -5 gpurun_out/groth16_2p20.err
timeout 600 python -m pytest tests/test_gpu_groth16.py -m gpu -q -x 2>&1 | tail -3
