#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; cat gpurun_out/bench_v3.json; tail -5 gpurun_out/bench_v3.err
for c in 19 20 21; do
G753_MSM_C=$c timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-fft > gpurun_out/bench_v3_c$c.json 2>&1; cat gpurun_out/bench_v3_c$c.json
done
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-fft --log-n 19 > gpurun_out/bench_v3_2p19.json 2>&1; cat gpurun_out/bench_v3_2p19.json
