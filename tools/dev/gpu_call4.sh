#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; cat gpurun_out/bench_v2.json; tail -5 gpurun_out/bench_v2.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_bench_v2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --no-fft > gpurun_out/bench_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_bucket_acc -s 2 -c 1 -o /tmp/acc22 python bench.py --steps 1 --warmup 3 --no-cpu --no-fft > gpurun_out/ncu_acc22.log 2>&1
ncu -i /tmp/acc22.ncu-rep --page raw --csv > gpurun_out/r01_bucket_acc_2p22.raw.csv 2>/dev/null
ncu -i /tmp/acc22.ncu-rep --page details --csv > gpurun_out/r01_bucket_acc_2p22.details.csv 2>/dev/null
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
