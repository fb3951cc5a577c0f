#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "ntt or mixed" > gpurun_out/pytest_gpu_ntt2.log 2>&1; tail -3 gpurun_out/pytest_gpu_ntt2.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; cat gpurun_out/bench_v5.json; tail -5 gpurun_out/bench_v5.err
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu --no-groth16 > gpurun_out/bench_plain5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_bench_v5.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-groth16 > gpurun_out/ncu_bench5.log 2>&1
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --no-groth16 > gpurun_out/bench_plain5b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -s 6 -c 1 -o /tmp/nttpass python bench.py --steps 1 --warmup 3 --no-cpu --no-groth16 > gpurun_out/ncu_nttpass.log 2>&1
ncu -i /tmp/nttpass.ncu-rep --page raw --csv > gpurun_out/r01_ntt_pass_2p22.raw.csv 2>/dev/null
ncu -i /tmp/nttpass.ncu-rep --page details --csv > gpurun_out/r01_ntt_pass_2p22.details.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_bucket_acc -s 3 -c 1 -o /tmp/acc5 python bench.py --steps 1 --warmup 3 --no-cpu --no-groth16 --no-fft > gpurun_out/ncu_acc5.log 2>&1
ncu -i /tmp/acc5.ncu-rep --page raw --csv > gpurun_out/r01_bucket_acc_2p22_v5.raw.csv 2>/dev/null
ncu -i /tmp/acc5.ncu-rep --page details --csv > gpurun_out/r01_bucket_acc_2p22_v5.details.csv 2>/dev/null
