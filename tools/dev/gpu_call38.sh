#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --no-groth16 --no-fft --no-config4 > gpurun_out/bench_plain18.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_bucket_acc -s 3 -c 1 -o /tmp/acc18 python bench.py --steps 1 --warmup 3 --no-cpu --no-groth16 --no-fft --no-config4 > gpurun_out/ncu_acc18.log 2>&1
ncu -i /tmp/acc18.ncu-rep --page raw --csv > gpurun_out/r01_bucket_acc_2p22_v18.raw.csv 2>/dev/null
ncu -i /tmp/acc18.ncu-rep --page details --csv > gpurun_out/r01_bucket_acc_2p22_v18.details.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_reduce_level -s 6 -c 1 -o /tmp/red18 python bench.py --steps 1 --warmup 3 --no-cpu --no-groth16 --no-fft --no-config4 > gpurun_out/ncu_red18.log 2>&1
ncu -i /tmp/red18.ncu-rep --page details --csv > gpurun_out/r01_reduce_level0_2p22_v18.details.csv 2>/dev/null
ls -la gpurun_out/*v18*
