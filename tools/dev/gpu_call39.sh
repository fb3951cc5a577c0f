#!/bin/bash
# six-slot mixed addition (12 warps per SM): targeted parity, MSM-only bench, ncu capture of k_bucket_acc
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "msm_small or accumulate_exceptions or domain_public or msm_vs_cpp" > gpurun_out/pytest_gpu39.log 2>&1
tail -3 gpurun_out/pytest_gpu39.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-groth16 --no-fft --no-config4 > gpurun_out/bench_plain19.log 2> gpurun_out/bench_plain19.err
tail -c 1500 gpurun_out/bench_plain19.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bucket_acc -s 3 -c 1 -o /tmp/acc19 python bench.py --steps 1 --warmup 3 --no-cpu --no-groth16 --no-fft --no-config4 > gpurun_out/ncu_acc19.log 2>&1
ncu -i /tmp/acc19.ncu-rep --page raw --csv > gpurun_out/r01_bucket_acc_2p22_v19.raw.csv 2>/dev/null
ncu -i /tmp/acc19.ncu-rep --page details --csv > gpurun_out/r01_bucket_acc_2p22_v19.details.csv 2>/dev/null
ls -la gpurun_out/*v19*
