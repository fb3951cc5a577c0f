#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
export PROBE_MSM_LOGS=16,18,20,22
timeout 600 python tools/gpu_probe.py msm > gpurun_out/probe_msm_v1.log 2>&1; cat gpurun_out/probe_msm_v1.log
export PROBE_MSM_LOGS=18
prof() {  # name regex skip what
  timeout 300 python tools/gpu_probe.py $4 > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o /tmp/$1 python tools/gpu_probe.py $4 > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details --csv > gpurun_out/$1.details.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
  gzip -f gpurun_out/$1.source.csv
}
prof r01_bucket_acc_v1 k_bucket_acc 1 msm
prof r01_reduce_v1 k_reduce_level 4 msm
