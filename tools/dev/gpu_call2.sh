#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_debug.py sweep > gpurun_out/debug_sweep2.log 2>&1
grep -c "ok True" gpurun_out/debug_sweep2.log; grep -c "ok False" gpurun_out/debug_sweep2.log
export PROBE_MSM_LOGS=18 PROBE_NTT_LOGS=20
timeout 300 python tools/gpu_probe.py msm ntt > gpurun_out/plain_probe.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_v0.csv python tools/gpu_probe.py msm ntt > gpurun_out/ncu1.log 2>&1
prof() {  # name regex skip what
  timeout 300 python tools/gpu_probe.py $4 > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o /tmp/$1 python tools/gpu_probe.py $4 > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details --csv > gpurun_out/$1.details.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
  gzip -f gpurun_out/$1.source.csv
}
prof r01_bucket_acc_v0 k_bucket_acc 1 msm
prof r01_reduce_v0 k_reduce_level 4 msm
prof r01_ntt_stage_v0 k_ntt_stage 30 ntt
ls -la gpurun_out; du -sh gpurun_out
