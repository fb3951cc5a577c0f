#!/bin/bash
mkdir -p gpurun_out
export G753_MSM_C=16
timeout 300 python tools/gpu_msm_groups.py 20 0 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_bucket_acc -s 2 -c 1 -o /tmp/acc20 python tools/gpu_msm_groups.py 20 0 1 > gpurun_out/ncu_acc20.log 2>&1
ncu -i /tmp/acc20.ncu-rep --page raw --csv > gpurun_out/r01_bucket_acc_2p20_c16.raw.csv 2>/dev/null
ncu -i /tmp/acc20.ncu-rep --page details --csv > gpurun_out/r01_bucket_acc_2p20_c16.details.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_msm20_c16.csv python tools/gpu_msm_groups.py 20 0 1 > /dev/null 2>&1
tail -30 gpurun_out/r01_launches_msm20_c16.csv | cut -c1-200
