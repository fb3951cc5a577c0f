#!/bin/bash
# The round's ncu evidence, one call on ONE GPU:  gpurun --timeout 1500 -- bash tools/gpu_profile.sh <tag>
# Writes gpurun_out/<tag>_*: the launch list of the MSM bench (gpu__time_duration.sum: kernel SHARES of a step) and
# `ncu --set full` captures of the dominant kernels (G1 / Fq2 / Fq3 bucket accumulation, the first reduction level,
# the batched NTT pass of the sharded transform), exported on the box as details / raw CSV (the .ncu-rep files
# together exceed what gpurun copies back, so only the CSV summaries return; they are committed under profiles/).
# ncu runs every command once without profiling first (its own rule), so the commands are kept short.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
export_rep() {  # <name>: .ncu-rep -> details + raw csv, then drop the report
  ncu -i $out/$1.ncu-rep --page details --csv > $out/$1.ncu_details.csv 2>/dev/null
  ncu -i $out/$1.ncu-rep --page raw --csv > $out/$1.ncu_raw.csv 2>/dev/null
  rm -f $out/$1.ncu-rep
}
MSM="python bench.py --steps 2 --warmup 3 --no-cpu --no-fft --no-configs --no-groth16"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_bench.csv $MSM > $out/${tag}_launches.log 2>&1
cp gpurun_out/ncu_plain_run.log $out/${tag}_launches_plain_run.log 2>/dev/null
ncu --set full --clock-control none --import-source on -k "regex:k_bucket_acc|k_reduce_level" -s 2 -c 2 -f -o $out/${tag}_g1_acc_reduce_2p22 python tools/gpu_msm_groups.py 22 0 0 > $out/${tag}_ncu_g1.log 2>&1
export_rep ${tag}_g1_acc_reduce_2p22
for g in 1 3; do
  ncu --set full --clock-control none --import-source on -k regex:k_bucket_acc -s 1 -c 1 -f -o $out/${tag}_bucket_acc_g${g}_2p20 python tools/gpu_msm_groups.py 20 $g 0 > $out/${tag}_ncu_g${g}.log 2>&1
  export_rep ${tag}_bucket_acc_g${g}_2p20
done
ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -s 12 -c 1 -f -o $out/${tag}_ntt_pass_batched_2p22 python tools/gpu_ntt_batched_single.py 22 > $out/${tag}_ncu_ntt.log 2>&1
export_rep ${tag}_ntt_pass_batched_2p22
du -sh $out; ls -la $out/${tag}_* | head -30
