#!/bin/bash
# ncu --set full of the addition-tree rounds (the first `cnt` levels of one 2^ln MSM of group g, 8 key copies), exported
# on the box as details / raw / per-instruction source CSV, plus the launch list of one MSM:
#   gpurun --timeout 1200 -- bash tools/gpu_profile_tree.sh <tag> [group] [log_n] [cnt]
tag=${1:-r02}; g=${2:-0}; ln=${3:-22}; cnt=${4:-2}
out=gpurun_out
mkdir -p $out
python tools/gpu_msm_groups.py $ln $g 8 > $out/${tag}_tree_plain_g${g}.log 2>&1 || { tail -5 $out/${tag}_tree_plain_g${g}.log; exit 1; }
tail -1 $out/${tag}_tree_plain_g${g}.log
name=${tag}_tree_round_g${g}_2p${ln}
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_tree -c 40 --csv --log-file $out/${name}_launches.csv python tools/gpu_msm_groups.py $ln $g 8 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tree_round -s 0 -c $cnt -f -o $out/$name python tools/gpu_msm_groups.py $ln $g 8 > $out/${tag}_ncu_tree_g${g}.log 2>&1
ncu -i $out/$name.ncu-rep --page details --csv > $out/$name.ncu_details.csv 2>/dev/null
ncu -i $out/$name.ncu-rep --page raw --csv > $out/$name.ncu_raw.csv 2>/dev/null
ncu -i $out/$name.ncu-rep --page source --print-source sass --csv > $out/$name.ncu_source_sass.csv 2>/dev/null
rm -f $out/$name.ncu-rep
ls -la $out/${name}*
