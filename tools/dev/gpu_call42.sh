#!/bin/bash
# final round-1 validation: GPU parity suite, smoke, default bench, reference arm, launch list, ncu of the Fq2 accumulation
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v20.json 2> gpurun_out/bench_v20.err; tail -3 gpurun_out/bench_v20.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_v20.json') if l.startswith('{')][-1]); print(d["value"], d["e2e"]["value"], d["roofline"]["phases_ms"], d["roofline"]["frac"], d["gpu_launches"], d["fft"]["ms"]); print(d["groth16"]["value"], d["groth16"]["phases_s"]); print(d["config4"]["msm"]["ms"], d["config4"]["mixed_radix_fft"]["ms"]); print(d["cpu_baseline"]["value"], d["clocks"])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_v20.json 2> gpurun_out/bench_reference_v20.err; tail -c 600 gpurun_out/bench_reference_v20.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_bench_v20.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-groth16 --no-fft --no-config4 > gpurun_out/ncu_launches20.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_bucket_acc -s 2 -c 1 -o /tmp/accg2 python tools/gpu_msm_groups.py 20 1 0 > gpurun_out/ncu_accg2.log 2>&1
ncu -i /tmp/accg2.ncu-rep --page raw --csv > gpurun_out/r01_bucket_acc_fq2_2p20_v20.raw.csv 2>/dev/null
ncu -i /tmp/accg2.ncu-rep --page details --csv > gpurun_out/r01_bucket_acc_fq2_2p20_v20.details.csv 2>/dev/null
ls -la gpurun_out/*v20*
