#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v12.json 2> gpurun_out/bench_v12.err; tail -3 gpurun_out/bench_v12.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_v12.json') if l.startswith('{')][-1]); print(d["value"], d["e2e"]["value"], d["roofline"]["phases_ms"], d["gpu_launches"], d["fft"]["ms"]); print(d["groth16"]["value"], d["groth16"]["phases_s"]); print(d["config4"]["msm"]["ms"], d["config4"]["mixed_radix_fft"]["ms"]); print(d["cpu_baseline"])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_v12.json 2>&1; tail -1 gpurun_out/bench_ref_v12.json | cut -c1-300
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu --no-groth16 --no-config4 > gpurun_out/bench_plain12.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_bench_v12.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-groth16 --no-config4 > gpurun_out/ncu_bench12.log 2>&1
wc -l gpurun_out/r01_launches_bench_v12.csv
