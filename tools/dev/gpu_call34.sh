#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v16.json 2> gpurun_out/bench_v16.err; tail -3 gpurun_out/bench_v16.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_v16.json') if l.startswith('{')][-1]); print(d["value"], d["e2e"]["value"], d["roofline"]["phases_ms"], d["roofline"]["frac"], d["gpu_launches"], d["fft"]["ms"]); print(d["groth16"]["value"], d["groth16"]["phases_s"]); print(d["config4"]["msm"]["ms"], d["config4"]["mixed_radix_fft"]["ms"]); print(d["cpu_baseline"]["value"], d["clocks"])
PY
