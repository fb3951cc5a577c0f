#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --no-config4 --no-groth16 --no-cpu --no-fft > gpurun_out/bench_v17.json 2> gpurun_out/bench_v17.err; tail -3 gpurun_out/bench_v17.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_v17.json') if l.startswith('{')][-1]); print(d["value"], d["e2e"]["value"], d["roofline"]["phases_ms"], d["gpu_launches"])
PY
timeout 300 python tools/gpu_msm_groups.py 20 0,1 0 2>&1 | tail -2
