#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_msm_groups.py 19 0 64 2>&1 | tail -1
timeout 300 python tools/gpu_msm_groups.py 20 0 32 2>&1 | tail -1
timeout 300 python tools/gpu_msm_groups.py 16 0 64 2>&1 | tail -1
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v10.json 2> gpurun_out/bench_v10.err; tail -3 gpurun_out/bench_v10.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_v10.json') if l.startswith('{')][-1]); print(d["value"], d["e2e"]["value"], d["roofline"]["phases_ms"], d["groth16"]["value"], d["groth16"]["best_ms"], d["groth16"]["phases_s"], d["groth16"]["config"]["key"]); print(d["config4"]["msm"]["ms"], d["config4"]["msm"]["phases_ms"], d["config4"]["msm"]["workload"])
PY
