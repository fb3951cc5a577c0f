#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py --steps 3 --warmup 3 --no-config4 > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.err; tail -3 gpurun_out/bench_v9.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_v9.json') if l.startswith('{')][-1]); print(d["value"], d["e2e"]["value"], d["roofline"]["phases_ms"], d["groth16"]["value"], d["groth16"]["best_ms"], d["groth16"]["phases_s"])
PY
timeout 300 python tools/gpu_msm_groups.py 16 0 1,8,32 2>&1 | tail -3
timeout 300 python tools/gpu_msm_groups.py 19 0 8 2>&1 | tail -1
