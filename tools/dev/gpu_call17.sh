#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-groth16 --no-fft --copies 16 > gpurun_out/bench_v6_c16.json 2>&1; cat gpurun_out/bench_v6_c16.json | cut -c1-300; 
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_v6_c16.json') if l.startswith('{')][-1]); print(d["value"], d["roofline"]["phases_ms"], d["key_precompute_s"])
PY
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-groth16 --no-fft --copies 32 > gpurun_out/bench_v6_c32.json 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_v6_c32.json') if l.startswith('{')][-1]); print(d["value"], d["roofline"]["phases_ms"], d["key_precompute_s"])
PY
timeout 600 python tools/gpu_msm_groups.py 20 0,1 16 > gpurun_out/msm_2p20_c16.jsonl 2>&1; cat gpurun_out/msm_2p20_c16.jsonl
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
