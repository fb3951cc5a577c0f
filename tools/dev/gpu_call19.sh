#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; cat gpurun_out/bench_v7.json | cut -c1-200; tail -3 gpurun_out/bench_v7.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_v7.json') if l.startswith('{')][-1]); print(d["value"], d["e2e"]["value"], d["roofline"]["phases_ms"], d["fft"]["ms"], d["groth16"]["value"], d["groth16"]["best_ms"], d["groth16"]["phases_s"])
PY
timeout 300 python tools/gpu_inv_probe.py 2>&1 | head -2
