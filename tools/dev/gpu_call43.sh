#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py > gpurun_out/bench_v21.json 2> gpurun_out/bench_v21.err; tail -3 gpurun_out/bench_v21.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_v21.json') if l.startswith('{')][-1]); print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["groth16"]["value"]); print(d["cpu_baseline"])
PY
