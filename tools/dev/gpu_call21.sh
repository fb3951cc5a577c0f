#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "kats or msm or generated" > gpurun_out/pytest_gpu_tp3.log 2>&1; tail -3 gpurun_out/pytest_gpu_tp3.log
timeout 600 python tools/gpu_msm_groups.py 20 3 8 > gpurun_out/msm_2p20_g3_tp4.jsonl 2>&1; cat gpurun_out/msm_2p20_g3_tp4.jsonl
timeout 1200 python bench.py --steps 3 --warmup 3 --no-groth16 > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err; tail -3 gpurun_out/bench_v8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_v8.json') if l.startswith('{')][-1]); print(d["value"], d["roofline"]["phases_ms"]); print(json.dumps(d["config4"]))
PY
