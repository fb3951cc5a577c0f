#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n${N}_v13.json 2> gpurun_out/bench_n${N}_v13.err; tail -3 gpurun_out/bench_n${N}_v13.err | cut -c1-300
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads([l for l in open('gpurun_out/bench_n%s_v13.json'%N) if l.startswith('{')][-1]); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["phases_ms"], d["groth16"]["value"] if d.get("groth16") else None, d["clocks"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 tools/gpu_ntt_sharded.py 22 > gpurun_out/ntt_sharded_fused_n${N}.json 2> gpurun_out/ntt_sharded_fused_n${N}.err; grep '^{' gpurun_out/ntt_sharded_fused_n${N}.json; tail -3 gpurun_out/ntt_sharded_fused_n${N}.err | cut -c1-300
