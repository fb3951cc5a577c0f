#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/gpu_ntt_sharded.py 22 > gpurun_out/ntt_sharded_fused_n${N}.json 2> gpurun_out/ntt_sharded_fused_n${N}.err; grep '^{' gpurun_out/ntt_sharded_fused_n${N}.json; tail -5 gpurun_out/ntt_sharded_fused_n${N}.err | cut -c1-400
