#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/gpu_ntt_sharded.py 22 > gpurun_out/ntt_sharded_n2.json 2> gpurun_out/ntt_sharded_n2.err; cat gpurun_out/ntt_sharded_n2.json; tail -5 gpurun_out/ntt_sharded_n2.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_v5.json 2> gpurun_out/bench_n2_v5.err; cat gpurun_out/bench_n2_v5.json; tail -5 gpurun_out/bench_n2_v5.err
