#!/bin/bash
mkdir -p gpurun_out
CWA_BENCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | grep "trace rank" | cut -c1-330
