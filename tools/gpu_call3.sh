#!/bin/bash
mkdir -p gpurun_out
CWA_BENCH_PHASES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2_phases.log 2>&1
grep -E "phases|kernels rank|early|metric" gpurun_out/bench_n2_phases.log | cut -c1-1500
