#!/bin/bash
mkdir -p gpurun_out
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 2>gpurun_out/bench_n2.err | tee gpurun_out/bench_n2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step']); print({k['kernel']: round(k['avg_us'],1) for k in d['roofline_kernels']})"
tail -c 300 gpurun_out/bench_n2.err | grep -v OMP
echo "== trace"; CWA_BENCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | grep "trace rank" | awk 'NR%4<2' | cut -c1-260 | tail -12
