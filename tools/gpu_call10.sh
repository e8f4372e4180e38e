#!/bin/bash
mkdir -p gpurun_out
N=$1
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 2>gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step']); print({k['kernel']: round(k['avg_us'],1) for k in d['roofline_kernels']})"
grep -v "OMP\|\*\*\*" gpurun_out/bench_n$N.err | tail -c 1500
echo "== reference arm N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 4 --warmup 1 2>/dev/null | cut -c1-400
