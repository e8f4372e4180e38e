#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_nb_variants.py tests/test_gpu_sph3.py tests/test_gpu_golden.py -m gpu -x -q ) 2>&1 | tail -4
timeout 120 python tools/kernel_times.py 10 100 | grep -E "us/frame|heavy|density|force"
timeout 300 python tools/state_evolution.py 1000 3000 2>&1 | cut -c1-330
echo "== N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'], 'value', d['value']); print({k['kernel']: round(k['avg_us'],1) for k in d['roofline_kernels'][:6]})"
