#!/bin/bash
mkdir -p gpurun_out
for cfg in "64 192" "128 192" "128 384" "192 512" "256 768"; do set -- $cfg
echo "== nbr_k $1 extreme $2"; CWA_NBR_K=$1 CWA_EXTREME=$2 timeout 120 python tools/kernel_times.py 10 100 | grep -E "us/frame|density|force|heavy"
done
for cfg in "64 192" "128 384" "192 512"; do set -- $cfg
echo "== N=2 nbr_k $1 extreme $2"; CWA_NBR_K=$1 CWA_EXTREME=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'], 'value', d['value']); print({k['kernel']: round(k['avg_us'],1) for k in d['roofline_kernels'][:5]})"
done
