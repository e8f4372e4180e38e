#!/bin/bash
mkdir -p gpurun_out
for cpl in 0 1; do
echo "== coupling $cpl"
CWA_DIST_COUPLING=$cpl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 2>&1 | grep -E '^\{|Error|error|Traceback|assert' | head -8
done
for fp in 0 1; do
echo "== bench fused_pack $fp"
CWA_FUSED_PACK=$fp timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 2>gpurun_out/bench_n2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'], 'value', d['value']); print({k['kernel']: round(k['avg_us'],1) for k in d['roofline_kernels']})"
done
