#!/bin/bash
mkdir -p gpurun_out
for sl in 2 3 4; do echo "== e2e slots $sl"; CWA_E2E_SLOTS=$sl timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_unpipelined'])"; done
for sc in 0 1 2 3; do echo "== scan_config $sc"; CWA_SCAN_CONFIG=$sc timeout 120 python tools/kernel_times.py 10 100 | grep -E "us/frame|scan"; done
