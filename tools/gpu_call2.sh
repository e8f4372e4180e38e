#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
{
for tr in 0 1; do for pl in 0 3; do echo "== transpose $tr pipeline $pl"; CWA_WAVE_TRANSPOSE=$tr CWA_PIPELINE=$pl timeout 120 python tools/kernel_times.py 10 100; done; done
} > gpurun_out/sweep2.log 2>&1
cat gpurun_out/sweep2.log | grep -E "==|us/frame|integrate|scan|density|other"
