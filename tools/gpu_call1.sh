#!/bin/bash
# one GPU session: parity tests, tile-shape / pipeline sweeps, bench lines
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
{
for sc in 0 1 2 3; do echo "== scan_config $sc pipeline 0"; CWA_PIPELINE=0 CWA_SCAN_CONFIG=$sc timeout 120 python tools/kernel_times.py 10 100 | grep -E "us/frame|scan|clear|hash|insert"; done
for pl in 1 2 3; do echo "== pipeline $pl"; CWA_PIPELINE=$pl timeout 120 python tools/kernel_times.py 10 100; done
} > gpurun_out/sweep1.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 600 gpurun_out/bench_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "e2e", d["e2e"])
    for k in d["roofline_kernels"]: print(k["kernel"], round(k["avg_us"],1), round(k["frac"],3), k["launches"])
except Exception as e:
    print("bench parse failed", e)
PY
