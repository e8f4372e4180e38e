#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_sph3.py tests/test_gpu_nb_variants.py -m gpu -x -q ) > gpurun_out/pytest_gpu2.log 2>&1
tail -4 gpurun_out/pytest_gpu2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 400 gpurun_out/bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
    print("N=2 ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"])
    for k in d["roofline_kernels"]: print("  ", k["kernel"], round(k["avg_us"],1), k["launches"])
except Exception as e:
    print("bench parse failed", e)
PY
