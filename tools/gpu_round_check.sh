set -x
mkdir -p gpurun_out/r1e
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1e/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1e/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r1e/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/r1e/bench_n1.json 2> gpurun_out/r1e/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r1e/bench_ref.json 2> gpurun_out/r1e/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1e/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1e/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 110 --launch-count 11 -f -o gpurun_out/r1e/c4_frame_full python tools/profile_c4.py 12 > gpurun_out/r1e/ncu_full.log 2>&1
tail -3 gpurun_out/r1e/pytest_gpu.log; cat gpurun_out/r1e/smoke.log | tail -2; cat gpurun_out/r1e/bench_n1.json | cut -c1-600
