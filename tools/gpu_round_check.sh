# round check: GPU parity tests, smoke, both bench arms, ncu launch list and a full ncu capture of one steady-state frame
set -x
OUT=gpurun_out/${1:-check}
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 560 --launch-count 24 -f -o $OUT/c4_frame_full python tools/profile_c4.py 70 > $OUT/ncu_full.log 2>&1
tail -3 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; cut -c1-400 $OUT/bench_n1.json; cat $OUT/bench_n1.err | tail -5
