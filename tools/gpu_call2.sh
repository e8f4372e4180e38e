mkdir -p gpurun_out/r1m
timeout 900 python -m pytest tests/test_gpu_nb_variants.py tests/test_gpu_sph3.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/r1m/pytest_nb.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1m/pytest_nb.log
tail -8 gpurun_out/r1m/pytest_nb.log
timeout 1200 bash tools/sweep_rows.sh > gpurun_out/r1m/sweep_rows.log 2>&1
cat gpurun_out/r1m/sweep_rows.log
CWA_BENCH_GRID=h_y30 CWA_NB_CONFIG=10 timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 600 --launch-count 24 -f -o gpurun_out/r1m/c4_list_h_full python tools/profile_c4.py 62 > gpurun_out/r1m/ncu_full_h.log 2>&1
tail -3 gpurun_out/r1m/ncu_full_h.log
