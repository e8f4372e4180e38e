mkdir -p gpurun_out/r2c
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c/pytest.log
tail -6 gpurun_out/r2c/pytest.log
python tools/state_evolution.py 10 200 1000 5000 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 720 --launch-count 13 -f -o gpurun_out/r2c/frame_full python tools/profile_c4.py 62 > gpurun_out/r2c/ncu.log 2>&1; tail -2 gpurun_out/r2c/ncu.log
