mkdir -p gpurun_out/r1u
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1u/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1u/pytest.log
tail -4 gpurun_out/r1u/pytest.log
python tools/state_evolution.py 10 60 200 1000 3000 5000 > gpurun_out/r1u/state_evolution.log 2>&1; cat gpurun_out/r1u/state_evolution.log
