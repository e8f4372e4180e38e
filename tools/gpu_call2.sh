mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests/test_gpu_nb_variants.py tests/test_gpu_sph3.py -m gpu -x -q > gpurun_out/r2b/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b/pytest.log
tail -3 gpurun_out/r2b/pytest.log
python tools/state_evolution.py 10 200 1000 5000 2>&1 | tail -4
