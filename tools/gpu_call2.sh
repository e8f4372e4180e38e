mkdir -p gpurun_out/r1z
echo "[default h16 t128 s3 c4]"; python tools/wave_bench.py
for v in build/variants/wave_*.so; do echo "[$v]"; CWA_LIB_PATH=/root/repo/$v python tools/wave_bench.py; done
timeout 300 python -m pytest tests/test_gpu_wave.py -m gpu -x -q 2>&1 | tail -2
