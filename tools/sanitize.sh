# compute-sanitizer passes over the kernels added or changed last (run on the GPU box): memcheck, racecheck, synccheck
set -x
SEL='tests/test_gpu_stencil1d.py tests/test_gpu_state.py tests/test_gpu_grid_scan.py::test_scan_every_tile_shape tests/test_gpu_nb_variants.py::test_pipelined_frames_equal_the_plain_sequence tests/test_gpu_nb_variants.py::test_variant_dense_cluster'
K='not 20000 and not 4194381'
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $SEL -m gpu -x -q -k "$K" 2>&1 | tail -4
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_stencil1d.py tests/test_gpu_grid_scan.py::test_scan_every_tile_shape tests/test_gpu_nb_variants.py::test_variant_dense_cluster tests/test_gpu_wave.py -m gpu -x -q -k "$K" 2>&1 | tail -4
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_nb_variants.py::test_pipelined_frames_equal_the_plain_sequence tests/test_gpu_grid_scan.py::test_scan_every_tile_shape -m gpu -x -q -k "$K" 2>&1 | tail -4
