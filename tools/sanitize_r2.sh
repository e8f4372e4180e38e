# compute-sanitizer passes over the kernels added or changed in round 2 (run on the GPU box): memcheck, racecheck, synccheck
# row-mask density / force kernels (nb_config 8), clump paths, SM-balanced all-pairs kernels, 2-D kernels with lanes, slab count-ahead
set -x
OUT=${1:-gpurun_out/sanitize_r2}
mkdir -p $OUT
K='not 20000 and not 4194381'
SEL='tests/test_gpu_nb_variants.py::test_variant_matches_oracle tests/test_gpu_nb_variants.py::test_variant_dense_cluster tests/test_gpu_nb_variants.py::test_pipelined_frames_equal_the_plain_sequence tests/test_gpu_sph3.py::test_allpairs_kernel_variants tests/test_gpu_sph3.py::test_force_pass tests/test_gpu_sph2.py tests/test_oracle_kats.py tests/test_gpu_slab_group.py::test_count_ahead_across_the_exchange_changes_nothing'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $SEL -m gpu -x -q -k "$K and (8- or cfg8 or not variant_matches)" > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?" >> $OUT/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_nb_variants.py::test_variant_dense_cluster tests/test_gpu_sph3.py::test_allpairs_kernel_variants tests/test_gpu_sph2.py::test_c2_size_64k_particles_runs_and_conserves_count -m gpu -x -q -k "$K" > $OUT/racecheck.log 2>&1; echo "racecheck rc=$?" >> $OUT/racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_nb_variants.py::test_variant_dense_cluster tests/test_gpu_sph3.py::test_allpairs_kernel_variants tests/test_gpu_sph2.py::test_c2_size_64k_particles_runs_and_conserves_count -m gpu -x -q -k "$K" > $OUT/synccheck.log 2>&1; echo "synccheck rc=$?" >> $OUT/synccheck.log
for f in memcheck racecheck synccheck; do tail -n 4 $OUT/$f.log; done
