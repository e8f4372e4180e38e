# compute-sanitizer over what the last part of round 2 added: the eight-lane shape of the clump force pass, tile culling of the all-pairs kernels,
# the folded 2-D launches + fused order/gather, all of it with dependent launches on
set -x
OUT=${1:-gpurun_out/sanitize_r2b}
mkdir -p $OUT
SEL='tests/test_gpu_nb_variants.py::test_queued_clump_targets_eight_lanes_each tests/test_gpu_nb_variants.py::test_variant_dense_cluster tests/test_gpu_sph3.py::test_allpairs_tile_culling_changes_nothing tests/test_gpu_sph2.py::test_one_frame_two_substeps tests/test_gpu_sph2.py::test_frame_graph_replays_the_same_frames'
for tool in memcheck racecheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $SEL -m gpu -x -q > $OUT/$tool.log 2>&1; echo "$tool rc=$?" >> $OUT/$tool.log
done
for f in memcheck racecheck synccheck; do tail -n 5 $OUT/$f.log; done
