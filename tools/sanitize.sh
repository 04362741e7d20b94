#!/usr/bin/env bash
# compute-sanitizer passes over the kernels that contain no tcgen05 / TMA instructions (the sanitizer's instrumentation of those is slow
# and was not needed so far) -- run on a GPU box:   gpurun --timeout 1500 -- 'bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1'
# memcheck: out-of-bounds / misaligned accesses;  racecheck: shared-memory hazards between warps;  synccheck: divergent barriers.
set -u
cd "$(dirname "$0")/.."
export PYTHONPATH="$PWD:$PWD/retrieval-augmented-diffusion-models_b200${PYTHONPATH:+:$PYTHONPATH}"
TESTS="tests/test_zz_rarm_gpu.py::test_cached_logits_match_reference_code tests/test_zz_rarm_gpu.py::test_guided_topk_draw_kernel_matches_oracle tests/test_zz_rarm_gpu.py::test_sampling_loop_token_by_token"
KNN_TESTS="tests/test_knn_gpu.py::test_query_normalisation_inside_the_library_is_bit_identical_to_numpy tests/test_knn_gpu.py::test_scann_shaped_api_and_gather tests/test_knn_gpu.py::test_shards_merge_to_the_unsharded_result tests/test_knn_gpu.py::test_many_exact_duplicates_tie_break"
rc=0
for tool in memcheck racecheck synccheck; do
    echo "=== compute-sanitizer --tool $tool ==="
    timeout 1200 compute-sanitizer --tool "$tool" --error-exitcode 9 --kernel-name kns=rarm_ \
        python -m pytest $TESTS -x -q -p no:cacheprovider 2>&1 | tail -25
    s=${PIPESTATUS[0]}; [ "$s" -ne 0 ] && rc=$s
    echo "=== compute-sanitizer --tool $tool: kNN normalise / select / merge / gather kernels ==="
    timeout 1200 compute-sanitizer --tool "$tool" --error-exitcode 9 --kernel-name kns=knn_normalize --kernel-name kns=knn_select --kernel-name kns=knn_merge --kernel-name kns=knn_gather \
        python -m pytest $KNN_TESTS -x -q -p no:cacheprovider 2>&1 | tail -12
    s=${PIPESTATUS[0]}; [ "$s" -ne 0 ] && rc=$s
done
exit $rc
