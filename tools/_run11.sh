set -x
timeout 1200 python -m pytest tests/test_unet_gpu.py tests/test_variants_gpu.py tests/test_zx_benchmarked_config_gpu.py tests/test_zy_ref_golden_gpu.py tests/test_mirror_gpu.py tests/test_script_flow_gpu.py -m gpu -q -x -k "not knn" > gpurun_out/pytest_r2k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2k.log
tail -6 gpurun_out/pytest_r2k.log
grep -h "rel-L2" gpurun_out/pytest_r2k.log | head -30
timeout 100 python tools/profile_forward.py 4 30
RDM_GN_NO_EPI_STATS=1 timeout 100 python tools/profile_forward.py 4 30
timeout 300 python tools/ablate_forward.py 4 > gpurun_out/ablate_r2k.log 2>&1; cat gpurun_out/ablate_r2k.log
