set -x
timeout 1200 python -m pytest tests/test_unet_gpu.py tests/test_variants_gpu.py tests/test_zx_benchmarked_config_gpu.py tests/test_zy_ref_golden_gpu.py tests/test_mirror_gpu.py tests/test_script_flow_gpu.py tests/test_vqdecoder_gpu.py -m gpu -q -x -k "not knn" > gpurun_out/pytest_r2j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2j.log
tail -12 gpurun_out/pytest_r2j.log
grep -h "rel-L2" gpurun_out/pytest_r2j.log | head -30
timeout 300 python tools/ablate_forward.py 4 > gpurun_out/ablate_r2j.log 2>&1; cat gpurun_out/ablate_r2j.log
RDM_GN_NO_EPI_STATS=1 timeout 100 python tools/profile_forward.py 4 30
