set -x
timeout 1200 python -m pytest tests/test_unet_gpu.py tests/test_variants_gpu.py tests/test_zx_benchmarked_config_gpu.py tests/test_zy_ref_golden_gpu.py tests/test_mirror_gpu.py tests/test_vqdecoder_gpu.py tests/test_clip_gpu.py -m gpu -q -x -k "not knn" > gpurun_out/pytest_r2h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2h.log
tail -12 gpurun_out/pytest_r2h.log
grep -h "rel-L2" gpurun_out/pytest_r2h.log | head -20
timeout 300 python tools/ablate_forward.py 4 > gpurun_out/ablate_r2h.log 2>&1; cat gpurun_out/ablate_r2h.log
RDM_GN_TWO_KERNELS=1 timeout 100 python tools/profile_forward.py 4 30
timeout 300 python tools/profile_forward.py 4 > gpurun_out/profile_fwd_fp16_r2h.log 2>&1; head -32 gpurun_out/profile_fwd_fp16_r2h.log
