set -x
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_zx_benchmarked_config_gpu.py tests/test_zy_ref_golden_gpu.py tests/test_clip_gpu.py -m gpu -q -x -k "not knn" 2>&1 | tail -4
timeout 100 python tools/profile_forward.py 4 30
RDM_TC_GEGLU_TRANSPOSED=1 timeout 100 python tools/profile_forward.py 4 30
timeout 200 python tools/profile_forward.py 4 2>&1 | grep "act=2\|act=4"
