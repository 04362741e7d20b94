set -x
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -8
timeout 100 python tools/profile_forward.py 4 30
RDM_TC_NO_HALO=1 timeout 100 python tools/profile_forward.py 4 30
timeout 200 python tools/profile_forward.py 4 2>&1 | grep "ks=3" | head -24
timeout 900 python -m pytest tests/test_zx_benchmarked_config_gpu.py tests/test_zy_ref_golden_gpu.py tests/test_variants_gpu.py tests/test_vqdecoder_gpu.py tests/test_mirror_gpu.py -m gpu -q -x -k "not knn" 2>&1 | tail -5
