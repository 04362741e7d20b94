set -x
timeout 100 python tools/profile_forward.py 4 30
RDM_GN_FUSED_MAX_HW=256 timeout 100 python tools/profile_forward.py 4 30
RDM_GN_FUSED_MAX_HW=16 timeout 100 python tools/profile_forward.py 4 30
timeout 900 python -m pytest tests/test_variants_gpu.py tests/test_unet_gpu.py tests/test_zy_ref_golden_gpu.py tests/test_mirror_gpu.py -m gpu -q -x -k "not knn" 2>&1 | tail -4
