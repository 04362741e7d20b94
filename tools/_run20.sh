set -x
timeout 900 python -m pytest tests/test_variants_gpu.py tests/test_vqdecoder_gpu.py tests/test_zx_benchmarked_config_gpu.py tests/test_unet_gpu.py -m gpu -q -x -k "not knn" 2>&1 | tail -6
for i in 1 2; do timeout 100 python tools/profile_forward.py 4 30; RDM_TC_2SM=0 timeout 100 python tools/profile_forward.py 4 30; done
timeout 200 python tools/profile_forward_r.py 4 10; RDM_TC_2SM=0 timeout 200 python tools/profile_forward_r.py 4 10; RDM_TC_2SM=2 timeout 200 python tools/profile_forward_r.py 4 10
