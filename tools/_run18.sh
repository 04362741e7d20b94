set -x
for i in 1 2; do
timeout 100 python tools/profile_forward.py 4 30
RDM_TC_NO_HALO=1 timeout 100 python tools/profile_forward.py 4 30
done
timeout 200 python tools/profile_forward_r.py 4 10
RDM_TC_NO_HALO=1 timeout 200 python tools/profile_forward_r.py 4 10
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_zx_benchmarked_config_gpu.py -m gpu -q -x -k "not knn" 2>&1 | tail -3
