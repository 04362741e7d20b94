set -x
timeout 300 python -m pytest tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -12
timeout 100 python tools/profile_forward.py 4 30
RDM_TC_2SM=0 timeout 100 python tools/profile_forward.py 4 30
timeout 200 python tools/profile_forward.py 4 2>&1 | grep "M=32768\|M=8192" | head -24
