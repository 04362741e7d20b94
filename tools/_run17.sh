set -x
timeout 100 python tools/profile_forward.py 4 30
RDM_TC_NO_HALO=1 timeout 100 python tools/profile_forward.py 4 30
RDM_TC_NO_L2_AHEAD=1 timeout 100 python tools/profile_forward.py 4 30
RDM_TC_NO_HALO=1 RDM_TC_NO_L2_AHEAD=1 timeout 100 python tools/profile_forward.py 4 30
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_clip_gpu.py tests/test_vqdecoder_gpu.py -m gpu -q -x 2>&1 | tail -3
