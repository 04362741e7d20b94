set -x
timeout 100 python tools/profile_forward.py 4 30
for v in "RDM_TC_FIX=3" "RDM_TC_FIX=10" "RDM_TC_RED=2" "RDM_TC_RED=10" "RDM_TC_MINKB=2" "RDM_TC_MINKB=8" "RDM_TC_FIX=10 RDM_TC_RED=10" "RDM_TC_FIX=3 RDM_TC_RED=2 RDM_TC_MINKB=2"; do echo "== $v"; env $v timeout 100 python tools/profile_forward.py 4 30; done
timeout 900 python -m pytest tests/test_variants_gpu.py tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -4
