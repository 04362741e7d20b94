set -x
for i in 1 2; do
python tools/profile_forward.py 4 200
RDM_TC_2SM_WAVES=0 python tools/profile_forward.py 4 200
done
timeout 1400 python -m pytest tests/test_unet_gpu.py tests/test_variants_gpu.py tests/test_zx_benchmarked_config_gpu.py tests/test_vqdecoder_gpu.py tests/test_zy_ref_golden_gpu.py tests/test_script_flow_gpu.py tests/test_mirror_gpu.py -m gpu -q -x 2>&1 | tail -8
