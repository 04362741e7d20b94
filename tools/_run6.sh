set -x
RDM_B200_LIB=retrieval-augmented-diffusion-models_b200/csrc/build/ab/timing.so timeout 300 python tools/profile_forward.py 4 2 > gpurun_out/tct_fp16_r2f.log 2>&1
wc -l gpurun_out/tct_fp16_r2f.log; tail -3 gpurun_out/tct_fp16_r2f.log
