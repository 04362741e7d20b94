set -x
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_r2e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2e.log
tail -15 gpurun_out/pytest_r2e.log
timeout 300 python tools/profile_forward.py 4 > gpurun_out/profile_fwd_fp16_r2e.log 2>&1; tail -60 gpurun_out/profile_fwd_fp16_r2e.log
timeout 300 python tools/ablate_forward.py 4 > gpurun_out/ablate_r2e.log 2>&1; cat gpurun_out/ablate_r2e.log
