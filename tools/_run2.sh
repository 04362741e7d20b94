set -x
python -m pytest tests -m gpu -q --durations=12 > gpurun_out/pytest_r2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2b.log
tail -40 gpurun_out/pytest_r2b.log
python tools/chains_sweep.py 3,4 1,2,3,4,8 > gpurun_out/chains_r2b.log 2>&1
python tools/chains_sweep.py 3,4 1,2,4 16,3,64 > gpurun_out/chains_r2b_rshape.log 2>&1
cat gpurun_out/chains_r2b.log gpurun_out/chains_r2b_rshape.log
