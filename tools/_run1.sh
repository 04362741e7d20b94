set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/pytest_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2a.log
tail -30 gpurun_out/pytest_r2a.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --mode fp16x2 > gpurun_out/bench_r2a_fp16x2.json 2> gpurun_out/bench_r2a.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --mode fp16 > gpurun_out/bench_r2a_fp16.json 2>> gpurun_out/bench_r2a.err
python tools/rarm_bench.py > gpurun_out/rarm_bench_r2a.json 2> gpurun_out/rarm_bench_r2a.err
python tools/rarm_bench.py --guidance 2.0 >> gpurun_out/rarm_bench_r2a.json 2>> gpurun_out/rarm_bench_r2a.err
python tools/ablate_forward.py 3 > gpurun_out/ablate_r2a.log 2>&1
tail -3 gpurun_out/bench_r2a_fp16x2.json gpurun_out/bench_r2a_fp16.json gpurun_out/rarm_bench_r2a.json gpurun_out/ablate_r2a.log
