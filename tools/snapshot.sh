#!/usr/bin/env bash
# Full measurement snapshot of one state of the tree on a 1-GPU box:  gpurun --timeout 3000 -- 'bash tools/snapshot.sh r2m'
set -x
tag=${1:-snap}
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log; tail -8 gpurun_out/pytest_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cat gpurun_out/bench_$tag.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference_arm.json 2> /dev/null; cat gpurun_out/bench_${tag}_reference_arm.json | cut -c1-200
timeout 600 python bench.py --mode fp16x2 --no-cpu-baseline --no-r-shape > gpurun_out/bench_${tag}_fp16x2.json 2> /dev/null; cut -c1-120 gpurun_out/bench_${tag}_fp16x2.json
timeout 600 python tools/cfg_bench.py --cfg 3 > gpurun_out/cfg3_n1_$tag.json 2> gpurun_out/cfg3_n1_$tag.err; cat gpurun_out/cfg3_n1_$tag.json
timeout 600 python tools/cfg_bench.py --cfg 4 > gpurun_out/cfg4_n1_$tag.json 2> gpurun_out/cfg4_n1_$tag.err; cat gpurun_out/cfg4_n1_$tag.json
timeout 300 python tools/rarm_bench.py > gpurun_out/rarm_bench_$tag.json 2> /dev/null; cut -c1-250 gpurun_out/rarm_bench_$tag.json
timeout 300 python tools/vqdec_bench.py > gpurun_out/vqdec_bench_$tag.json 2> /dev/null; cut -c1-300 gpurun_out/vqdec_bench_$tag.json
timeout 600 python tools/db_load_bench.py 1000000 4 > gpurun_out/db_load_$tag.json 2> gpurun_out/db_load_$tag.err; cat gpurun_out/db_load_$tag.json
# cfg5 grid: rows x queries x k x dtype, clocks recorded per line
python tools/knn_sweep.py --n 1281167 --q 1 4 16 32 64 128 256 1024 --k 4 --dtypes float16 float32 --out gpurun_out/knn_sweep_${tag}_1m_k4.json > gpurun_out/knn_sweep_$tag.log 2>&1
python tools/knn_sweep.py --n 1281167 --q 16 64 --k 8 --dtypes float16 --out gpurun_out/knn_sweep_${tag}_1m_k8.json >> gpurun_out/knn_sweep_$tag.log 2>&1
python tools/knn_sweep.py --n 1281167 --q 16 64 --k 20 --dtypes float16 --out gpurun_out/knn_sweep_${tag}_1m_k20.json >> gpurun_out/knn_sweep_$tag.log 2>&1
python tools/knn_sweep.py --n 20927907 --q 1 16 64 256 --k 4 --dtypes float16 --out gpurun_out/knn_sweep_${tag}_20m_k4.json >> gpurun_out/knn_sweep_$tag.log 2>&1
python tools/knn_sweep.py --n 50000000 --q 16 64 --k 4 --dtypes float16 --out gpurun_out/knn_sweep_${tag}_50m_k4.json >> gpurun_out/knn_sweep_$tag.log 2>&1
cat gpurun_out/knn_sweep_$tag.log
