set -x
timeout 900 python -m pytest tests/test_knn_gpu.py tests/test_zy_ref_golden_gpu.py tests/test_zx_benchmarked_config_gpu.py tests/test_db_loader_gpu.py -m gpu -q -x -k "knn or search or normalisation or shards or duplicates or load" > gpurun_out/pytest_r2i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2i.log
tail -5 gpurun_out/pytest_r2i.log
python tools/knn_sweep.py --n 1281167 --q 1 4 16 32 64 --dtypes float16 --out gpurun_out/knn_sweep_r2i.json > gpurun_out/knn_sweep_r2i.log 2>&1; cat gpurun_out/knn_sweep_r2i.log
python tools/knn_sweep.py --n 20000000 --q 16 64 --dtypes float16 > gpurun_out/knn_sweep_r2i_20m.log 2>&1; cat gpurun_out/knn_sweep_r2i_20m.log
timeout 900 python bench.py > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; tail -3 gpurun_out/bench_r2i.err; cat gpurun_out/bench_r2i.json
timeout 600 python -m pytest tests/test_zz_rarm_gpu.py -m gpu -q -x > gpurun_out/pytest_r2i_rarm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2i_rarm.log; tail -4 gpurun_out/pytest_r2i_rarm.log
timeout 300 python tools/rarm_bench.py > gpurun_out/rarm_bench_r2i.json 2> gpurun_out/rarm_bench_r2i.err; cat gpurun_out/rarm_bench_r2i.json | cut -c1-400
