set -x
python tools/chains_debug.py > gpurun_out/chains_debug_r2c.log 2>&1
python -m pytest tests/test_unet_gpu.py tests/test_zx_benchmarked_config_gpu.py -m gpu -q -rP -k "chains or cfg2_ddim100 or reference_shape" > gpurun_out/pytest_r2c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2c.log
grep -n "rel-L2\|passed\|failed\|rc=" gpurun_out/pytest_r2c.log | tail -30
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -c 3000 gpurun_out/bench_r2c.json; tail -5 gpurun_out/bench_r2c.err
python tools/cfg_bench.py --cfg 3 > gpurun_out/cfg3_n1_r2c.json 2> gpurun_out/cfg3_n1_r2c.err; cat gpurun_out/cfg3_n1_r2c.json; tail -3 gpurun_out/cfg3_n1_r2c.err
python tools/cfg_bench.py --cfg 4 > gpurun_out/cfg4_n1_r2c.json 2> gpurun_out/cfg4_n1_r2c.err; cat gpurun_out/cfg4_n1_r2c.json; tail -3 gpurun_out/cfg4_n1_r2c.err
timeout 600 bash tools/sanitize.sh > gpurun_out/sanitize_r2c.log 2>&1; tail -15 gpurun_out/sanitize_r2c.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:knn_scan_tc -c 4 -o gpurun_out/knn_q64_r2c -f python tools/knn_sweep.py --n 1281167 --q 64 --dtypes float16 > gpurun_out/ncu_knn_q64_r2c.log 2>&1; tail -3 gpurun_out/ncu_knn_q64_r2c.log
cat gpurun_out/chains_debug_r2c.log
