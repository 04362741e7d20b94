set -x
python -m pytest tests/test_knn_gpu.py tests/test_zy_ref_golden_gpu.py tests/test_zx_benchmarked_config_gpu.py -m gpu -q -x -k "knn or search or normalisation or shards or duplicates" > gpurun_out/pytest_r2d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2d.log
tail -15 gpurun_out/pytest_r2d.log
python tools/knn_sweep.py --n 1281167 --q 1 4 16 32 64 256 --dtypes float16 --out gpurun_out/knn_sweep_r2d.json > gpurun_out/knn_sweep_r2d.log 2>&1; cat gpurun_out/knn_sweep_r2d.log
RDM_KNN_NO_FUSED=1 python tools/knn_sweep.py --n 1281167 --q 16 64 --dtypes float16 > gpurun_out/knn_sweep_r2d_nofused.log 2>&1; cat gpurun_out/knn_sweep_r2d_nofused.log
python tools/knn_sweep.py --n 20000000 --q 16 64 --dtypes float16 > gpurun_out/knn_sweep_r2d_20m.log 2>&1; cat gpurun_out/knn_sweep_r2d_20m.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:knn_ -c 40 --csv --log-file gpurun_out/knn_launches_r2d.csv python tools/knn_sweep.py --n 1281167 --q 16 --dtypes float16 > /dev/null 2>&1
grep -v "^==" gpurun_out/knn_launches_r2d.csv | cut -d, -f5,14- | tail -14
