set -x
timeout 900 python -m pytest tests/test_knn_gpu.py tests/test_zy_ref_golden_gpu.py tests/test_zx_benchmarked_config_gpu.py tests/test_db_loader_gpu.py -m gpu -q -x -k "knn or search or normalisation or shards or duplicates or load or hi_only" > gpurun_out/pytest_r2n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2n.log
tail -8 gpurun_out/pytest_r2n.log
python tools/knn_sweep.py --n 1281167 --q 16 32 64 128 256 1024 --k 4 --dtypes float16 > gpurun_out/knn_sweep_r2n.log 2>&1; cat gpurun_out/knn_sweep_r2n.log
RDM_KNN_HILO=1 python tools/knn_sweep.py --n 1281167 --q 32 64 --k 4 --dtypes float16
python tools/knn_sweep.py --n 20927907 --q 64 128 256 --k 4 --dtypes float16
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 100 python tools/profile_forward.py 4 30
