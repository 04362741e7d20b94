#!/usr/bin/env bash
# ncu / sanitizer evidence of one round, run on a GPU box:  gpurun --timeout 2400 -- 'bash tools/evidence.sh r2'
# Outputs go to gpurun_out/ (scratch); tools/evidence_summarise.py turns them into the tracked files under profiles/.
set -x
tag=${1:-r2}
# (1) every launch of the first DDIM steps of the benchmark with its device time (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_$tag.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-r-shape > /dev/null 2>&1
# (2) DRAM traffic of every tcgen05 GEMM launch of one graph-replayed forward (182 launches since the skip_connection fusion; the eager pass + capture come first)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tc_kernel --launch-skip 188 --launch-count 182 \
    --csv --log-file gpurun_out/gemm_tc_dram_$tag.csv python tools/profile_forward.py 4 1 > /dev/null 2>&1
# (3) --set full of the glue kernel classes inside the replayed forward
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'gn_stats|gn_apply|layernorm_kernel|attention_mma' --launch-skip 150 --launch-count 8 \
    -o gpurun_out/glue_$tag python tools/profile_forward.py 4 1 > gpurun_out/ncu_glue_$tag.log 2>&1
# (4) RARM decode step: launch list + --set full of the weight-streaming GEMV and the cached attention
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rarm_ --launch-skip 600 -c 300 --csv --log-file gpurun_out/launches_rarm_$tag.csv \
    python tools/rarm_bench.py > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'rarm_gemv|rarm_attn' --launch-skip 600 --launch-count 9 \
    -o gpurun_out/rarm_$tag python tools/rarm_bench.py > gpurun_out/ncu_rarm_$tag.log 2>&1
# (5) one kNN search at the headline size
timeout 600 ncu --set full --import-source on --clock-control none -k regex:knn_ --launch-skip 12 --launch-count 6 \
    -o gpurun_out/knn_$tag python tools/knn_sweep.py --n 1281167 --q 16 --dtypes float16 > gpurun_out/ncu_knn_$tag.log 2>&1
# (6) compute-sanitizer over the SIMT kernels
timeout 1500 bash tools/sanitize.sh > gpurun_out/sanitize_$tag.log 2>&1; echo "sanitize rc=$?" >> gpurun_out/sanitize_$tag.log
ls -la gpurun_out | head -30
