set -x
# launches: skip the eager warm-up forward (197 GEMMs) + capture, profile a window of the graph-replayed forward
timeout 900 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel --launch-skip 203 --launch-count 12 -o gpurun_out/gemm_tc_r2g python tools/profile_forward.py 4 1 > gpurun_out/ncu_gemm_r2g.log 2>&1
tail -5 gpurun_out/ncu_gemm_r2g.log
ls -la gpurun_out/gemm_tc_r2g.ncu-rep
