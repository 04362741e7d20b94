set -x
RDM_GN_TRACE=1 timeout 100 python tools/profile_forward.py 4 1 2> gpurun_out/gn_trace.log; sort gpurun_out/gn_trace.log | uniq -c | sort -nr | head -60
timeout 600 python -m pytest tests/test_script_flow_gpu.py -m gpu -q -x 2>&1 | tail -5
for v in "RDM_TC_NOSPLIT=1" "RDM_TC_CLUSTER=0" "RDM_TC_CLUSTER=2" "RDM_PDL_GLUE=1"; do env $v timeout 100 python tools/profile_forward.py 4 30; done
