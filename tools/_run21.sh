set -x
for i in 1 2; do timeout 100 python tools/profile_forward.py 4 30; RDM_GN_CLUSTER=2 timeout 100 python tools/profile_forward.py 4 30; done
