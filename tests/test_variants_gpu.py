"""Every scheduling variant of the tcgen05 engine must give the same answer: split-K through the reduce kernel vs through a thread-block
cluster (DSMEM reduction), fused vs un-fused cross-attention, with / without programmatic dependent launch.  The variants are chosen by
environment variables that librdm_b200 reads once at load time, hence one sub-process each (tools/variant_check.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = {"0": 1e-4, "1": 1e-4, "2": 3e-2, "3": 3e-3, "4": 4e-3}          # ("2" = plain bf16: 8-bit significands)          # same per-forward tolerances as tests/test_unet_gpu.py


@pytest.mark.parametrize("env", [{"RDM_TC_CLUSTER": "1"}, {"RDM_TC_CLUSTER": "2"}, {"RDM_SKIP": "64"}, {"RDM_PDL": "0"}, {"RDM_TC_NOSPLIT": "1"}, {"RDM_GN_EPI_STATS": "1"}, {"RDM_GN_FUSED_MAX_HW": "4096"}, {"RDM_TC_2SM": "2", "RDM_TC_2SM_MIN_M": "256", "MODES": "4,2"}, {"RDM_TC_2SM": "0", "MODES": "4"}, {"RDM_RES_SKIP_FUSED": "0", "MODES": "4,2"}, {"RDM_UP_FOLD": "0", "MODES": "4,3"}, {"RDM_TC_2SM_WAVES": "3", "MODES": "4"}],
                         ids=["splitk-cost-model-with-clusters", "splitk-cluster-dsmem", "unfused-cross-attention", "no-pdl", "no-splitk", "gn-epilogue-statistics", "gn-one-launch-everywhere", "cta-pairs-everywhere", "no-cta-pairs", "skip-connection-as-its-own-gemm", "upsample-materialised", "cta-pairs-for-geglu"])
def test_engine_variant_matches_oracle(cuda, env):
    env = dict(env)
    modes = env.pop("MODES", "1,3")                                # engine modes to check (the CTA-pair kernel exists for the one-MMA modes only)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "variant_check.py"), modes], env=dict(os.environ, **env),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    errs = json.loads(r.stdout.strip().splitlines()[-1])
    for mode, err in errs.items():
        assert err < TOL[mode], f"{env} mode {mode}: rel-L2 {err:.2e} (tolerance {TOL[mode]})"
