import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "retrieval-augmented-diffusion-models_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


# The CLIP merge table is a data file of the CLIP release that this repository does not ship; CPU tests in the build container may use the
# reference's copy (tests may read /root/reference; the product never does).
_BPE = "/root/reference/rdm/modules/custom_clip/bpe_simple_vocab_16e6.txt.gz"
if "CLIP_BPE_PATH" not in os.environ and os.path.isfile(_BPE):
    os.environ["CLIP_BPE_PATH"] = _BPE


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
