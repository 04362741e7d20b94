"""CPU: the strict fp32 mode of the CLIP executor (csrc/clip.cu over kernels.cu / gemm_simt.cu, as written) under the host emulation of
CUDA (tests/emu/), through the C ABI and the product's Python wrapper, against tests/golden/clip_small.npz -- features computed by the
REFERENCE's own vendored CLIP (rdm/modules/custom_clip/model.py) -- and against torch's bicubic resize for the retriever preprocessing."""
import contextlib
import ctypes
import os
import shutil
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import clip as oclip

sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ to build the emulated library")


@pytest.fixture()
def emulated(monkeypatch):
    import build_emu
    from rdm_b200 import _lib
    L = _lib.bind(ctypes.CDLL(build_emu.build()), [n for n in _lib.SIGNATURES if n.startswith("rdm_clip_")] + ["rdm_last_error", "rdm_launch_count"])
    monkeypatch.setattr(_lib, "_lib", L)
    monkeypatch.setattr(_lib, "resolve_device", lambda d: torch.device("cpu"))
    monkeypatch.setattr(_lib, "device_ctx", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(_lib, "stream_ptr", lambda d=None: None)
    return L


def rel(a, b):
    a, b = a.double(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm())


def test_small_clip_matches_reference_golden(emulated):
    from rdm_b200.clip import B200Clip, cfg_from_state_dict
    d = np.load(os.path.join(ROOT, "tests", "golden", "clip_small.npz"))
    sd = {k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd:")}
    m = B200Clip("cpu", **cfg_from_state_dict(sd))
    m.load_state_dict(sd)
    m.set_mode(0)
    assert rel(m.encode_image(torch.from_numpy(d["image"])), d["image_features"]) < 1e-5
    assert rel(m.encode_text(torch.from_numpy(d["tokens"])), d["text_features"]) < 1e-5
    x = torch.rand(1, 3, 20, 28, generator=torch.Generator().manual_seed(2)) * 2 - 1
    assert float((m.preprocess(x, size=32) - oclip.preprocess(x, size=32)).abs().max()) < 2e-5
