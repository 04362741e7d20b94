"""CPU: the product's `rdm.models.diffusion.ddim.DDIMSampler` (generic per-step path, reachable with any eps-model) against
tests/golden/ref_sampler_options.npz -- the REFERENCE's own DDIMSampler.sample (ddim.py:59-268) run on the closed-form model of
tests/golden/retro_stub.py with the options outside the plain guided loop: inpainting mask, eta > 0 with temperature, noise dropout,
style / content conditioning switched by SNR, callbacks, x_T drawn by the sampler, intermediates every log_every_t.  Same seed of the
global torch generator -> same tensors, same conditioning seen by the model at every step, and the same generator state left behind."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLD)
import make_golden_ref as gen  # noqa: E402  (SAMPLER_CASES / run_sampler_case only; nothing reference-side is imported)
import retro_stub  # noqa: E402


@pytest.mark.parametrize("tag", list(gen.SAMPLER_CASES))
def test_option_paths_match_the_reference_sampler(tag):
    import rdm  # noqa: F401
    from rdm.models.diffusion.ddim import DDIMSampler
    g = np.load(os.path.join(GOLD, "ref_sampler_options.npz"))
    got = gen.run_sampler_case(DDIMSampler, retro_stub.RetroStub().setup(), gen.SAMPLER_CASES[tag], retro_stub)
    for k, v in got.items():
        want = g[f"{tag}:{k}"]
        assert v.shape == want.shape, (tag, k, v.shape, want.shape)
        assert np.allclose(v, want, rtol=1e-5, atol=1e-6), (tag, k)
