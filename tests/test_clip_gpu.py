"""GPU parity of the CLIP sink: librdm_b200 vs (a) golden vectors of the REFERENCE's own implementation, (b) the pinned oracle at
full ViT-B/32 size with random weights, (c) torch bicubic for the retriever preprocessing.  Tolerance 1e-3 relative;
observed ~1e-5 in the default bf16x3 mode."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import clip as oclip

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("mode", [0, 1])
def test_small_clip_matches_reference_golden(cuda, mode):
    from rdm_b200.clip import B200Clip, cfg_from_state_dict
    d = np.load(os.path.join(ROOT, "tests", "golden", "clip_small.npz"))
    sd = {k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd:")}
    cfg = dict(zip([str(k) for k in d["cfg_keys"]], [int(v) for v in d["cfg_vals"]]))
    assert cfg_from_state_dict(sd) == cfg                                   # build_model shape inference (model.py:363-391)
    m = B200Clip(cuda, **cfg)
    m.load_state_dict(sd)
    m.set_mode(mode)
    assert rel(m.encode_image(torch.from_numpy(d["image"])), d["image_features"]) < 1e-4
    assert rel(m.encode_text(torch.from_numpy(d["tokens"])), d["text_features"]) < 1e-4


def test_vit_b32_text_and_image_match_oracle(cuda):
    from rdm_b200.clip import VIT_B32, B200Clip
    cfg = dict(VIT_B32, vocab_size=2048)                                    # full towers; small vocab keeps the fixture light
    sd = oclip.random_state_dict(**cfg, seed=5)
    m = B200Clip(cuda, **cfg)
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(4)
    tok = torch.zeros(5, 77, dtype=torch.long)
    for b, L in enumerate((3, 12, 20, 77, 9)):                              # [SOT, ids, EOT, 0...] with EOT = largest id
        tok[b, 0] = 2046; tok[b, 1:L - 1] = torch.randint(1, 2000, (L - 2,), generator=g); tok[b, L - 1] = 2047
    img = torch.randn(3, 3, 224, 224, generator=g)
    assert rel(m.encode_text(tok), oclip.encode_text(sd, tok, 8)) < 1e-3
    assert rel(m.encode_image(img), oclip.encode_image(sd, img)) < 1e-3


def test_preprocess_matches_torch_bicubic(cuda):
    from rdm_b200.clip import VIT_B32, B200Clip
    m = B200Clip(cuda, **dict(VIT_B32, vocab_size=64, vision_layers=1, transformer_layers=1))
    g = torch.Generator().manual_seed(2)
    for H, W in ((256, 256), (64, 96), (300, 200)):
        x = torch.rand(2, 3, H, W, generator=g) * 2 - 1
        got = m.preprocess(x).cpu()
        want = oclip.preprocess(x)
        assert float((got - want).abs().max()) < 2e-5
