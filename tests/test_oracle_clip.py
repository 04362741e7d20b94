"""CPU: the CLIP oracle restatement vs golden vectors produced by the REFERENCE's own vendored implementation."""
import os

import numpy as np
import torch

from conftest import ROOT
from oracle import clip as oclip


def _golden():
    d = np.load(os.path.join(ROOT, "tests", "golden", "clip_small.npz"))
    sd = {k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd:")}
    cfg = dict(zip([str(k) for k in d["cfg_keys"]], [int(v) for v in d["cfg_vals"]]))
    return d, sd, cfg


def test_image_tower_matches_reference_golden():
    d, sd, cfg = _golden()
    got = oclip.encode_image(sd, torch.from_numpy(d["image"]))
    assert np.abs(got.numpy() - d["image_features"]).max() < 2e-5 * np.abs(d["image_features"]).max()


def test_text_tower_matches_reference_golden():
    d, sd, cfg = _golden()
    got = oclip.encode_text(sd, torch.from_numpy(d["tokens"]), cfg["transformer_heads"])
    assert np.abs(got.numpy() - d["text_features"]).max() < 2e-5 * np.abs(d["text_features"]).max()


def test_random_state_dict_has_the_reference_vit_b32_inventory():
    sd = oclip.random_state_dict(vocab_size=1000)            # small vocab: the count below excludes the embedding table
    n = sum(v.numel() for k, v in sd.items() if k != "token_embedding.weight")
    # custom_clip ViT-B/32 = 151,277,313 parameters (151.28 M) of which 49408*512 are the token embedding
    assert n == 151_277_313 - 49408 * 512


def test_cfg_inference_from_state_dict_matches_build_model():
    from rdm_b200.clip import VIT_B32, cfg_from_state_dict
    d, sd, cfg = _golden()
    assert cfg_from_state_dict(sd) == cfg
    big = oclip.random_state_dict(**dict(VIT_B32, vocab_size=1000))
    assert cfg_from_state_dict(big) == dict(VIT_B32, vocab_size=1000)
