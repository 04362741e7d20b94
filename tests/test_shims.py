"""CPU: the import shims give the unchanged reference callers what they use."""
import os
import sys

from conftest import PKG


def _shims():
    from rdm_b200 import compat
    compat.install_shims()


def test_omegaconf_shim_is_mutable_attribute_mapping(tmp_path):
    _shims()
    from omegaconf import OmegaConf
    p = tmp_path / "c.yaml"
    p.write_text("model:\n  target: torch.nn.Identity\n  params:\n    retrieval_cfg:\n      params:\n        gpu: true\n        retriever_config:\n          params: {model: ViT-B/32}\n    lst: [8, 4, 2]\n")
    cfg = OmegaConf.load(p)
    cfg.model.params.retrieval_cfg.params.gpu = False                      # scripts/rdm_sample.py:157-160
    cfg.model.params.retrieval_cfg.params.retriever_config.params.device = "cpu"
    assert cfg.model.params.retrieval_cfg.params.gpu is False and cfg["model"]["params"]["lst"] == [8, 4, 2]
    assert cfg.model.params.retrieval_cfg.params.retriever_config.params.device == "cpu"


def test_instantiate_from_config_semantics():
    _shims()
    import torch
    from ldm.util import instantiate_from_config
    assert isinstance(instantiate_from_config({"target": "torch.nn.Identity"}), torch.nn.Identity)
    assert instantiate_from_config("__is_unconditional__") is None and instantiate_from_config("__is_first_stage__") is None
    m = instantiate_from_config({"target": "torch.nn.Linear", "params": {"in_features": 3, "out_features": 2}})
    assert m.weight.shape == (2, 3)


def test_seed_everything_seeds_numpy_and_torch():
    _shims()
    import numpy as np
    import torch
    from pytorch_lightning import seed_everything
    seed_everything(7); a, b = np.random.rand(), torch.rand(1)
    seed_everything(7); assert a == np.random.rand() and torch.equal(b, torch.rand(1))
