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


def test_tokenizer_matches_reference_golden_ids():
    """Golden ids come from the reference's own tokenizer (tests/golden/make_golden_tokens.py).  Needs the CLIP merge table, a data
    file that is not shipped in this repo: skipped where it is absent (e.g. on the GPU box)."""
    import json
    import pytest
    _shims()
    import clip
    try:
        clip._find_vocab()
    except FileNotFoundError:
        pytest.skip("bpe_simple_vocab_16e6.txt.gz not available")
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "clip_tokens.json")))
    toks = clip.tokenize(list(g["prompts"].keys()))
    assert toks.shape == (len(g["prompts"]), 77) and toks.dtype.is_floating_point is False
    for row, (prompt, ids) in zip(toks, g["prompts"].items()):
        assert row[:len(ids)].tolist() == ids and int(row[len(ids):].abs().sum()) == 0, prompt
        assert int(row.argmax()) == len(ids) - 1                      # EOT is the largest id: encode_text gathers it (model.py:318)


def test_vendored_tokenizer_cuts_long_captions_like_the_reference():
    """rdm/modules/custom_clip/clip.py:127-143: over-long captions are cut to 77 ids with a warning and EOT is NOT re-inserted
    (OpenAI's clip.tokenize raises instead); CLIPTextEmbedder.preprocess goes through it (retrievers.py:9,111)."""
    import pytest
    _shims()
    import clip
    try:
        clip._find_vocab()
    except FileNotFoundError:
        pytest.skip("bpe_simple_vocab_16e6.txt.gz not available")
    import rdm  # noqa: F401
    from rdm.modules.custom_clip.clip import tokenize
    long_caption = " ".join(["a photo of a very fluffy corgi wearing a tiny hat"] * 12)
    with pytest.raises(RuntimeError):
        clip.tokenize(long_caption)
    toks = tokenize([long_caption, "a corgi"])
    assert toks.shape == (2, 77) and int(toks[0, 0]) == 49406 and int((toks[0] == 49407).sum()) == 0 and int(toks[0, -1]) != 0
    assert toks[1].tolist()[:4] == clip.tokenize("a corgi")[0].tolist()[:4] and int(toks[1].argmax()) == 3
    ref_path = "/root/reference/rdm/modules/custom_clip"
    if os.path.isdir(ref_path):                                            # the reference's own vendored tokenizer on the same caption
        import importlib.util
        import sys
        import types
        sys.modules.setdefault("ftfy", types.SimpleNamespace(fix_text=lambda x: x))          # absent here; an identity on ASCII captions
        spec = importlib.util.spec_from_file_location("ref_simple_tokenizer", os.path.join(ref_path, "simple_tokenizer.py"))
        mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
        t = mod.SimpleTokenizer()
        want = ([t.encoder["<|startoftext|>"]] + t.encode(long_caption) + [t.encoder["<|endoftext|>"]])[:77]
        assert toks[0].tolist() == want
