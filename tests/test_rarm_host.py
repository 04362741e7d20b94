"""CPU: host side of the RARM mirror (rdm.modules.attention.RetrievalPatchTransformer, rdm.models.autoregression.transformer.
LatentImageRETRO): constructor keys of the shipped YAML, state-dict layout of the reference, sampling control flow over a stand-in
engine (the oracle), failure without a CUDA device."""
import ast
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import rarm as orarm

GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLD)
import ref_weights  # noqa: E402

import rdm  # noqa: F401,E402  (installs the shims)
from rdm.models.autoregression.transformer import LatentImageRETRO, SOSProvider  # noqa: E402
from rdm.modules.attention import RetrievalPatchTransformer  # noqa: E402

SMALL = ast.literal_eval(str(np.load(os.path.join(GOLD, "ref_rarm_small.npz"))["cfg_json"]))


def model_cfg(tcfg):
    """models/rarm/imagenet/dogs/config.yaml:2-27 with a small transformer and without retrieval / first stage."""
    return dict(mask_token=tcfg["in_channels"] - 2, sos_token=tcfg["in_channels"] - 1, p_mask_max=0.0, nn_key="nn_embeddings",
                nn_reshaper_cfg={"target": "rdm.modules.encoders.nn_encoders.CLIPEmbeddingReshaper"},
                nn_encoder_cfg={"target": "rdm.modules.encoders.nn_encoders.IdentityEncoder"},
                transformer_config={"target": "rdm.modules.attention.RetrievalPatchTransformer", "params": dict(tcfg)},
                first_stage_config=None, retrieval_cfg=None, cond_stage_config="__is_unconditional__")


class OracleEngine:
    """Stand-in for B200Rarm on a CUDA-less host: same calls, evaluated by the oracle."""

    def __init__(self, sd, heads):
        self.sd, self.heads = sd, heads

    def set_context(self, r):
        self.r = r

    def sample(self, prefix, steps, temperature=1.0, top_k=None, guidance_scale=1.0, uniforms=None):
        B = prefix.shape[0]
        toks, _ = orarm.sample(self.sd, self.heads, prefix[:, :1], prefix[:, 1:], self.r[:B], steps, temperature, top_k, guidance_scale, uniforms)
        return torch.cat([prefix[:, :1], toks], 1)


def test_transformer_state_dict_is_the_reference_layout():
    d = np.load(os.path.join(GOLD, "ref_rarm_small.npz"))
    m = RetrievalPatchTransformer(**SMALL)
    assert list(m.state_dict().keys()) == [str(k) for k in d["sd_keys"]]
    assert sum(p.numel() for p in m.parameters()) == int(d["n_params"])
    sd = ref_weights.state_dict_for(((k, v.shape) for k, v in m.state_dict().items()), 21)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 3, dtype=torch.long), context=torch.zeros(1, 4, SMALL["context_dim"]))
    with pytest.raises(NotImplementedError):
        RetrievalPatchTransformer(**dict(SMALL, continuous=True))


def test_sos_conditioning():
    _, _, (_, _, c) = SOSProvider(16385).encode(torch.zeros(3, 0))
    assert c.tolist() == [[16385]] * 3 and c.dtype == torch.int64


def test_model_builds_from_the_shipped_config_keys_and_loads_a_checkpoint_layout():
    m = LatentImageRETRO(**model_cfg(SMALL)).eval()
    keys = list(m.state_dict().keys())
    assert "sos_token" in keys and "mask_token" in keys and "transformer.positional_encoding" in keys and "transformer.proj_out.bias" in keys
    # a Lightning checkpoint also holds the first stage; without a first-stage container those keys are ignored, everything else is strict
    sd = dict(m.state_dict(), **{"first_stage_model.decoder.conv_in.weight": torch.zeros(1)})
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    with pytest.raises(NotImplementedError):
        m.training_step(None, 0)


def test_sample_control_flow_matches_the_oracle_loop():
    m = LatentImageRETRO(**model_cfg(SMALL)).eval()
    sd = ref_weights.state_dict_for(((k, v.shape) for k, v in m.transformer.state_dict().items()), 21)
    m.transformer.load_state_dict(sd)
    m.transformer.engine = lambda device: OracleEngine(sd, SMALL["n_heads"])
    g = torch.Generator().manual_seed(0)
    r = torch.randn(2, 4, SMALL["context_dim"], generator=g)
    _, c = m.encode_to_c(torch.zeros((2, 0)))
    z0 = torch.zeros((2, 0), dtype=torch.long)
    calls = []
    for scale in (1.0, 3.0):
        torch.manual_seed(7)
        got = m.sample(z0, r, c, steps=5, temperature=0.9, sample=True, top_k=8, guidance_scale=scale, callback=calls.append)
        torch.manual_seed(7)
        u = torch.rand((5, 2))
        want, _ = orarm.sample(sd, SMALL["n_heads"], c, z0, r, 5, temperature=0.9, top_k=8, guidance_scale=scale, uniforms=u)
        assert torch.equal(got, want) and got.shape == (2, 5)
    assert calls == list(range(5)) * 2
    greedy = m.sample(z0, r, c, steps=3, sample=False)
    assert torch.equal(greedy, orarm.sample(sd, SMALL["n_heads"], c, z0, r, 3)[0])
    # sample_from_rdata with given neighbour embeddings (scripts/rarm_sample.py --only_caption / --unconditional path)
    torch.manual_seed(3)
    out = m.sample_from_rdata(2, nn_embeddings=r, k_nn=4, top_k=8, code_side_len=2, z_dimensionality=8)
    assert out["samples_with_sampled_nns"].shape == (2, 4) and torch.equal(out["sampled_indices"], out["samples_with_sampled_nns"])
    with pytest.raises(NotImplementedError):
        m.decode_to_img(out["sampled_indices"], (2, 8, 2, 2))


def test_get_qids_follows_numpy_global_rng():
    m = LatentImageRETRO(**model_cfg(SMALL)).eval()
    m.use_memory, m.id_count = True, {i: i + 1 for i in range(100)}
    m.register_buffer("nn_memory", torch.arange(100, dtype=torch.int) * 3 % 100, persistent=False)
    np.random.seed(4)
    got = m.get_qids(0.5, 6)
    np.random.seed(4)
    assert np.array_equal(got, np.random.choice(m.nn_memory.numpy()[:50], size=6))
    np.random.seed(4)
    got = m.get_qids(20, 6, use_weights=True)
    mem = m.nn_memory.numpy()[:20]
    f = np.asarray([m.id_count[int(i)] for i in mem])
    np.random.seed(4)
    assert np.array_equal(got, np.random.choice(mem, size=6, p=f / f.sum(keepdims=True)))


def test_first_stage_in_the_taming_layout_decodes_the_sampled_ids(monkeypatch):
    """models/rarm/imagenet/*/config.yaml:28-51 at reduced widths: `first_stage_config.target: taming.models.vqgan.VQModel` resolves (to the
    stand-in when taming is not installed), loads a checkpoint in the taming key layout, and `decode_to_img` == the oracle's restatement of
    taming's `decode_to_img` (codebook entries -> post_quant_conv -> Decoder); the device decoder is replaced by the oracle's (no GPU here)."""
    from oracle import vqdecoder as ovq
    vq = ovq.TINY_VQ_WIDE
    dd = dict(vq["ddconfig"], resolution=8)
    cfg = model_cfg(dict(SMALL, sequence_length=16))
    cfg["first_stage_config"] = {"target": "taming.models.vqgan.VQModel",
                                 "params": dict(embed_dim=vq["embed_dim"], n_embed=vq["n_embed"], ddconfig=dd, lossconfig={"target": "torch.nn.Identity"})}
    m = LatentImageRETRO(**cfg).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):                    # the product's first stage only runs on the device
        m.decode_to_img(torch.zeros(1, 16, dtype=torch.long), (1, vq["embed_dim"], 4, 4))
    from ldm.models.autoencoder import VQModelInterface

    def first_stage_decode(self, h, force_not_quantize=False):                 # stand-in for the device decoder: the oracle's
        ref = ovq.VQModelInterface(self.embed_dim, self.quantize.embedding.num_embeddings, self._ddconfig).eval()
        ref.load_state_dict(self.state_dict())
        with torch.no_grad():
            return ref.decode(h, force_not_quantize)
    monkeypatch.setattr(VQModelInterface, "decode", first_stage_decode)
    fs = ovq.randomize_(ovq.VQModelInterface(embed_dim=vq["embed_dim"], n_embed=vq["n_embed"], ddconfig=dd), 34).eval()
    sd = {"first_stage_model." + k: v for k, v in fs.state_dict().items()}
    sd.update({k: v for k, v in m.state_dict().items() if not k.startswith("first_stage_model.")})
    missing, unexpected = m.load_state_dict(sd, strict=True)                  # the notebook loads strictly (demo_rarm.ipynb cell 5)
    assert not missing and not unexpected
    tsd = {k: v for k, v in m.transformer.state_dict().items()}
    m.transformer.engine = lambda device: OracleEngine(tsd, SMALL["n_heads"])
    torch.manual_seed(1)
    r = torch.randn(2, 4, SMALL["context_dim"])
    out = m.sample_from_rdata(2, nn_embeddings=r, k_nn=4, top_k=12, guidance_scale=2.0, code_side_len=4, z_dimensionality=vq["embed_dim"])
    ids, img = out["sampled_indices"], out["samples_with_sampled_nns"]
    assert ids.shape == (2, 16) and img.shape == (2, 3, 8, 8)
    with torch.no_grad():
        assert torch.equal(img, ovq.decode_indices(fs, ids, (2, vq["embed_dim"], 4, 4)))
