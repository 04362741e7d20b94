"""GPU parity of the RARM decoder (csrc/rarm.cu) through the C ABI: key/value-cached logits vs logits of the REFERENCE's own
RetrievalPatchTransformer (tests/golden/ref_rarm_small.npz) and vs the pinned oracle at the ImageNet model size; the fused
guidance / top-k / draw kernel vs the oracle's restatement of transformer.py:249-266; the graph-replayed sampling loop token by token.
Tolerances: 1e-4 rel-L2 on logits with fp32 weights, 5e-3 with fp16 weights; a drawn token must be the inverse-CDF pick of its
uniform up to 1e-4 of probability mass (index-exact away from CDF boundaries).  (File name: runs after the U-Net / kNN suites.)"""
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

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def rel(a, b):
    a, b = a.double().cpu(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm())


def small(cuda, mode):
    from rdm_b200.rarm import B200Rarm
    d = np.load(os.path.join(GOLD, "ref_rarm_small.npz"))
    cfg = ast.literal_eval(str(d["cfg_json"]))
    net = B200Rarm(cuda, **cfg)
    assert list(net.shapes) == [str(k) for k in d["sd_keys"]]
    sd = ref_weights.state_dict_for(net.shapes.items(), int(d["weight_seed"]))
    net.load_state_dict(sd)
    assert net.missing() == 0
    net.set_mode(mode)
    return d, cfg, sd, net


def consistent_with_uniform(probs, token, u, tol=1e-4):
    """token is the inverse-CDF pick for u under `probs` (float64 CDF), allowing `tol` of probability mass at the boundaries."""
    cdf = probs.double().cumsum(-1)
    cdf = cdf / cdf[-1]
    lo = float(cdf[token - 1]) if token > 0 else 0.0
    return float(probs[token]) > 0 and lo - tol <= u <= float(cdf[token]) + tol


@pytest.mark.parametrize("mode,tol", [(0, 1e-4), (4, 5e-3)])
def test_cached_logits_match_reference_code(cuda, mode, tol):
    d, cfg, sd, net = small(cuda, mode)
    tok, ctx = torch.from_numpy(d["tokens"]), torch.from_numpy(d["context"])
    got = net.forward(tok, ctx)
    assert rel(got, d["logits"]) < tol
    assert rel(got[:, :5], d["logits_prefix5"]) < tol                       # causality: the cache never leaks later positions
    if mode == 4:                                                           # tight: the oracle on the same fp16-rounded dense weights
        assert rel(got, orarm.forward(ref_weights.round_dense_weights_to_fp16(sd), tok, ctx, cfg["n_heads"])) < 1e-4
    # guidance doubling (transformer.py:233-248): B tokens against 2B context rows [r | zeros]
    net.set_context(torch.cat([ctx, torch.zeros_like(ctx)]))
    both = torch.stack([net.forward_token(tok[:, t], t) for t in range(tok.shape[1])], 1)
    assert rel(both[:3], d["logits"]) < tol and rel(both[3:], d["logits_uncond"]) < tol


@pytest.mark.parametrize("V,top_k,guided", [(16384, 256, True), (16384, 256, False), (48, 5, True), (1000, None, False), (16384, 1, True)])
def test_guided_topk_draw_kernel_matches_oracle(cuda, V, top_k, guided):
    _, _, _, net = small(cuda, 0)
    g = torch.Generator().manual_seed(V + (top_k or 0))
    B = 4
    lc, lu = torch.randn(B, V, generator=g) * 3, torch.randn(B, V, generator=g) * 3
    lc[0, 7] = lc[0, 9]                                                       # a tie inside the candidates
    scale, temp = (2.5, 0.8) if guided else (1.0, 1.3)
    u = torch.rand(B, generator=g)
    logits = torch.cat([lc, lu]) if guided else lc
    tok, probs = net.sample_step(logits, guidance_scale=scale, temperature=temp, top_k=top_k, uniforms=u, want_probs=True)
    want = orarm.step_probs(lc, lu if guided else None, scale, temp, top_k)
    assert torch.equal(probs.cpu() > 0, want > 0)                             # the same candidate set (ties at the k-th value kept)
    assert float((probs.cpu() - want).abs().max()) < 1e-6
    for b in range(B):
        assert consistent_with_uniform(want[b], int(tok[b]), float(u[b]), 1e-5)
    greedy, _ = net.sample_step(logits, guidance_scale=scale, temperature=temp, top_k=top_k, uniforms=None)
    assert torch.equal(greedy.cpu(), want.argmax(-1))


@pytest.mark.parametrize("mode", [0, 4])
@pytest.mark.parametrize("scale", [1.0, 2.0])
def test_sampling_loop_token_by_token(cuda, mode, scale):
    d, cfg, sd, net = small(cuda, mode)
    ctx = torch.from_numpy(d["context"])
    B, steps, top_k, temp = 3, 11, 6, 0.9
    c = torch.full((B, 1), cfg["in_channels"] - 1)
    g = torch.Generator().manual_seed(5)
    u = torch.rand(steps, B, generator=g)
    r = torch.cat([ctx, torch.zeros_like(ctx)]) if scale > 1.0 else ctx
    net.set_context(r)
    toks = net.sample(c, steps, temperature=temp, top_k=top_k, guidance_scale=scale, uniforms=u).cpu()
    assert toks.shape == (B, 1 + steps) and torch.equal(toks[:, :1], c) and int(toks[:, 1:].max()) < cfg["out_channels"]
    net.set_graph(False)                                                      # eager launches give the same tokens as the graph replay
    net.set_context(r)
    assert torch.equal(net.sample(c, steps, temperature=temp, top_k=top_k, guidance_scale=scale, uniforms=u).cpu(), toks)
    net.set_graph(True)
    # teacher-forced check against the oracle: every drawn token is the inverse-CDF pick of its uniform under the ORACLE's probabilities
    # (fp16 mode: the oracle on the same fp16-rounded dense weights, so candidate sets and CDF boundaries agree to fp32 rounding)
    osd = sd if mode == 0 else ref_weights.round_dense_weights_to_fp16(sd)
    lc = orarm.forward(osd, toks[:, :-1], ctx, cfg["n_heads"])
    lu = orarm.forward(osd, toks[:, :-1], torch.zeros_like(ctx), cfg["n_heads"]) if scale > 1.0 else None
    tol = 1e-4
    for t in range(steps):
        p = orarm.step_probs(lc[:, t], None if lu is None else lu[:, t], scale, temp, top_k)
        for b in range(B):
            assert consistent_with_uniform(p[b], int(toks[b, t + 1]), float(u[t, b]), tol), (t, b)
    # a start prefix (half-sampling, transformer.py:452-457): given tokens are kept, the rest continues from them
    net.set_context(r)
    cont = net.sample(toks[:, :5], steps - 4, temperature=temp, top_k=top_k, guidance_scale=scale, uniforms=u[4:]).cpu()
    assert torch.equal(cont, toks)
    # greedy decoding == argmax of the oracle's probabilities at every position (fp32 weights)
    if mode == 0:
        net.set_context(r)
        gr = net.sample(c, 6, temperature=1.0, top_k=None, guidance_scale=scale, uniforms=None).cpu()
        want, _ = orarm.sample(sd, cfg["n_heads"], c, torch.zeros((B, 0), dtype=torch.long), ctx, 6, guidance_scale=scale)
        assert torch.equal(gr[:, 1:], want)


def test_imagenet_size_decoder_against_oracle(cuda):
    """models/rarm/imagenet/*/config.yaml: 18 x (12 x 64), vocabulary 16386 -> 16384, k = 4 CLIP neighbours; 230.9 M parameters."""
    from rdm_b200.rarm import RARM_IMAGENET, B200Rarm
    net = B200Rarm(cuda, **RARM_IMAGENET)
    sd = ref_weights.state_dict_for(net.shapes.items(), 31)
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(6)
    tok = torch.randint(0, 16384, (2, 4), generator=g)
    tok[:, 0] = 16385
    ctx = torch.randn(2, 4, 512, generator=g)
    want = orarm.forward(sd, tok, ctx, 12)
    want16 = orarm.forward(ref_weights.round_dense_weights_to_fp16(sd), tok, ctx, 12)
    net.set_mode(0)
    assert rel(net.forward(tok, ctx), want) < 1e-4
    net.set_mode(4)                                                            # fp16 dense weights: tight against the oracle on the same rounded weights
    got16 = net.forward(tok, ctx)
    assert rel(got16, want16) < 1e-4 and rel(got16, want) < 3e-2
    # the full 256-step loop of scripts/rarm_sample.py (guided, batch 2 -> 4 rows): runs, stays in range, is reproducible
    u = torch.rand(256, 2, generator=g)
    r = torch.cat([ctx, torch.zeros_like(ctx)])
    net.set_context(r)
    a = net.sample(tok[:, :1], 256, temperature=1.0, top_k=256, guidance_scale=2.0, uniforms=u)
    net.set_context(r)
    b = net.sample(tok[:, :1], 256, temperature=1.0, top_k=256, guidance_scale=2.0, uniforms=u)
    assert a.shape == (2, 257) and torch.equal(a, b) and 0 <= int(a[:, 1:].min()) and int(a[:, 1:].max()) < 16384


# ---- first stage of the RARM models: taming VQGAN-f16 decode of the sampled ids (wide-latent path of the device decoder) -------------
@pytest.mark.parametrize("mode,tol", [(3, 5e-3), (4, 1e-2)])
def test_wide_latent_decoder_small(cuda, mode, tol):
    """embed_dim = z_channels = 64, AttnBlocks inside the lowest up level (attn_resolutions = [16]) as in the taming VQGAN-f16."""
    from oracle import vqdecoder as ovq
    from rdm_b200.vqdecoder import B200VQDecoder
    cfg = ovq.TINY_VQ_WIDE
    ref = ovq.randomize_(ovq.VQModelInterface(**cfg), 7).eval()
    dec = B200VQDecoder(cuda, cfg["embed_dim"], cfg["n_embed"], cfg["ddconfig"])
    assert sorted(dec.names) == sorted(ref.state_dict().keys())
    dec.load_state_dict(ref.state_dict())
    dec.set_mode(mode)
    idx = torch.randint(0, cfg["n_embed"], (2, 16 * 16), generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        want = ovq.decode_indices(ref, idx, (2, 64, 16, 16))
    quant = ref.quantize.embedding.weight.detach()[idx].view(2, 16, 16, 64).permute(0, 3, 1, 2).contiguous()
    got = dec.decode(quant.to(cuda), force_not_quantize=True)
    again = dec.decode(quant.to(cuda), force_not_quantize=True)               # CUDA-graph replay
    assert got.shape == want.shape == (2, 3, 32, 32)
    assert rel(got, want) < tol and rel(again, want) < tol
    with pytest.raises(RuntimeError, match="quantize"):
        dec.decode(quant.to(cuda))                                             # nearest-codebook search is not part of the sampling path


def test_latent_image_retro_samples_and_decodes_on_the_device(cuda):
    """scripts/rarm_sample.py end to end at reduced widths: YAML-style config -> LatentImageRETRO -> sample_from_rdata (given neighbour
    embeddings) -> 64 sampled ids per image (8 x 8 codes) -> taming-layout first stage -> images; the ids are checked token by token against the oracle."""
    import rdm  # noqa: F401
    from ldm.util import instantiate_from_config
    from omegaconf import OmegaConf
    from oracle import vqdecoder as ovq
    tcfg = dict(in_channels=98, n_heads=2, d_head=64, depth=2, context_dim=128, positional_encodings=True, sequence_length=64, out_channels=96,
                cross_attend=True, causal=True, continuous=False)
    vq = ovq.TINY_VQ_WIDE
    cfg = {"target": "rdm.models.autoregression.transformer.LatentImageRETRO",
           "params": dict(mask_token=96, sos_token=97, p_mask_max=0.0, nn_key="nn_embeddings",
                          nn_reshaper_cfg={"target": "rdm.modules.encoders.nn_encoders.CLIPEmbeddingReshaper"},
                          nn_encoder_cfg={"target": "rdm.modules.encoders.nn_encoders.IdentityEncoder"},
                          transformer_config={"target": "rdm.modules.attention.RetrievalPatchTransformer", "params": tcfg},
                          first_stage_config={"target": "taming.models.vqgan.VQModel",
                                              "params": dict(embed_dim=vq["embed_dim"], n_embed=vq["n_embed"], ddconfig=dict(vq["ddconfig"], resolution=16),
                                                             lossconfig={"target": "torch.nn.Identity"})},
                          retrieval_cfg=None, cond_stage_config="__is_unconditional__")}
    model = instantiate_from_config(OmegaConf.create(cfg)).eval()
    sd = ref_weights.state_dict_for(((k, v.shape) for k, v in model.transformer.state_dict().items()), 33)
    model.transformer.load_state_dict(sd)
    model.transformer.engine_mode = "fp32"
    fs = ovq.randomize_(ovq.VQModelInterface(embed_dim=vq["embed_dim"], n_embed=vq["n_embed"], ddconfig=dict(vq["ddconfig"], resolution=16)), 34).eval()
    model.first_stage_model.load_state_dict(fs.state_dict())
    model = model.to(cuda)
    g = torch.Generator().manual_seed(35)
    r = torch.randn(2, 4, 128, generator=g)
    torch.manual_seed(36)
    out = model.sample_from_rdata(2, nn_embeddings=r.to(cuda), k_nn=4, top_k=12, temperature=1.0, guidance_scale=2.0, code_side_len=8, z_dimensionality=64)
    ids, img = out["sampled_indices"].cpu(), out["samples_with_sampled_nns"]
    assert ids.shape == (2, 64) and img.shape == (2, 3, 16, 16) and int(ids.max()) < 96
    torch.manual_seed(36)
    u = torch.rand((64, 2), device=cuda).cpu()                                # the uniforms LatentImageRETRO.sample drew
    c = torch.full((2, 1), 97)
    full = torch.cat([c, ids], 1)
    lc = orarm.forward(sd, full[:, :-1], r, 2)
    lu = orarm.forward(sd, full[:, :-1], torch.zeros_like(r), 2)
    for t in range(64):
        p = orarm.step_probs(lc[:, t], lu[:, t], 2.0, 1.0, 12)
        for b in range(2):
            assert consistent_with_uniform(p[b], int(ids[b, t]), float(u[t, b]), 1e-4), (t, b)
    with torch.no_grad():
        want = ovq.decode_indices(fs, ids, (2, 64, 8, 8))
    assert rel(img, want) < 1e-2                                              # fp16x2 decoder (default mode) vs the fp32 oracle


def test_product_sampling_flow_matches_reference_code(cuda):
    """tests/golden/ref_rarm_sampling.npz -- the REFERENCE's own LatentImageRETRO.sample_from_rdata / sample over its own
    RetrievalPatchTransformer -- against this repository's LatentImageRETRO on the device decoder (fp32 weights): same seeds, database and
    weights -> same query ids, token ids and decoded images.  (The 128-wide toy database is searched by the oracle: the device searcher
    takes d in {256, 512, 768, 1024}.)"""
    import types
    import make_golden_ref as gen
    import rdm  # noqa: F401
    from oracle import knn as oknn
    from rdm.models.autoregression.transformer import LatentImageRETRO
    g = np.load(os.path.join(GOLD, "ref_rarm_sampling.npz"))
    model = LatentImageRETRO(**gen.rarm_model_cfg()).eval()
    tsd = ref_weights.state_dict_for(((k, v.shape) for k, v in model.transformer.state_dict().items()), 21)
    model.transformer.load_state_dict(tsd)
    model.transformer.engine_mode = "fp32"
    model = model.to(cuda)
    db = ref_weights.make_db(int(g["n_db"]))[0][:, :128].copy()

    class Searcher:
        def search_device(self, q_hat, k):
            i, d = oknn.search(db, q_hat.cpu().numpy(), k)
            return torch.from_numpy(i).to(cuda), torch.from_numpy(d).to(cuda)

        def gather_device(self, idx):
            return torch.from_numpy(db[idx.cpu().numpy()].astype(np.float32)).to(cuda)
    model.retriever = types.SimpleNamespace(searcher=Searcher(), data_pool={"embedding": db})
    real_rand = torch.rand
    for tag, kw in gen.RARM_SAMPLING_CASES.items():
        u = torch.from_numpy(g[f"{tag}:uniforms"])
        torch.rand = lambda *a, **k: u.to(k.get("device", "cpu"))              # the CPU generator's uniforms of the fixture (CUDA draws differ)
        try:
            np.random.seed(kw["seed"])
            logs = model.sample_from_rdata(2, qids=None, k_nn=4, memsize=100, top_k=kw["top_k"], temperature=kw["temperature"], code_side_len=3,
                                           z_dimensionality=8, guidance_scale=kw["guidance_scale"])
        finally:
            torch.rand = real_rand
        assert np.array_equal(np.asarray(logs["qids"]), g[f"{tag}:qids"])
        # token ids are compared EXACTLY; they are recovered from the fixture's images through the stub first stage (decode = tanh of the
        # first three code channels, one pixel per token).  The images themselves are compared to 1 ulp-scale tolerance only: the stub's
        # torch.tanh runs on the device here and on the CPU in the fixture, and the two math libraries round differently.
        cb = torch.tanh(model.first_stage_model.quantize.embedding.weight.detach().cpu()[:, :3])            # [n_embed, 3]
        want_img = torch.from_numpy(g[f"{tag}:images"])
        want_ids = (want_img.permute(0, 2, 3, 1).reshape(2, -1, 1, 3) - cb[None, None]).abs().sum(-1).argmin(-1)
        got_ids = logs["sampled_indices"].cpu().reshape(2, -1)
        assert torch.equal(got_ids, want_ids), (tag, got_ids.tolist(), want_ids.tolist())
        assert torch.allclose(logs["samples_with_sampled_nns"].cpu(), want_img, rtol=0, atol=2e-6), tag
    _, c = model.encode_to_c(torch.zeros((2, 0)))
    got = model.sample(torch.from_numpy(g["greedy:start"]).to(cuda), torch.from_numpy(g["greedy:r"]).to(cuda), c.to(cuda), steps=6, sample=False, top_k=None,
                       guidance_scale=3.0)
    assert torch.equal(got.cpu(), torch.from_numpy(g["greedy:tokens"]))
