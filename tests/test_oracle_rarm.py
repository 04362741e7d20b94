"""oracle/rarm.py against logits of the REFERENCE's own RetrievalPatchTransformer (tests/golden/ref_rarm_small.npz), the KV-cache
formulation against the full-prefix one, and the sampling arithmetic of LatentImageRETRO.sample (transformer.py:224-270)."""
import ast
import os
import sys

import numpy as np
import torch

from conftest import ROOT
from oracle import rarm as orarm

GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLD)
import ref_weights  # noqa: E402


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm())


def golden():
    d = np.load(os.path.join(GOLD, "ref_rarm_small.npz"))
    cfg = ast.literal_eval(str(d["cfg_json"]))
    shapes = orarm.param_shapes(**cfg)
    assert list(shapes) == [str(k) for k in d["sd_keys"]]                       # the reference's keys in its registration order
    assert sum(int(np.prod(s)) for s in shapes.values()) == int(d["n_params"])
    return d, cfg, ref_weights.state_dict_for(shapes.items(), int(d["weight_seed"]))


def test_forward_matches_reference_code():
    d, cfg, sd = golden()
    tok, ctx = torch.from_numpy(d["tokens"]), torch.from_numpy(d["context"])
    assert rel(orarm.forward(sd, tok, ctx, cfg["n_heads"]), d["logits"]) < 2e-6
    assert rel(orarm.forward(sd, tok[:, :5], ctx, cfg["n_heads"]), d["logits_prefix5"]) < 2e-6
    assert rel(orarm.forward(sd, tok, torch.zeros_like(ctx), cfg["n_heads"]), d["logits_uncond"]) < 2e-6


def test_kv_cache_formulation_equals_full_prefix_recompute():
    d, cfg, sd = golden()
    tok, ctx = torch.from_numpy(d["tokens"]), torch.from_numpy(d["context"])
    inc = orarm.forward_incremental(sd, tok, ctx, cfg["n_heads"])
    assert rel(inc, d["logits"]) < 2e-6
    # causality: logits of a prefix are the prefix of the logits (what makes the cache valid)
    assert rel(inc[:, :5], d["logits_prefix5"]) < 2e-6


def test_imagenet_config_parameter_count():
    """models/rarm/imagenet/*/config.yaml:14-27: 18 layers x (12 x 64), vocabulary 16386 -> 16384."""
    shapes = orarm.param_shapes(**orarm.RARM_IMAGENET)
    n = sum(int(np.prod(s)) for s in shapes.values())
    C = 768
    per_layer = 3 * C * C + C * C + C + (C * C + 2 * C * 512 + C * C + C) + (8 * C * C + 8 * C + 4 * C * C + C) + 6 * C
    assert n == C * 256 + 16386 * C + 18 * per_layer + 16384 * C + 16384 == 230_874_112


def test_top_k_filter_and_guidance():
    g = torch.Generator().manual_seed(0)
    lc, lu = torch.randn(3, 40, generator=g), torch.randn(3, 40, generator=g)
    p = orarm.step_probs(lc, lu, 3.0, 0.7, 5)
    guided = (lu + 3.0 * (lc - lu)) / 0.7
    assert ((p > 0).sum(-1) == 5).all() and torch.allclose(p.sum(-1), torch.ones(3))
    kept = guided.topk(5).indices
    for b in range(3):
        assert set(torch.nonzero(p[b]).flatten().tolist()) == set(kept[b].tolist())
        assert torch.allclose(p[b, kept[b]], torch.softmax(guided[b, kept[b]], -1))
    # ties at the threshold are all kept (`out < v_k`, not a fixed count)
    t = torch.tensor([[1.0, 3.0, 2.0, 2.0, 0.0]])
    assert (orarm.top_k_logits(t, 2) == torch.tensor([[-float("inf"), 3.0, 2.0, 2.0, -float("inf")]])).all()


def test_draw_is_the_inverse_cdf():
    p = torch.tensor([[0.0, 0.25, 0.0, 0.5, 0.25], [1.0, 0.0, 0.0, 0.0, 0.0]])
    for u, want in ((0.0, [1, 0]), (0.2499, [1, 0]), (0.25, [3, 0]), (0.7499, [3, 0]), (0.75, [4, 0]), (0.9999, [4, 0])):
        assert orarm.draw(p, torch.tensor([u, u])).tolist() == want
    # empirical distribution
    g = torch.Generator().manual_seed(1)
    u = torch.rand(20000, generator=g)
    idx = orarm.draw(p[:1].expand(20000, 5), u)
    freq = torch.bincount(idx, minlength=5).float() / 20000
    assert torch.allclose(freq, p[0], atol=0.01)


def test_sample_loop_shapes_and_prefix_consistency():
    d, cfg, sd = golden()
    ctx = torch.from_numpy(d["context"])
    c = torch.full((3, 1), 49)
    x0 = torch.zeros((3, 0), dtype=torch.long)
    g = torch.Generator().manual_seed(3)
    u = torch.rand(6, 3, generator=g)
    toks, probs = orarm.sample(sd, cfg["n_heads"], c, x0, ctx, 6, temperature=1.0, top_k=8, guidance_scale=2.0, uniforms=u)
    assert toks.shape == (3, 6) and probs.shape == (6, 3, cfg["out_channels"]) and int(toks.max()) < cfg["out_channels"]
    # a continuation from the first 3 sampled tokens with the remaining uniforms reproduces the rest (the loop has no hidden state)
    toks2, _ = orarm.sample(sd, cfg["n_heads"], c, toks[:, :3], ctx, 3, temperature=1.0, top_k=8, guidance_scale=2.0, uniforms=u[3:])
    assert torch.equal(toks2, toks)
    greedy, gp = orarm.sample(sd, cfg["n_heads"], c, x0, ctx, 4, top_k=None, guidance_scale=1.0)
    assert torch.equal(greedy, gp.argmax(-1).t())


def test_sampling_loop_matches_reference_code():
    """tests/golden/ref_rarm_sampling.npz: the REFERENCE's LatentImageRETRO.sample_from_rdata / sample run end to end (retrieval, SOS
    conditioning, guidance on the logits, temperature, top-k, softmax, decode) with `torch.multinomial` swapped for the inverse-CDF draw
    on recorded uniforms.  The oracle loop must hand the draw the same probabilities at every step and end in the same images."""
    from oracle import knn as oknn
    import make_golden_ref as gen
    import retro_stub
    g = np.load(os.path.join(GOLD, "ref_rarm_sampling.npz"))
    cfg = gen.RARM_CFG
    sd = ref_weights.state_dict_for(orarm.param_shapes(**cfg).items(), 21)
    db = ref_weights.make_db(int(g["n_db"]))[0][:, :128].copy()
    fs = retro_stub.StubFirstStage(n_embed=48, embed_dim=8)
    c, x0 = torch.full((2, 1), 49), torch.zeros((2, 0), dtype=torch.long)
    for tag, kw in gen.RARM_SAMPLING_CASES.items():
        q = db[g[f"{tag}:qids"]].astype(np.float32)
        nns, _ = oknn.search(db, oknn.normalize_queries(q), 4)
        r = torch.from_numpy(db[nns].astype(np.float32))
        toks, probs = orarm.sample(sd, cfg["n_heads"], c, x0, r, 9, kw["temperature"], kw["top_k"], kw["guidance_scale"], torch.from_numpy(g[f"{tag}:uniforms"]))
        assert float((probs - torch.from_numpy(g[f"{tag}:probs"])).abs().max()) < 5e-6, tag      # fp32 rounding, amplified by the guidance scale
        with torch.no_grad():
            img = fs.decode(fs.quantize.get_codebook_entry(toks.reshape(-1), shape=(2, 3, 3, 8)))
        assert torch.equal(img, torch.from_numpy(g[f"{tag}:images"])), tag
    r, start = torch.from_numpy(g["greedy:r"]), torch.from_numpy(g["greedy:start"])
    toks, _ = orarm.sample(sd, cfg["n_heads"], c, start, r, 6, guidance_scale=3.0)
    assert torch.equal(toks, torch.from_numpy(g["greedy:tokens"]))
