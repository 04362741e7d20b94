"""Oracle RARM decoder: torch-CPU fp32 restatement.  TEST INFRASTRUCTURE (oracle/__init__.py).

Follows the reference's `rdm/modules/attention.py`: `RetrievalPatchTransformer` (:199-272) as configured by
`models/rarm/imagenet/*/config.yaml:14-27` (`continuous: false` -> `proj_in` is an nn.Embedding over the 16386-entry input vocabulary,
`positional_encoding` [inner_dim, sequence_length] added column-wise, `depth` x `BasicTransformerBlock` (:77-96) with a CAUSAL
self-attention, a non-causal cross-attention to the k retrieved CLIP vectors (`CrossAttention` :20-74, causal mask :58-65) and the
ldm GEGLU feed-forward, `proj_out` = Conv1d(inner_dim, out_channels, 1); no final LayerNorm), and the sampling
loop `LatentImageRETRO.sample` (`rdm/models/autoregression/transformer.py:224-270`): classifier-free guidance on the LOGITS with an
all-zeros retrieval context (:233-253), temperature, top-k filtering (taming `top_k_logits`: everything below the k-th largest logit
becomes -inf), softmax, one draw per step.

PINNED: `tests/golden/ref_rarm_small.npz` holds logits computed by the REFERENCE's own `RetrievalPatchTransformer`
(tests/golden/make_golden_ref.py); `tests/test_oracle_rarm.py` checks `forward` against it and `forward_incremental`
(the KV-cache formulation the CUDA executor uses) against `forward`.

Sampling draw: `torch.multinomial` consumes the global (CUDA Philox) generator and cannot be reproduced by a device kernel, so the
draw is defined on explicit uniforms: token = first index (ascending) whose cumulative probability exceeds u.  Same distribution.
Parameters are addressed by the reference's state-dict names below `transformer.` (e.g. `transformer_blocks.0.attn1.to_q.weight`).
"""
import torch
import torch.nn.functional as F


def _ln(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _heads(t, h):
    B, N, C = t.shape
    return t.reshape(B, N, h, C // h).transpose(1, 2)           # [B, h, N, d]


def _attention(q, k, v, heads, causal):
    """CrossAttention.forward (attention.py:42-74) after the projections; q [B,Nq,C], k/v [B,Nk,C]."""
    q, k, v = _heads(q, heads), _heads(k, heads), _heads(v, heads)
    sim = (q @ k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)
    if causal:                                                   # :58-65: query i sees keys j <= i + (Nk - Nq)
        i, j = sim.shape[-2:]
        mask = torch.ones(i, j, dtype=torch.bool).triu(j - i + 1)
        sim = sim.masked_fill(mask, -torch.finfo(sim.dtype).max)
    out = sim.softmax(dim=-1) @ v
    return out.transpose(1, 2).reshape(out.shape[0], out.shape[2], -1)


def _block(x, ctx, sd, p, heads):
    """BasicTransformerBlock._forward (attention.py:92-96)."""
    n = _ln(x, sd, p + "norm1")
    a = _attention(n @ sd[p + "attn1.to_q.weight"].t(), n @ sd[p + "attn1.to_k.weight"].t(), n @ sd[p + "attn1.to_v.weight"].t(), heads, True)
    x = a @ sd[p + "attn1.to_out.0.weight"].t() + sd[p + "attn1.to_out.0.bias"] + x
    n = _ln(x, sd, p + "norm2")
    a = _attention(n @ sd[p + "attn2.to_q.weight"].t(), ctx @ sd[p + "attn2.to_k.weight"].t(), ctx @ sd[p + "attn2.to_v.weight"].t(), heads, False)
    x = a @ sd[p + "attn2.to_out.0.weight"].t() + sd[p + "attn2.to_out.0.bias"] + x
    n = _ln(x, sd, p + "norm3")
    h = n @ sd[p + "ff.net.0.proj.weight"].t() + sd[p + "ff.net.0.proj.bias"]
    val, gate = h.chunk(2, dim=-1)
    return (val * F.gelu(gate)) @ sd[p + "ff.net.2.weight"].t() + sd[p + "ff.net.2.bias"] + x


def depth_of(sd):
    i = 0
    while f"transformer_blocks.{i}.norm1.weight" in sd:
        i += 1
    return i


@torch.no_grad()
def forward(sd, tokens, context, heads):
    """RetrievalPatchTransformer.forward (attention.py:247-272), discrete input: tokens int64 [B,T], context [B,k,ctx] -> [B,T,out]."""
    x = sd["proj_in.weight"][tokens]                                              # Embedding; 'b t c'
    x = x + sd["positional_encoding"][:, :x.shape[1]].t()[None]
    for i in range(depth_of(sd)):
        x = _block(x, context, sd, f"transformer_blocks.{i}.", heads)
    return x @ sd["proj_out.weight"][:, :, 0].t() + sd["proj_out.bias"]


@torch.no_grad()
def forward_incremental(sd, tokens, context, heads):
    """Same function evaluated one position at a time with per-layer key/value caches (what the CUDA executor does):
    returns logits [B,T,out]; row t only ever touches cached rows 0..t."""
    B, T = tokens.shape
    L = depth_of(sd)
    kc = [None] * L
    vc = [None] * L
    ck = [context @ sd[f"transformer_blocks.{i}.attn2.to_k.weight"].t() for i in range(L)]     # step-invariant: projected once
    cv = [context @ sd[f"transformer_blocks.{i}.attn2.to_v.weight"].t() for i in range(L)]
    outs = []
    for t in range(T):
        x = sd["proj_in.weight"][tokens[:, t]][:, None] + sd["positional_encoding"][:, t][None, None]
        for i in range(L):
            p = f"transformer_blocks.{i}."
            n = _ln(x, sd, p + "norm1")
            k_new, v_new = n @ sd[p + "attn1.to_k.weight"].t(), n @ sd[p + "attn1.to_v.weight"].t()
            kc[i] = k_new if t == 0 else torch.cat([kc[i], k_new], 1)
            vc[i] = v_new if t == 0 else torch.cat([vc[i], v_new], 1)
            a = _attention(n @ sd[p + "attn1.to_q.weight"].t(), kc[i], vc[i], heads, False)       # all cached keys are <= t
            x = a @ sd[p + "attn1.to_out.0.weight"].t() + sd[p + "attn1.to_out.0.bias"] + x
            n = _ln(x, sd, p + "norm2")
            a = _attention(n @ sd[p + "attn2.to_q.weight"].t(), ck[i], cv[i], heads, False)
            x = a @ sd[p + "attn2.to_out.0.weight"].t() + sd[p + "attn2.to_out.0.bias"] + x
            n = _ln(x, sd, p + "norm3")
            val, gate = (n @ sd[p + "ff.net.0.proj.weight"].t() + sd[p + "ff.net.0.proj.bias"]).chunk(2, dim=-1)
            x = (val * F.gelu(gate)) @ sd[p + "ff.net.2.weight"].t() + sd[p + "ff.net.2.bias"] + x
        outs.append(x @ sd["proj_out.weight"][:, :, 0].t() + sd["proj_out.bias"])
    return torch.cat(outs, 1)


def top_k_logits(logits, k):
    """taming Net2NetTransformer.top_k_logits (called at transformer.py:258): entries below the k-th largest value -> -inf (ties kept)."""
    v, _ = torch.topk(logits, k)
    out = logits.clone()
    out[out < v[..., [-1]]] = -float("inf")
    return out


def step_probs(logits_cond, logits_uncond, guidance_scale, temperature, top_k):
    """transformer.py:249-260 for the last position: guided logits / temperature -> top-k -> softmax.  float32 like the reference."""
    logits = logits_cond if logits_uncond is None else logits_uncond + guidance_scale * (logits_cond - logits_uncond)
    logits = logits / temperature
    if top_k is not None:
        logits = top_k_logits(logits, top_k)
    return F.softmax(logits, dim=-1)


def draw(probs, u):
    """First index whose cumulative probability (float64, ascending index) exceeds u * total; u in [0,1).  probs [B,V], u [B]."""
    cdf = probs.double().cumsum(-1)
    target = u.double() * cdf[:, -1]
    idx = (cdf > target[:, None]).float().argmax(-1)
    return idx


@torch.no_grad()
def sample(sd, heads, c, x, r, steps, temperature=1.0, top_k=None, guidance_scale=1.0, uniforms=None):
    """LatentImageRETRO.sample (transformer.py:224-270): c [B,Tc] conditioning tokens (the sos token), x [B,T0] start tokens,
    r [B,k,ctx] retrieval context.  uniforms [steps,B] -> inverse-CDF draws; None -> greedy (`sample=False`).
    Recomputes the whole prefix every step, like the reference.  Returns (tokens [B, T0+steps], probs of every step [steps,B,V])."""
    x = torch.cat((c, x), 1)
    bs = x.shape[0]
    if guidance_scale > 1.0:
        r = torch.cat((r, torch.zeros_like(r)), dim=0)
    all_probs = []
    for k in range(steps):
        xx = torch.cat((x, x), dim=0) if guidance_scale > 1.0 else x
        logits = forward(sd, xx, r, heads)[:, -1, :]
        lc, lu = (logits[:bs], logits[bs:]) if guidance_scale > 1.0 else (logits, None)
        probs = step_probs(lc, lu, guidance_scale, temperature, top_k)
        ix = draw(probs, uniforms[k]) if uniforms is not None else probs.argmax(-1)
        all_probs.append(probs)
        x = torch.cat((x, ix[:, None]), dim=1)
    return x[:, c.shape[1]:], torch.stack(all_probs)


def param_shapes(in_channels, n_heads, d_head, depth, context_dim, sequence_length, out_channels, **_):
    """name -> shape in the reference's registration order (attention.py:224-245, :79-87, :24-38; ldm FeedForward)."""
    C = n_heads * d_head
    P = {"positional_encoding": (C, sequence_length), "proj_in.weight": (in_channels, C)}
    for i in range(depth):
        p = f"transformer_blocks.{i}."
        for a, cd in (("attn1", C), ("attn2", context_dim)):
            if a == "attn2":
                P[p + "ff.net.0.proj.weight"], P[p + "ff.net.0.proj.bias"] = (8 * C, C), (8 * C,)
                P[p + "ff.net.2.weight"], P[p + "ff.net.2.bias"] = (C, 4 * C), (C,)
            P[p + a + ".to_q.weight"], P[p + a + ".to_k.weight"], P[p + a + ".to_v.weight"] = (C, C), (C, cd), (C, cd)
            P[p + a + ".to_out.0.weight"], P[p + a + ".to_out.0.bias"] = (C, C), (C,)
        for nm in ("norm1", "norm2", "norm3"):
            P[p + nm + ".weight"], P[p + nm + ".bias"] = (C,), (C,)
    P["proj_out.weight"], P["proj_out.bias"] = (out_channels, C, 1), (out_channels,)
    return P


RARM_IMAGENET = dict(in_channels=16386, n_heads=12, d_head=64, depth=18, context_dim=512, sequence_length=256, out_channels=16384)
"""`models/rarm/imagenet/{dogs,mammals,animals}/config.yaml:14-27`."""
