"""Oracle U-Net: torch-CPU fp32 restatement of the reference's cross-attending U-Net.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
``rdm/modules/diffusionmodules/openaimodel.py:66-317`` (layer plan) and ``:335-371``
(forward), ``rdm/modules/attention.py:20-74`` (CrossAttention), ``:77-96``
(BasicTransformerBlock), ``:122-196`` (SpatialTransformer) and, for the pieces
the reference imports from un-vendored ``latent-diffusion@main``, SURVEY.md
Appendix A (ResBlock, Downsample, Upsample, GroupNorm32, timestep_embedding,
GEGLU FeedForward).  Sub-module names reproduce the checkpoint key layout of
SURVEY.md Appendix C so state dicts are interchangeable with the product's
``rdm.modules.diffusionmodules.openaimodel.UNetModel``.

Only the configuration space the shipped configs use is restated
(``use_spatial_transformer=True``, ``resblock_updown=False``,
``use_scale_shift_norm=False``, ``dims=2``, ``transformer_depth=1``).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def timestep_embedding(t, dim, max_period=10000):
    """ldm ``timestep_embedding``: [cos | sin] of t * freqs."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


class GroupNorm32(nn.GroupNorm):
    """ldm ``normalization(C)``: GroupNorm(32, C, eps=1e-5) evaluated in fp32."""

    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


class ResBlock(nn.Module):
    """ldm ResBlock without scale-shift / updown (openaimodel.py call sites :160-168,:225-232,:258-266)."""

    def __init__(self, channels, emb_channels, out_channels=None):
        super().__init__()
        out_channels = out_channels or channels
        self.channels, self.out_channels = channels, out_channels
        self.in_layers = nn.Sequential(GroupNorm32(32, channels), nn.SiLU(),
                                       nn.Conv2d(channels, out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, out_channels))
        self.out_layers = nn.Sequential(GroupNorm32(32, out_channels), nn.SiLU(), nn.Dropout(0.0),
                                        nn.Conv2d(out_channels, out_channels, 3, padding=1))
        if out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = nn.Conv2d(channels, out_channels, 1)

    def forward(self, x, emb):
        h = self.in_layers(x)
        h = h + self.emb_layers(emb)[:, :, None, None]
        h = self.out_layers(h)
        return self.skip_connection(x) + h


class Downsample(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.op = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        return self.op(x)


class Upsample(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2, mode="nearest"))


class CrossAttention(nn.Module):
    """attention.py:20-74 without mask/causal (never set on the U-Net path)."""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64):
        super().__init__()
        inner = heads * dim_head
        context_dim = query_dim if context_dim is None else context_dim
        self.heads, self.scale = heads, dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(0.0))

    def forward(self, x, context=None):
        context = x if context is None else context
        b, n, _ = x.shape
        h = self.heads
        q, k, v = self.to_q(x), self.to_k(context), self.to_v(context)
        split = lambda t: t.reshape(b, t.shape[1], h, -1).permute(0, 2, 1, 3)
        q, k, v = split(q), split(k), split(v)
        sim = torch.matmul(q, k.transpose(-1, -2)) * self.scale
        attn = sim.softmax(dim=-1)
        out = torch.matmul(attn, v).permute(0, 2, 1, 3).reshape(b, n, -1)
        return self.to_out(out)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        a, g = self.proj(x).chunk(2, dim=-1)
        return a * F.gelu(g)


class FeedForward(nn.Module):
    """ldm FeedForward(glu=True, mult=4)."""

    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.Sequential(GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim))

    def forward(self, x):
        return self.net(x)


class BasicTransformerBlock(nn.Module):
    """attention.py:77-96."""

    def __init__(self, dim, n_heads, d_head, context_dim=None):
        super().__init__()
        self.attn1 = CrossAttention(dim, heads=n_heads, dim_head=d_head)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, context_dim=context_dim, heads=n_heads, dim_head=d_head)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)

    def forward(self, x, context=None):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context=context) + x
        x = self.ff(self.norm3(x)) + x
        return x


class SpatialTransformer(nn.Module):
    """attention.py:122-196 (dims=2)."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, context_dim=None):
        super().__init__()
        inner = n_heads * d_head
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6, affine=True)   # attention.py:16-17
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, n_heads, d_head, context_dim=context_dim) for _ in range(depth)])
        self.proj_out = nn.Conv2d(inner, in_channels, 1)

    def forward(self, x, context=None):
        if isinstance(context, (list, tuple)):
            assert len(context) == 1
            context = context[0]
        b, c, h, w = x.shape
        x_in = x
        x = self.proj_in(self.norm(x))
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, -1)
        for blk in self.transformer_blocks:
            x = blk(x, context=context)
        x = x.reshape(b, h, w, -1).permute(0, 3, 1, 2)
        return self.proj_out(x) + x_in


class TimestepEmbedSequential(nn.Sequential):
    """openaimodel.py:17-33: route emb to ResBlocks, context to SpatialTransformers."""

    def forward(self, x, emb, context=None):
        for layer in self:
            if isinstance(layer, ResBlock):
                x = layer(x, emb)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, context)
            else:
                x = layer(x)
        return x


class UNetModel(nn.Module):
    """openaimodel.py:66-317 (plan) and :335-371 (forward) for the shipped configuration space."""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks,
                 attention_resolutions, channel_mult=(1, 2, 4, 8), num_head_channels=32,
                 transformer_depth=1, context_dim=512, **ignored):
        super().__init__()
        self.in_channels, self.out_channels, self.model_channels = in_channels, out_channels, model_channels
        ted = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, ted), nn.SiLU(), nn.Linear(ted, ted))
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(in_channels, model_channels, 3, padding=1))])
        chans, ch, ds = [model_channels], model_channels, 1
        self.head_counts = []

        def st(c):
            self.head_counts.append(c // num_head_channels)
            return SpatialTransformer(c, c // num_head_channels, num_head_channels,
                                      depth=transformer_depth, context_dim=context_dim)

        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [ResBlock(ch, ted, mult * model_channels)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(st(ch))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch)))
                chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(ResBlock(ch, ted), st(ch), ResBlock(ch, ted))
        self.output_blocks = nn.ModuleList()
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [ResBlock(ch + chans.pop(), ted, model_channels * mult)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(st(ch))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(GroupNorm32(32, ch), nn.SiLU(), nn.Conv2d(model_channels, out_channels, 3, padding=1))

    def forward(self, x, timesteps, context=None):
        emb = self.time_embed(timestep_embedding(timesteps, self.model_channels))
        hs, h = [], x
        for m in self.input_blocks:
            h = m(h, emb, context)
            hs.append(h)
        h = self.middle_block(h, emb, context)
        for m in self.output_blocks:
            h = m(torch.cat([h, hs.pop()], dim=1), emb, context)
        return self.out(h)


def randomize_(model, seed=0, gain=1.0):
    """Seeded random weights that keep activations O(1) and leave NO layer at the reference's
    zero-init (``zero_module`` convs would make parity tests uninformative)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.ndim >= 2:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * (gain / math.sqrt(fan_in)))
            elif name.endswith("weight"):        # norm scales
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            else:                                # biases / norm shifts
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    return model


IMAGENET_UNET = dict(image_size=64, in_channels=3, out_channels=3, model_channels=192,
                     attention_resolutions=[8, 4, 2], num_res_blocks=2, channel_mult=[1, 2, 3, 5],
                     num_head_channels=32, transformer_depth=1, context_dim=512)
"""``models/rdm/imagenet/config.yaml:36-59``."""

BASELINE_UNET = dict(IMAGENET_UNET, image_size=32, in_channels=4, out_channels=4)
"""BASELINE.json cfg2: same widths on the 32x32x4 (VQ-f8) latent."""

TINY_UNET = dict(image_size=16, in_channels=4, out_channels=4, model_channels=64,
                 attention_resolutions=[2, 4], num_res_blocks=1, channel_mult=[1, 2, 3],
                 num_head_channels=32, transformer_depth=1, context_dim=512)
"""Small arch for fast parity tests: exercises every layer kind incl. non-aligned concat GN groups."""
