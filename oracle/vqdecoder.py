"""ORACLE (test infrastructure; imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs).

torch-CPU fp32 restatement of the first-stage decode path `VQModelInterface.decode` = VectorQuantizer lookup -> post_quant_conv ->
Decoder.  The source of these modules is NOT in /root/reference (un-vendored dependency `latent-diffusion@main`,
`ldm/modules/diffusionmodules/model.py` + `taming/modules/vqvae/quantize.py`); the call sites this
follows are rdm/models/diffusion/ddpm.py:840,981 (`decode_first_stage`) and the configuration models/rdm/imagenet/config.yaml:60-80.
Parity unpinned by the reference (no fixtures); module / key names follow the latent-diffusion checkpoint layout."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _norm(c):
    return nn.GroupNorm(32, c, eps=1e-6, affine=True)


class ResnetBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.norm1, self.conv1 = _norm(cin), nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2, self.conv2 = _norm(cout), nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.nin_shortcut = nn.Conv2d(cin, cout, 1)
        self.cin, self.cout = cin, cout

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (self.nin_shortcut(x) if self.cin != self.cout else x) + h


class AttnBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.norm = _norm(c)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(c, c, 1) for _ in range(4))

    def forward(self, x):
        b, c, h, w = x.shape
        n = self.norm(x)
        q, k, v = self.q(n).reshape(b, c, h * w), self.k(n).reshape(b, c, h * w), self.v(n).reshape(b, c, h * w)
        o = F.scaled_dot_product_attention(q.transpose(1, 2)[:, None], k.transpose(1, 2)[:, None], v.transpose(1, 2)[:, None])[:, 0]
        return x + self.proj_out(o.transpose(1, 2).reshape(b, c, h, w))


class Upsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class Decoder(nn.Module):
    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, resolution, z_channels, **ignored):
        super().__init__()
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        block_in, curr_res = ch * ch_mult[-1], resolution // 2 ** (len(ch_mult) - 1)
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, padding=1)
        self.mid = nn.Module()
        self.mid.block_1, self.mid.attn_1, self.mid.block_2 = ResnetBlock(block_in, block_in), AttnBlock(block_in), ResnetBlock(block_in, block_in)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn, block_out = nn.ModuleList(), nn.ModuleList(), ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(block_in, block_out)); block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            up = nn.Module(); up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = Upsample(block_in); curr_res *= 2
            self.up.insert(0, up)
        self.norm_out, self.conv_out = _norm(block_in), nn.Conv2d(block_in, out_ch, 3, padding=1)

    def forward(self, z):
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(self.conv_in(z))))
        for i_level in reversed(range(self.num_resolutions)):
            for i_block in range(self.num_res_blocks + 1):
                h = self.up[i_level].block[i_block](h)
                if len(self.up[i_level].attn) > 0:
                    h = self.up[i_level].attn[i_block](h)
            if i_level != 0:
                h = self.up[i_level].upsample(h)
        return self.conv_out(F.silu(self.norm_out(h)))


class VectorQuantizer(nn.Module):
    def __init__(self, n_e, e_dim):
        super().__init__()
        self.embedding = nn.Embedding(n_e, e_dim)

    def forward(self, z):
        zf = z.permute(0, 2, 3, 1).reshape(-1, z.shape[1])
        d = (zf ** 2).sum(1, keepdim=True) + (self.embedding.weight ** 2).sum(1)[None] - 2 * zf @ self.embedding.weight.t()
        idx = d.argmin(1)
        zq = self.embedding(idx).view(z.shape[0], z.shape[2], z.shape[3], -1).permute(0, 3, 1, 2).contiguous()
        return zq, None, (None, None, idx)


class VQModelInterface(nn.Module):
    def __init__(self, embed_dim, n_embed, ddconfig, lossconfig=None, **ignored):
        super().__init__()
        self.decoder = Decoder(**ddconfig)
        self.quantize = VectorQuantizer(n_embed, embed_dim)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.embed_dim = embed_dim

    def decode(self, h, force_not_quantize=False):
        quant = h if force_not_quantize else self.quantize(h)[0]
        return self.decoder(self.post_quant_conv(quant))


TINY_VQ = dict(embed_dim=3, n_embed=512, ddconfig=dict(double_z=False, z_channels=3, resolution=64, in_channels=3, out_ch=3, ch=64,
                                                         ch_mult=[1, 2], num_res_blocks=1, attn_resolutions=[], dropout=0.0))
RDM_VQ_F4 = dict(embed_dim=3, n_embed=8192, ddconfig=dict(double_z=False, z_channels=3, resolution=256, in_channels=3, out_ch=3, ch=128,
                                                          ch_mult=[1, 2, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0))
"""models/rdm/imagenet/config.yaml:60-80 (VQ-f4: 64x64x3 latent -> 256x256x3 image)."""


def randomize_(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if p.dim() >= 2 and "embedding" not in name:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / fan_in ** 0.5)
            elif "embedding" in name:
                p.copy_(torch.randn(p.shape, generator=g))
            elif name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    return m


TAMING_VQ_F16 = dict(embed_dim=256, n_embed=16384, ddconfig=dict(double_z=False, z_channels=256, resolution=256, in_channels=3, out_ch=3, ch=128,
                                                                  ch_mult=[1, 1, 2, 2, 4], num_res_blocks=2, attn_resolutions=[16], dropout=0.0))
"""models/rarm/imagenet/dogs/config.yaml:28-51 (taming VQGAN-f16 of the RARM models: 16x16 codes of width 256 -> 256x256x3 image; same Decoder
architecture as above, AttnBlocks after every ResnetBlock of the 16x16 level)."""
TINY_VQ_WIDE = dict(embed_dim=64, n_embed=96, ddconfig=dict(double_z=False, z_channels=64, resolution=32, in_channels=3, out_ch=3, ch=64,
                                                             ch_mult=[1, 2], num_res_blocks=1, attn_resolutions=[16], dropout=0.0))


def decode_indices(model, indices, zshape):
    """taming `Net2NetTransformer.decode_to_img` (called from rdm/models/autoregression/transformer.py:291-292) with the identity permuter:
    codebook entries of the sampled ids, reshaped (b, h, w, c) -> NCHW, then post_quant_conv -> Decoder (no re-quantisation)."""
    b, c, h, w = zshape
    quant = model.quantize.embedding(indices.reshape(-1)).view(b, h, w, c).permute(0, 3, 1, 2).contiguous()
    return model.decode(quant, force_not_quantize=True)
