"""`ldm.models.autoencoder.VQModelInterface` stand-in: the DECODE side only (quantise -> post_quant_conv -> Decoder), as
described in SURVEY.md Appendix A, so `decode_first_stage` (ddpm.py:840,981) returns images for the unchanged sampling scripts.
The nn.Modules below are parameter containers only (checkpoint key layout, no forward); `decode` runs the hand-written
decoder of librdm_b200 (`rdm_b200.vqdecoder.B200VQDecoder`) and raises for tensors that are not on a CUDA device.  Module names follow the latent-diffusion checkpoint layout (`first_stage_model.decoder.*`,
`first_stage_model.quantize.embedding.weight`, `first_stage_model.post_quant_conv.*`)."""
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



class AttnBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.norm = _norm(c)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(c, c, 1) for _ in range(4))



class Upsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)



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
        self._ddconfig = dict(ddconfig)

    def encode(self, x):
        raise NotImplementedError("the first-stage ENCODER is outside the sampling hot path (SURVEY.md section 2)")

    def _b200(self, device):
        """Decoder handle on `device`, re-synchronised whenever a parameter tensor was written (load_state_dict, .to, ...)."""
        from rdm_b200.vqdecoder import B200VQDecoder
        stamp = (str(device), tuple((p.data_ptr(), p._version) for p in self.parameters()))
        if getattr(self, "_b200_stamp", None) != stamp:
            dec = B200VQDecoder(device, self.embed_dim, self.quantize.embedding.num_embeddings, self._ddconfig)
            dec.load_state_dict(self.state_dict())
            self._b200_dec, self._b200_stamp = dec, stamp
        return self._b200_dec

    def decode(self, h, force_not_quantize=False):
        if not h.is_cuda:
            raise RuntimeError("the first-stage decoder (B200 build) has no CPU path: move the model and the latents to a CUDA device")
        return self._b200(h.device).decode(h, force_not_quantize)
