"""`taming.models.vqgan.VQModel` stand-in: the DECODE side only, for the RARM first stage
(`models/rarm/imagenet/*/config.yaml:28-51`, used by taming's `Net2NetTransformer.decode_to_img`, which
`rdm/models/autoregression/transformer.py:291-292` calls after sampling): `quantize.get_codebook_entry(indices, shape)` ->
`post_quant_conv` -> `Decoder`.  taming-transformers is not vendored by the reference; the decoder architecture is the one
latent-diffusion inherited from it, so the parameter containers of the `ldm` stand-in are reused (same checkpoint key layout:
`first_stage_model.decoder.*`, `first_stage_model.quantize.embedding.weight`, `first_stage_model.post_quant_conv.*`).
`decode` runs the hand-written decoder of librdm_b200 (wide-latent path of csrc/unet.cu) and raises for tensors that are not on a
CUDA device.  Appended to sys.path only when the real package is missing (rdm_b200/compat.py)."""
import torch.nn as nn

from ldm.models.autoencoder import Decoder, VQModelInterface


class VectorQuantizer(nn.Module):
    """taming `VectorQuantizer2`, lookup side (remap=None)."""

    def __init__(self, n_e, e_dim):
        super().__init__()
        self.n_e, self.e_dim = n_e, e_dim
        self.embedding = nn.Embedding(n_e, e_dim)

    def get_codebook_entry(self, indices, shape):
        z_q = self.embedding(indices)
        if shape is not None:
            z_q = z_q.view(shape).permute(0, 3, 1, 2).contiguous()          # (batch, height, width, channel) -> NCHW
        return z_q

    def forward(self, z):
        raise NotImplementedError("the first-stage ENCODER / quantiser search is outside the sampling hot path (SURVEY.md section 2)")


class VQModel(VQModelInterface):
    def __init__(self, ddconfig, lossconfig=None, n_embed=None, embed_dim=None, ckpt_path=None, ignore_keys=(), image_key="image", colorize_nlabels=None,
                 monitor=None, remap=None, sane_index_shape=False, **ignored):
        assert remap is None, "remapped codebooks are not used by the shipped RARM configs"
        super().__init__(embed_dim, n_embed, ddconfig, lossconfig)
        self.quantize = VectorQuantizer(n_embed, embed_dim)
        self.image_key = image_key

    def decode(self, quant):
        """quant: codebook entries NCHW [B, embed_dim, h, w] (taming `VQModel.decode`: post_quant_conv -> decoder; no re-quantisation)."""
        return super().decode(quant, force_not_quantize=True)
