"""`rdm.modules.encoders.nn_encoders`: only the pass-through encoders the shipped configs reference
(`rdm/modules/encoders/nn_encoders.py:127-145`).  VQ-based neighbour encoders are out of scope (unused by shipped configs)."""
import torch.nn as nn


class IdentityEncoder(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, x, *args, **kwargs):
        return x

    def encode(self, x, *args, **kwargs):
        return x


class CLIPEmbeddingReshaper(nn.Module):
    def forward(self, x, *args, **kwargs):
        return x.reshape(x.shape[0], -1, x.shape[-1]) if x.ndim > 3 else x


class VQGANAggregator(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("VQ-based neighbour encoders are outside the sampling hot path (SURVEY.md section 2)")


class VQGANNNAttender(VQGANAggregator):
    pass
