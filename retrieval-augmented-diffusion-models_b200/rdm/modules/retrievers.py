"""`rdm.modules.retrievers`: CLIP retriever wrappers (`rdm/modules/retrievers.py:67-117`).

The CLIP ViT-B/32 encoders are the second compute sink of the north star.  They are NOT
built yet in this round: constructing the wrappers works (so configs instantiate and DB-row / pre-embedded sampling runs),
calling them raises instead of silently falling back to a PyTorch implementation.
"""
import torch
import torch.nn as nn


class _NotBuiltCLIP(nn.Module):
    def encode_image(self, *a, **k):
        raise NotImplementedError("CLIP image encode on librdm_b200 is not built yet (SURVEY.md section 8a row a20)")

    encode_text = encode_image


class ClipImageRetriever(nn.Module):
    def __init__(self, model, jit=False, device='cuda' if torch.cuda.is_available() else 'cpu', antialias=False):
        super().__init__()
        self.model_name, self.antialias = model, antialias
        self.model = _NotBuiltCLIP()
        self.register_buffer('mean', torch.Tensor([0.48145466, 0.4578275, 0.40821073]), persistent=False)
        self.register_buffer('std', torch.Tensor([0.26862954, 0.26130258, 0.27577711]), persistent=False)

    def forward(self, x):
        return self.model.encode_image(x)


class CLIPTextEmbedder(nn.Module):
    def __init__(self, model="ViT-B/32", device="cuda", add_k_shape=False):
        super().__init__()
        self.model, self.device, self.add_k_shape = _NotBuiltCLIP(), device, add_k_shape

    def forward(self, txt):
        return self.model.encode_text(txt)
