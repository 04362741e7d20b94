"""`rdm.modules.retrievers`: CLIP retriever wrappers (`rdm/modules/retrievers.py:67-117`) on the librdm_b200 CLIP executor.

`load_clip(name)` replaces `clip.load` (`retrievers.py:6,76`): the ViT-B/32 weights are looked up LAZILY (first encode) at
`$RDM_CLIP_CKPT`, `~/.cache/clip/ViT-B-32.pt` (OpenAI's download location; TorchScript archive or plain state dict) -- there
is no network here and sampling from database rows never needs CLIP.  A state dict can also be injected with
`model.load_state_dict(sd)`.  No PyTorch compute path exists.
"""
import os

import torch
import torch.nn as nn

from rdm_b200.clip import VIT_B32, B200Clip, cfg_from_state_dict

_NAMES = {"ViT-B/32": "ViT-B-32.pt"}


class B200ClipModel(nn.Module):
    """Quacks like the reference `CLIP` module for the calls the sampling path makes: encode_image / encode_text."""

    def __init__(self, name="ViT-B/32", device="cuda"):
        super().__init__()
        self.name, self._device, self._engine, self._sd = name, torch.device(device if device != "cuda" else "cuda"), None, None
        self.register_buffer("_dev_probe", torch.zeros(1), persistent=False)

    def load_state_dict(self, sd, strict=True):
        self._sd = {k: v for k, v in sd.items() if isinstance(v, torch.Tensor)}
        self._engine = None
        return [], []

    def _find_weights(self):
        cands = [os.environ.get("RDM_CLIP_CKPT"), os.path.expanduser(os.path.join("~/.cache/clip", _NAMES.get(self.name, self.name.replace("/", "-") + ".pt")))]
        for c in cands:
            if c and os.path.isfile(c):
                try:
                    return torch.jit.load(c, map_location="cpu").state_dict()
                except Exception:
                    sd = torch.load(c, map_location="cpu")
                    return sd.get("state_dict", sd)
        raise FileNotFoundError(f"CLIP weights for {self.name} not found (looked at {cands}); set RDM_CLIP_CKPT or call load_state_dict()")

    def engine(self):
        dev = self._dev_probe.device
        if dev.type != "cuda":
            raise RuntimeError("CLIP (B200 build) has no CPU path: move the retriever to a CUDA device")
        if self._engine is None or self._engine.device != dev:
            sd = self._sd if self._sd is not None else self._find_weights()
            self._sd = sd
            self._engine = B200Clip(dev, **cfg_from_state_dict(sd))
            self._engine.load_state_dict(sd)
        return self._engine

    @torch.no_grad()
    def encode_image(self, image):
        return self.engine().encode_image(image)

    @torch.no_grad()
    def encode_text(self, text):
        return self.engine().encode_text(text)


def load_clip(name="ViT-B/32", device="cuda", jit=False):
    model = B200ClipModel(name, device)
    if str(device).startswith("cuda") and torch.cuda.is_available():
        model = model.to(device)
    preprocess = lambda x: model.engine().preprocess(x)
    return model, preprocess


class ClipImageRetriever(nn.Module):
    def __init__(self, model, jit=False, device='cuda' if torch.cuda.is_available() else 'cpu', antialias=False):
        super().__init__()
        assert not antialias, "antialiased resize is not implemented"
        self.model, _ = load_clip(name=model, device=device, jit=jit)
        self.antialias = antialias
        self.register_buffer('mean', torch.Tensor([0.48145466, 0.4578275, 0.40821073]), persistent=False)
        self.register_buffer('std', torch.Tensor([0.26862954, 0.26130258, 0.27577711]), persistent=False)

    def preprocess(self, x):
        return self.model.engine().preprocess(x)              # bicubic 224 (align_corners) + (x+1)/2 + mean/std, one kernel

    def forward(self, x):
        return self.model.encode_image(self.preprocess(x))    # x in [-1, 1]  (retrievers.py:93-95)


class CLIPTextEmbedder(nn.Module):
    def __init__(self, model="ViT-B/32", device="cuda", add_k_shape=False):
        super().__init__()
        self.model, _ = load_clip(model, device=device)
        self.device, self.add_k_shape = device, add_k_shape

    def preprocess(self, text):
        from rdm.modules.custom_clip.clip import tokenize          # retrievers.py:9,111: the vendored tokenizer (cuts over-long captions)
        return tokenize(text)

    def forward(self, txt):
        emb = self.model.encode_text(self.preprocess(txt).to(self.model._dev_probe.device))
        return emb[:, None] if self.add_k_shape else emb
