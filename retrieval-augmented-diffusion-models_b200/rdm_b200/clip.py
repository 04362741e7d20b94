"""Host wrapper of the CLIP executor in librdm_b200 (csrc/clip.cu): `encode_text` / `encode_image` / `preprocess` with the
semantics of the reference's vendored model (`rdm/modules/custom_clip/model.py:304-320`) and retriever (`retrievers.py:83-95`)."""
import ctypes

import torch

from . import _lib

VIT_B32 = dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=32, context_length=77,
               vocab_size=49408, transformer_width=512, transformer_heads=8, transformer_layers=12)
"""`clip.load("ViT-B/32")` architecture (custom_clip/model.py:363-399; 151.28 M parameters)."""


class ClipCfg(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in ("embed_dim", "image_resolution", "vision_layers", "vision_width", "vision_patch_size",
                                              "context_length", "vocab_size", "transformer_width", "transformer_heads", "transformer_layers")]


def cfg_from_state_dict(sd):
    """The reference's `build_model` shape inference (custom_clip/model.py:363-391) for ViT checkpoints."""
    vw = sd["visual.conv1.weight"].shape[0]
    P = sd["visual.conv1.weight"].shape[-1]
    G = round((sd["visual.positional_embedding"].shape[0] - 1) ** 0.5)
    tw = sd["ln_final.weight"].shape[0]
    nl = lambda p: len({k.split(".")[len(p.split(".")) + 1] for k in sd if k.startswith(p + ".resblocks.")})
    return dict(embed_dim=sd["text_projection"].shape[1], image_resolution=P * G, vision_layers=nl("visual.transformer"), vision_width=vw,
                vision_patch_size=P, context_length=sd["positional_embedding"].shape[0], vocab_size=sd["token_embedding.weight"].shape[0],
                transformer_width=tw, transformer_heads=tw // 64, transformer_layers=nl("transformer"))


class B200Clip:
    def __init__(self, device, **cfg):
        L = _lib.lib()
        self.device = _lib.resolve_device(device)
        self.cfg = dict(cfg)
        c = ClipCfg(**{k: int(cfg[k]) for k, _ in ClipCfg._fields_})
        self._h = ctypes.c_void_p()
        _lib.check(L.rdm_clip_create(ctypes.byref(self._h), ctypes.byref(c), int(self.device.index or 0)), "rdm_clip_create")
        self.names = [L.rdm_clip_param_name(self._h, i).decode() for i in range(L.rdm_clip_num_params(self._h))]

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().rdm_clip_destroy(h)
            except Exception:
                pass

    def load_state_dict(self, sd, strict=True):
        L = _lib.lib()
        missing = [k for k in self.names if k not in sd and k != "logit_scale"]
        if strict and missing:
            raise RuntimeError(f"missing CLIP parameters: {missing[:5]}")
        for k in self.names:
            if k not in sd:
                continue
            t = sd[k].detach().to("cpu", torch.float32).contiguous()
            if t.numel() != L.rdm_clip_param_numel(self._h, k.encode()):
                raise RuntimeError(f"size mismatch for {k}: {tuple(t.shape)}")
            _lib.check(L.rdm_clip_load(self._h, k.encode(), ctypes.c_void_p(t.data_ptr()), t.numel()), f"rdm_clip_load({k})")
        return missing

    def set_mode(self, mode):
        _lib.check(_lib.lib().rdm_clip_set_mode(self._h, int(mode)), "rdm_clip_set_mode")

    def encode_text(self, tokens):
        tokens = tokens.to(self.device, torch.int64).contiguous()
        assert tokens.ndim == 2 and tokens.shape[1] == self.cfg["context_length"]
        out = torch.empty((tokens.shape[0], self.cfg["embed_dim"]), dtype=torch.float32, device=self.device)
        with _lib.device_ctx(self.device):
            _lib.check(_lib.lib().rdm_clip_encode_text(self._h, _lib.ptr(tokens), tokens.shape[0], _lib.ptr(out), _lib.stream_ptr(self.device)), "rdm_clip_encode_text")
        return out

    def encode_image(self, image):
        image = image.to(self.device, torch.float32).contiguous()
        R = self.cfg["image_resolution"]
        assert tuple(image.shape[1:]) == (3, R, R), f"encode_image expects [B,3,{R},{R}] (preprocessed)"
        out = torch.empty((image.shape[0], self.cfg["embed_dim"]), dtype=torch.float32, device=self.device)
        with _lib.device_ctx(self.device):
            _lib.check(_lib.lib().rdm_clip_encode_image(self._h, _lib.ptr(image), image.shape[0], _lib.ptr(out), _lib.stream_ptr(self.device)), "rdm_clip_encode_image")
        return out

    def preprocess(self, x, size=None):
        x = x.to(self.device, torch.float32).contiguous()
        size = size or self.cfg["image_resolution"]
        out = torch.empty((x.shape[0], 3, size, size), dtype=torch.float32, device=self.device)
        with _lib.device_ctx(self.device):
            _lib.check(_lib.lib().rdm_clip_preprocess(_lib.ptr(x), x.shape[0], x.shape[2], x.shape[3], size, _lib.ptr(out), int(self.device.index or 0),
                                                      _lib.stream_ptr(self.device)), "rdm_clip_preprocess")
        return out
