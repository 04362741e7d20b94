"""Host wrapper of the first-stage VQ decoder in librdm_b200 (csrc/unet.cu, `rdm_vqdec_*`): `VQModelInterface.decode`
= VectorQuantizer lookup -> post_quant_conv -> Decoder, i.e. `decode_first_stage` of rdm/models/diffusion/ddpm.py:840,981
(configuration: `first_stage_config.params` of models/rdm/imagenet/config.yaml:60-80).
The handle shares the U-Net executor, so parameters are loaded through the `rdm_unet_*` state-dict calls."""
import ctypes

import torch

from . import _lib


class VQDecCfg(ctypes.Structure):
    _fields_ = [("embed_dim", ctypes.c_int32), ("n_embed", ctypes.c_int32), ("z_channels", ctypes.c_int32), ("resolution", ctypes.c_int32),
                ("out_ch", ctypes.c_int32), ("ch", ctypes.c_int32), ("num_res_blocks", ctypes.c_int32),
                ("n_ch_mult", ctypes.c_int32), ("ch_mult", ctypes.c_int32 * 8),
                ("n_attn_resolutions", ctypes.c_int32), ("attn_resolutions", ctypes.c_int32 * 8)]


class B200VQDecoder:
    """Owns one decoder handle.  ``decode(z)`` == ``VQModelInterface.decode(z, force_not_quantize)`` on a CUDA tensor."""

    def __init__(self, device, embed_dim, n_embed, ddconfig):
        L = _lib.lib()
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        c = VQDecCfg()
        c.embed_dim, c.n_embed = int(embed_dim), int(n_embed)
        c.z_channels, c.resolution, c.out_ch, c.ch = int(ddconfig["z_channels"]), int(ddconfig["resolution"]), int(ddconfig["out_ch"]), int(ddconfig["ch"])
        c.num_res_blocks = int(ddconfig["num_res_blocks"])
        cm, ar = [int(m) for m in ddconfig.get("ch_mult", (1, 2, 4, 8))], [int(a) for a in ddconfig.get("attn_resolutions", [])]
        c.n_ch_mult, c.n_attn_resolutions = len(cm), len(ar)
        for i, m in enumerate(cm):
            c.ch_mult[i] = m
        for i, a in enumerate(ar):
            c.attn_resolutions[i] = a
        self.cfg, self.up, self.out_ch, self.embed_dim = c, 2 ** (len(cm) - 1), c.out_ch, c.embed_dim
        self._h = ctypes.c_void_p()
        _lib.check(L.rdm_vqdec_create(ctypes.byref(self._h), ctypes.byref(c), self.device.index), "rdm_vqdec_create")
        self.names = [L.rdm_unet_param_name(self._h, i).decode() for i in range(L.rdm_unet_num_params(self._h))]

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().rdm_unet_destroy(h)
            except Exception:
                pass

    def load_state_dict(self, sd, strict=True):
        """sd: `decoder.*`, `quantize.embedding.weight`, `post_quant_conv.*` (latent-diffusion names, any device / dtype)."""
        L = _lib.lib()
        missing = [k for k in self.names if k not in sd]
        if strict and missing:
            raise RuntimeError(f"missing first-stage parameters: {missing[:5]}{'...' if len(missing) > 5 else ''}")
        for k in self.names:
            if k not in sd:
                continue
            t = sd[k].detach().to("cpu", torch.float32).contiguous()
            if t.numel() != L.rdm_unet_param_numel(self._h, k.encode()):
                raise RuntimeError(f"size mismatch for {k}: {tuple(t.shape)}")
            _lib.check(L.rdm_unet_load(self._h, k.encode(), ctypes.c_void_p(t.data_ptr()), t.numel()), f"rdm_unet_load({k})")
        return missing

    def set_mode(self, mode):
        _lib.check(_lib.lib().rdm_unet_set_mode(self._h, int(mode)), "rdm_unet_set_mode")

    def decode(self, z, force_not_quantize=False):
        z = z.to(self.device, torch.float32).contiguous()
        B, E, h, w = z.shape
        assert E == self.embed_dim, f"latent has {E} channels, the codebook {self.embed_dim}"
        out = torch.empty((B, self.out_ch, h * self.up, w * self.up), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().rdm_vqdec_decode(self._h, _lib.ptr(z), B, h, w, 0 if force_not_quantize else 1, _lib.ptr(out),
                                                   _lib.stream_ptr(self.device)), "rdm_vqdec_decode")
        return out
