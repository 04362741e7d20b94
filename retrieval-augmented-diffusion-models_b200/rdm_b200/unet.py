"""Host wrapper of the U-Net executor in librdm_b200 (csrc/unet.cu).

``unet_param_shapes`` restates the parameter inventory of the reference's ``UNetModel.__init__``
(``rdm/modules/diffusionmodules/openaimodel.py:137-317``; key names per SURVEY.md Appendix C) so the
Python mirror can own real ``nn.Parameter`` tensors with the checkpoint's names and shapes; the C++ side
builds the same inventory independently and ``B200UNet.__init__`` asserts that both agree name by name.
"""
import ctypes
from collections import OrderedDict

import torch

from . import _lib

MODE_FP32 = 0          # every contraction fp32 on CUDA cores (strict)
MODE_TC_BF16X3 = 1     # tcgen05, bf16 hi/lo split operands (fp32-grade)
MODE_TC_BF16 = 2       # tcgen05, plain bf16 operands
MODE_TC_FP16X2 = 3     # tcgen05, fp16 activations x (fp16 hi + lo) weights, 2 MMAs per product
MODE_TC_FP16 = 4       # tcgen05, plain fp16 operands
MODES = {"fp32": 0, "bf16x3": 1, "bf16": 2, "fp16x2": 3, "fp16": 4}


class UNetCfg(ctypes.Structure):
    _fields_ = [("in_channels", ctypes.c_int32), ("model_channels", ctypes.c_int32), ("out_channels", ctypes.c_int32),
                ("num_res_blocks", ctypes.c_int32), ("n_attention_resolutions", ctypes.c_int32),
                ("attention_resolutions", ctypes.c_int32 * 8), ("n_channel_mult", ctypes.c_int32),
                ("channel_mult", ctypes.c_int32 * 8), ("num_head_channels", ctypes.c_int32), ("num_heads", ctypes.c_int32),
                ("transformer_depth", ctypes.c_int32), ("context_dim", ctypes.c_int32)]


def make_cfg(in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, channel_mult,
             num_head_channels=32, num_heads=-1, transformer_depth=1, context_dim=512):
    c = UNetCfg()
    c.in_channels, c.model_channels, c.out_channels, c.num_res_blocks = in_channels, model_channels, out_channels, num_res_blocks
    ar, cm = [int(a) for a in attention_resolutions], [int(m) for m in channel_mult]
    assert len(ar) <= 8 and len(cm) <= 8
    c.n_attention_resolutions, c.n_channel_mult = len(ar), len(cm)
    for i, a in enumerate(ar):
        c.attention_resolutions[i] = a
    for i, m in enumerate(cm):
        c.channel_mult[i] = m
    c.num_head_channels, c.num_heads, c.transformer_depth, c.context_dim = num_head_channels, num_heads, transformer_depth, int(context_dim)
    return c


def unet_param_shapes(in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, channel_mult,
                      context_dim=512, **_):
    """name -> shape, in the reference's registration order (openaimodel.py:137-317)."""
    P = OrderedDict()
    mc, ted = model_channels, 4 * model_channels

    def lin(p, i, o, bias=True):
        P[p + ".weight"] = (o, i)
        if bias:
            P[p + ".bias"] = (o,)

    def conv(p, i, o, k):
        P[p + ".weight"] = (o, i, k, k)
        P[p + ".bias"] = (o,)

    def norm(p, c):
        P[p + ".weight"] = (c,)
        P[p + ".bias"] = (c,)

    def res(p, i, o):
        norm(p + ".in_layers.0", i); conv(p + ".in_layers.2", i, o, 3)
        lin(p + ".emb_layers.1", ted, o)
        norm(p + ".out_layers.0", o); conv(p + ".out_layers.3", o, o, 3)
        if i != o:
            conv(p + ".skip_connection", i, o, 1)

    def st(p, c):
        norm(p + ".norm", c); conv(p + ".proj_in", c, c, 1)
        t = p + ".transformer_blocks.0"
        def attn(a, cd):
            lin(t + a + ".to_q", c, c, False); lin(t + a + ".to_k", cd, c, False); lin(t + a + ".to_v", cd, c, False)
            lin(t + a + ".to_out.0", c, c)
        attn(".attn1", c)                                                   # registration order of attention.py:80-86
        lin(t + ".ff.net.0.proj", c, 8 * c); lin(t + ".ff.net.2", 4 * c, c)
        attn(".attn2", context_dim)
        for i in (1, 2, 3):
            norm(t + f".norm{i}", c)
        conv(p + ".proj_out", c, c, 1)

    lin("time_embed.0", mc, ted); lin("time_embed.2", ted, ted)
    conv("input_blocks.0.0", in_channels, mc, 3)
    chans, ch, ds, nb = [mc], mc, 1, 1
    for level, mult in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            res(f"input_blocks.{nb}.0", ch, mult * mc); ch = mult * mc
            if ds in attention_resolutions:
                st(f"input_blocks.{nb}.1", ch)
            chans.append(ch); nb += 1
        if level != len(channel_mult) - 1:
            conv(f"input_blocks.{nb}.0.op", ch, ch, 3); chans.append(ch); nb += 1; ds *= 2
    res("middle_block.0", ch, ch); st("middle_block.1", ch); res("middle_block.2", ch, ch)
    nb = 0
    for level, mult in list(enumerate(channel_mult))[::-1]:
        for i in range(num_res_blocks + 1):
            li = 0
            res(f"output_blocks.{nb}.{li}", ch + chans.pop(), mc * mult); ch = mc * mult; li += 1
            if ds in attention_resolutions:
                st(f"output_blocks.{nb}.{li}", ch); li += 1
            if level and i == num_res_blocks:
                conv(f"output_blocks.{nb}.{li}.conv", ch, ch, 3); ds //= 2
            nb += 1
    norm("out.0", ch); conv("out.2", mc, out_channels, 3)
    return P


class B200UNet:
    """Owns one ``rdm_unet_t`` handle.  ``forward`` == ``UNetModel.forward`` (openaimodel.py:335-371)."""

    def __init__(self, device, **cfg):
        L = _lib.lib()
        self.device = _lib.resolve_device(device)
        self.cfg = dict(cfg)
        self.shapes = unet_param_shapes(**cfg)
        c = make_cfg(cfg["in_channels"], cfg["model_channels"], cfg["out_channels"], cfg["num_res_blocks"],
                     cfg["attention_resolutions"], cfg["channel_mult"], cfg.get("num_head_channels", 32),
                     cfg.get("num_heads", -1), cfg.get("transformer_depth", 1), cfg.get("context_dim", 512))
        self._h = ctypes.c_void_p()
        _lib.check(L.rdm_unet_create(ctypes.byref(self._h), ctypes.byref(c), int(self.device.index or 0)), "rdm_unet_create")
        names = [L.rdm_unet_param_name(self._h, i).decode() for i in range(L.rdm_unet_num_params(self._h))]
        assert sorted(names) == sorted(self.shapes), "parameter inventory of csrc/unet.cu and unet_param_shapes() differ"
        for k, shp in self.shapes.items():
            n = 1
            for s in shp:
                n *= s
            assert L.rdm_unet_param_numel(self._h, k.encode()) == n, k
        self._ctx_key = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().rdm_unet_destroy(h)
            except Exception:
                pass

    def load_state_dict(self, sd, strict=True):
        """sd: name -> tensor with the reference's names/shapes (any device/dtype; converted to host fp32)."""
        L = _lib.lib()
        missing = [k for k in self.shapes if k not in sd]
        if strict and missing:
            raise RuntimeError(f"missing U-Net parameters: {missing[:5]}{'...' if len(missing) > 5 else ''}")
        for k, shp in self.shapes.items():
            if k not in sd:
                continue
            t = sd[k].detach()
            if tuple(t.shape) != tuple(shp):
                raise RuntimeError(f"size mismatch for {k}: {tuple(t.shape)} vs {tuple(shp)}")
            t = t.to("cpu", torch.float32).contiguous()
            _lib.check(L.rdm_unet_load(self._h, k.encode(), ctypes.c_void_p(t.data_ptr()), t.numel()), f"rdm_unet_load({k})")
        self._ctx_key = None
        return missing

    def missing(self):
        return int(_lib.lib().rdm_unet_missing(self._h))

    def set_mode(self, mode):
        _lib.check(_lib.lib().rdm_unet_set_mode(self._h, int(mode)), "rdm_unet_set_mode")

    def set_context(self, context):
        """context: CUDA float32 [B2, k, context_dim]."""
        if isinstance(context, (list, tuple)):
            assert len(context) == 1
            context = context[0]
        context = context.to(self.device, torch.float32).contiguous()
        B2, k, d = context.shape
        with _lib.device_ctx(self.device):
            _lib.check(_lib.lib().rdm_unet_set_context(self._h, _lib.ptr(context), B2, k, _lib.stream_ptr(self.device)), "rdm_unet_set_context")
        self._ctx_B2 = B2

    def forward(self, x, t, out=None):
        """x: CUDA float32 NCHW [Bx, C, H, W] (Bx == B2 or B2/2), t: int64 [B2] -> eps [B2, C_out, H, W]."""
        x = x.to(self.device, torch.float32).contiguous()
        t = t.to(self.device, torch.int64).contiguous()
        B2, (Bx, _, H, W) = t.shape[0], x.shape
        if out is None:
            out = torch.empty((B2, self.cfg["out_channels"], H, W), dtype=torch.float32, device=self.device)
        with _lib.device_ctx(self.device):
            _lib.check(_lib.lib().rdm_unet_forward(self._h, _lib.ptr(x), Bx, _lib.ptr(t), B2, H, W, _lib.ptr(out), _lib.stream_ptr(self.device)),
                       "rdm_unet_forward")
        return out


    def set_graph(self, on):
        _lib.check(_lib.lib().rdm_unet_set_graph(self._h, 1 if on else 0), "rdm_unet_set_graph")

    def set_chains(self, chains):
        """Concurrent batch chains (rdm_unet_set_chains): 1..8 sub-batches of a forward run the layer sequence on their own streams."""
        _lib.check(_lib.lib().rdm_unet_set_chains(self._h, int(chains)), "rdm_unet_set_chains")

    def set_ablation(self, mask):
        """Measurement aid (rdm_unet_set_ablation): bit mask of kernel classes that the following forwards do NOT launch."""
        _lib.check(_lib.lib().rdm_unet_set_ablation(self._h, int(mask)), "rdm_unet_set_ablation")

    def profile_forward(self, x, t):
        """Graph-replayed forward with event-record nodes around every GEMM -> dict(tc_ms, tc_flop, simt_ms, simt_flop, total_ms, n_tc, n_simt).
        (Event nodes cost a few microseconds each: use the per-layer lines for SHARES; bench.py takes the in-graph GEMM time by ablation.)"""
        x = x.to(self.device, torch.float32).contiguous(); t = t.to(self.device, torch.int64).contiguous()
        B2, (Bx, _, H, W) = t.shape[0], x.shape
        out = torch.empty((B2, self.cfg["out_channels"], H, W), dtype=torch.float32, device=self.device)
        o8 = (ctypes.c_double * 8)()
        with _lib.device_ctx(self.device):
            _lib.check(_lib.lib().rdm_unet_profile_forward(self._h, _lib.ptr(x), Bx, _lib.ptr(t), B2, H, W, _lib.ptr(out), o8, _lib.stream_ptr(self.device)),
                       "rdm_unet_profile_forward")
        return dict(tc_ms=o8[0], tc_flop=o8[1], simt_ms=o8[2], simt_flop=o8[3], total_ms=o8[4], n_tc=int(o8[5]), n_simt=int(o8[6]))

    def ddim_sample(self, x, timesteps, coef, cfg_scale=1.0, first_step=0, num_steps=None, noise=None, want_pred_x0=False):
        """Fused sampling loop (rdm_ddim_sample): x [B,C,H,W] CUDA float32 is consumed and the result returned.
        timesteps int64 [S] / coef float32 [S,8] in sampling order (see sampler.make_ddim_tables)."""
        x = x.to(self.device, torch.float32).contiguous().clone()
        B, _, H, W = x.shape
        S = timesteps.shape[0]
        num_steps = S - first_step if num_steps is None else num_steps
        p0 = torch.empty_like(x) if want_pred_x0 else None
        with _lib.device_ctx(self.device):
            _lib.check(_lib.lib().rdm_ddim_sample(self._h, _lib.ptr(x), B, H, W, _lib.ptr(timesteps), _lib.ptr(coef), int(first_step), int(num_steps),
                                                  float(cfg_scale), _lib.ptr(noise), _lib.ptr(p0), _lib.stream_ptr(self.device)), "rdm_ddim_sample")
        return (x, p0) if want_pred_x0 else x


def ddim_step(x, eps, coef_row, cfg_scale=None, noise=None, want_pred_x0=True):
    """One fused CFG + DDIM update (ddim.py:236-238,253-267).  x [B,...], eps [2B,...] if cfg_scale is not None else [B,...];
    coef_row: CUDA float32 [>=5] = {sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma}."""
    x, eps = x.contiguous(), eps.contiguous()
    xp = torch.empty_like(x)
    p0 = torch.empty_like(x) if want_pred_x0 else None
    with _lib.device_ctx(x.device):
        _lib.check(_lib.lib().rdm_ddim_step(_lib.ptr(x), _lib.ptr(eps), x.numel(), 1 if cfg_scale is not None else 0,
                                            float(cfg_scale if cfg_scale is not None else 1.0), _lib.ptr(coef_row),
                                            _lib.ptr(noise.contiguous() if noise is not None else None), _lib.ptr(xp), _lib.ptr(p0),
                                            int(x.device.index or 0), _lib.stream_ptr(x.device)), "rdm_ddim_step")
    return xp, p0
