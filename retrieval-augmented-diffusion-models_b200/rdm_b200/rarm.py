"""Host wrapper of the RARM decoder executor in librdm_b200 (csrc/rarm.cu): the reference's `RetrievalPatchTransformer`
(`rdm/modules/attention.py:199-272`, `continuous: false`) evaluated with key/value caches, and the sampling loop of
`LatentImageRETRO.sample` (`rdm/models/autoregression/transformer.py:224-270`) as one captured CUDA graph per position."""
import ctypes
from collections import OrderedDict

import torch

from . import _lib

RARM_IMAGENET = dict(in_channels=16386, n_heads=12, d_head=64, depth=18, context_dim=512, sequence_length=256, out_channels=16384)
"""`models/rarm/imagenet/{dogs,mammals,animals}/config.yaml:14-27` (230.9 M parameters)."""
MODE_FP32, MODE_FP16 = 0, 4


class RarmCfg(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in ("in_channels", "n_heads", "d_head", "depth", "context_dim", "sequence_length", "out_channels")]


def rarm_param_shapes(in_channels, n_heads, d_head, depth, context_dim, sequence_length, out_channels, **_):
    """name -> shape in the reference's registration order (attention.py:224-245; block order attn1, ff, attn2, norm1..3 :79-87)."""
    C = n_heads * d_head
    P = OrderedDict()
    P["positional_encoding"], P["proj_in.weight"] = (C, sequence_length), (in_channels, C)
    for i in range(depth):
        p = f"transformer_blocks.{i}."
        for a, cd in (("attn1", C), ("ff", None), ("attn2", context_dim)):
            if a == "ff":
                P[p + "ff.net.0.proj.weight"], P[p + "ff.net.0.proj.bias"] = (8 * C, C), (8 * C,)
                P[p + "ff.net.2.weight"], P[p + "ff.net.2.bias"] = (C, 4 * C), (C,)
                continue
            P[p + a + ".to_q.weight"], P[p + a + ".to_k.weight"], P[p + a + ".to_v.weight"] = (C, C), (C, cd), (C, cd)
            P[p + a + ".to_out.0.weight"], P[p + a + ".to_out.0.bias"] = (C, C), (C,)
        for nm in ("norm1", "norm2", "norm3"):
            P[p + nm + ".weight"], P[p + nm + ".bias"] = (C,), (C,)
    P["proj_out.weight"], P["proj_out.bias"] = (out_channels, C, 1), (out_channels,)
    return P


class B200Rarm:
    """Owns one `rdm_rarm_t` handle."""

    def __init__(self, device, **cfg):
        L = self._library()
        self.device, index = self._resolve_device(device)
        self.cfg = {k: int(cfg[k]) for k, _ in RarmCfg._fields_}
        self.shapes = rarm_param_shapes(**self.cfg)
        c = RarmCfg(**self.cfg)
        self._h = ctypes.c_void_p()
        self._check(L.rdm_rarm_create(ctypes.byref(self._h), ctypes.byref(c), index), "rdm_rarm_create")
        names = [L.rdm_rarm_param_name(self._h, i).decode() for i in range(L.rdm_rarm_num_params(self._h))]
        assert names == list(self.shapes), "parameter inventory of csrc/rarm.cu and rarm_param_shapes() differ"
        self._B2 = 0

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._library().rdm_rarm_destroy(h)
            except Exception:
                pass

    # The three hooks below are the only places that touch the CUDA device; they go through the same `_lib` indirections as the other
    # wrappers (tests/emu drives the SAME C ABI compiled against a host emulation of CUDA by patching those).  The product has no other
    # implementation: _library() raises without the .so, resolve_device() raises for anything but a CUDA device.
    def _library(self):
        return _lib.lib()

    def _resolve_device(self, device):
        device = _lib.resolve_device(device)
        return device, int(device.index or 0)

    def _run(self, fn, *args):
        with _lib.device_ctx(self.device):
            self._check(getattr(self._library(), fn)(self._h, *args, _lib.stream_ptr(self.device)), fn)

    def _check(self, rc, what=""):
        if rc != 0:
            msg = self._library().rdm_last_error()
            raise RuntimeError(f"librdm_b200 {what} failed ({rc}): {msg.decode() if msg else ''}")

    def load_state_dict(self, sd, strict=True):
        L = self._library()
        missing = [k for k in self.shapes if k not in sd]
        if strict and missing:
            raise RuntimeError(f"missing RARM parameters: {missing[:5]}{'...' if len(missing) > 5 else ''}")
        for k, shp in self.shapes.items():
            if k not in sd:
                continue
            t = sd[k].detach()
            if tuple(t.shape) != tuple(shp):
                raise RuntimeError(f"size mismatch for {k}: {tuple(t.shape)} vs {tuple(shp)}")
            t = t.to("cpu", torch.float32).contiguous()
            self._check(L.rdm_rarm_load(self._h, k.encode(), ctypes.c_void_p(t.data_ptr()), t.numel()), f"rdm_rarm_load({k})")
        return missing

    def missing(self):
        return int(self._library().rdm_rarm_missing(self._h))

    def set_mode(self, mode):
        self._check(self._library().rdm_rarm_set_mode(self._h, int(mode)), "rdm_rarm_set_mode")

    def set_graph(self, on):
        self._check(self._library().rdm_rarm_set_graph(self._h, int(bool(on))), "rdm_rarm_set_graph")

    def set_context(self, context):
        """context: float32 [B2, k, context_dim]; restarts the sequences."""
        context = context.to(self.device, torch.float32).contiguous()
        assert context.ndim == 3 and context.shape[2] == self.cfg["context_dim"]
        self._run("rdm_rarm_set_context", _lib.ptr(context), context.shape[0], context.shape[1])
        self._B2 = context.shape[0]

    def forward_token(self, tokens, pos):
        """tokens int64 [B] (B == B2 or B2 / 2) at position `pos` -> logits float32 [B2, out_channels]."""
        tokens = tokens.to(self.device, torch.int64).contiguous()
        out = torch.empty((self._B2, self.cfg["out_channels"]), dtype=torch.float32, device=self.device)
        self._run("rdm_rarm_forward_token", _lib.ptr(tokens), tokens.shape[0], int(pos), _lib.ptr(out))
        return out

    def forward(self, tokens, context):
        """`RetrievalPatchTransformer.forward(x, context)` for discrete x: int64 [B, T] -> logits [B, T, out_channels]
        (positions fed in order through the caches; used for parity tests and teacher-forced scoring)."""
        self.set_context(context)
        tokens = tokens.to(self.device, torch.int64)
        return torch.stack([self.forward_token(tokens[:, t], t) for t in range(tokens.shape[1])], dim=1)

    def sample_step(self, logits, guidance_scale=1.0, temperature=1.0, top_k=None, uniforms=None, want_probs=False):
        """transformer.py:249-266 on last-position logits [B or 2B, V] -> (tokens int64 [B], probs [B, V] or None)."""
        guided = guidance_scale > 1.0
        logits = logits.to(self.device, torch.float32).contiguous()
        B = logits.shape[0] // 2 if guided else logits.shape[0]
        tok = torch.empty((B,), dtype=torch.int64, device=self.device)
        probs = torch.empty((B, logits.shape[1]), dtype=torch.float32, device=self.device) if want_probs else None
        u = None if uniforms is None else uniforms.to(self.device, torch.float32).contiguous()
        self._run("rdm_rarm_sample_step", _lib.ptr(logits), B, int(logits.shape[1]), int(guided), float(guidance_scale), float(temperature), int(top_k or 0),
                  _lib.ptr(u), _lib.ptr(tok), _lib.ptr(probs))
        return tok, probs

    def sample(self, prefix, steps, temperature=1.0, top_k=None, guidance_scale=1.0, uniforms=None):
        """`LatentImageRETRO.sample` after `x = cat((c, x), 1)`: prefix int64 [B, n_prefix >= 1] -> int64 [B, n_prefix + steps].
        uniforms float32 [steps, B] in [0, 1) (None: greedy).  The context must hold B rows, or 2B ([r | zeros]) when guided."""
        prefix = prefix.to(self.device, torch.int64)
        B, n_prefix = prefix.shape
        tokens = torch.zeros((B, n_prefix + steps), dtype=torch.int64, device=self.device)
        tokens[:, :n_prefix] = prefix
        u = None if uniforms is None else uniforms.to(self.device, torch.float32).contiguous()
        assert u is None or tuple(u.shape) == (steps, B)
        self._run("rdm_rarm_sample", _lib.ptr(tokens), B, n_prefix, int(steps), float(temperature), int(top_k or 0), float(guidance_scale), _lib.ptr(u))
        return tokens
