"""`rdm.modules.attention.RetrievalPatchTransformer` -- same constructor and state-dict layout as the reference
(`rdm/modules/attention.py:199-272`) for the configuration the shipped RARM models use (`models/rarm/imagenet/*/config.yaml:14-27`:
`continuous: false`, `positional_encodings: true`, `cross_attend: true`, `causal: true`), executed by the key/value-cached decoder of
librdm_b200 (csrc/rarm.cu).  The U-Net's SpatialTransformer / CrossAttention blocks (`attention.py:20-196`) live inside the U-Net
engine (csrc/unet.cu) and are not separate Python modules here.  No PyTorch compute path exists: forward needs a CUDA device."""
import math
import os

import torch
import torch.nn as nn

from rdm_b200 import _lib as _binding
from rdm_b200.rarm import MODE_FP16, MODE_FP32, B200Rarm, rarm_param_shapes


def _set_param(root, dotted, tensor):
    mod, parts = root, dotted.split(".")
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, nn.Module())
        mod = mod._modules[p]
    mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


class RetrievalPatchTransformer(nn.Module):
    def __init__(self, in_channels, n_heads, d_head, depth=1, context_dim=None, dropout=0., positional_encodings=False, sequence_length=None,
                 residual=False, checkpoint=False, out_channels=None, cross_attend=False, causal=False, continuous=True):
        super().__init__()
        if cross_attend:
            assert context_dim is not None
        unsupported = dict(continuous=continuous, no_positional_encodings=not positional_encodings, not_causal=not causal, residual=residual,
                           no_context=context_dim is None, d_head=d_head != 64)
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError(f"RetrievalPatchTransformer options outside the shipped RARM configurations are not implemented: {bad}")
        assert sequence_length is not None, 'Need sequence length for positional embedding'
        self.in_channels, self.residual, self.checkpoint, self.continuous = in_channels, residual, checkpoint, continuous
        out_channels = in_channels if out_channels is None else out_channels
        self._cfg = dict(in_channels=int(in_channels), n_heads=int(n_heads), d_head=int(d_head), depth=int(depth), context_dim=int(context_dim),
                         sequence_length=int(sequence_length), out_channels=int(out_channels))
        inner = n_heads * d_head
        for name, shp in rarm_param_shapes(**self._cfg).items():           # torch default initialisations of the reference's layers
            if name == "positional_encoding":
                t = torch.randn(shp) / inner ** 0.5                          # attention.py:239
            elif name == "proj_in.weight":
                t = torch.randn(shp)                                         # nn.Embedding
            elif len(shp) >= 2:
                bound = 1.0 / math.sqrt(int(torch.tensor(shp[1:]).prod()))
                t = torch.empty(shp).uniform_(-bound, bound)
            elif ".norm" in name:
                t = torch.ones(shp) if name.endswith("weight") else torch.zeros(shp)
            else:
                t = torch.zeros(shp)
            _set_param(self, name, t)
        self.engine_mode = os.environ.get("RDM_B200_RARM_MODE", "fp16")      # "fp32": fp32 weights (strict parity)
        self._engine, self._loaded_key = None, None

    def engine(self, device):
        device = _binding.resolve_device(device)          # raises for anything but a CUDA device: there is no CPU path
        if self._engine is None or self._engine.device != device:
            self._engine, self._loaded_key = B200Rarm(device, **self._cfg), None
        key = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (self.engine_mode,)
        if key != self._loaded_key:
            self._engine.load_state_dict(self.state_dict())
            self._engine.set_mode(MODE_FP32 if self.engine_mode == "fp32" else MODE_FP16)
            self._loaded_key = key
        return self._engine

    def forward(self, x, context=None):
        """x: int64 token ids [b, t]; context [b, k, context_dim] -> logits [b, t, out_channels] (attention.py:247-272)."""
        assert context is not None and x.dtype in (torch.int64, torch.int32)
        return self.engine(x.device).forward(x, context)
