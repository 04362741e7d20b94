"""Stand-ins for the packages the REFERENCE imports but does not vendor, so that its own hot-path modules can be imported and
run on CPU in the build container (used ONLY by tests/golden/make_golden_ref.py; never on the GPU box, never by the product).

The reference's `rdm/modules/attention.py`, `rdm/modules/diffusionmodules/openaimodel.py` and `rdm/models/diffusion/ddim.py`
import `ldm` (latent-diffusion@main, un-pinned git dependency, `environment.yaml:38`), `kornia` and `main`.  What they need from
there is restated below from the published latent-diffusion sources with ldm's signatures: these few
functions are third-party arithmetic, everything else executed by the generator is the reference's own code, unmodified, from
/root/reference.  Kept independent of `oracle/` on purpose, so the fixtures check the oracle rather than echo it.
"""
import importlib
import inspect
import math
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ---- ldm.util ----------------------------------------------------------------------------------------------
def exists(x):
    return x is not None


def default(val, d):
    if exists(val):
        return val
    return d() if inspect.isfunction(d) else d


def instantiate_from_config(config):
    module, cls = config["target"].rsplit(".", 1)
    return getattr(importlib.import_module(module), cls)(**config.get("params", dict()))


# ---- ldm.modules.diffusionmodules.util ---------------------------------------------------------------------
def checkpoint(func, inputs, params, flag):
    return func(*inputs)            # gradient checkpointing is a training-memory device; the forward value is func(*inputs)


def conv_nd(dims, *args, **kwargs):
    return {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}[dims](*args, **kwargs)


def linear(*args, **kwargs):
    return nn.Linear(*args, **kwargs)


def avg_pool_nd(dims, *args, **kwargs):
    return {1: nn.AvgPool1d, 2: nn.AvgPool2d, 3: nn.AvgPool3d}[dims](*args, **kwargs)


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


class GroupNorm32(nn.GroupNorm):
    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


def normalization(channels):
    return GroupNorm32(32, channels)


def timestep_embedding(timesteps, dim, max_period=10000, repeat_only=False):
    assert not repeat_only
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(device=timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    embedding = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        embedding = torch.cat([embedding, torch.zeros_like(embedding[:, :1])], dim=-1)
    return embedding


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    assert ddim_discr_method == "uniform"
    c = num_ddpm_timesteps // num_ddim_timesteps
    ddim_timesteps = np.asarray(list(range(0, num_ddpm_timesteps, c)))
    return ddim_timesteps + 1        # "add one to get the final alpha values right (the ones from first scale to data during sampling)"


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


def noise_like(shape, device, repeat=False):
    assert not repeat
    return torch.randn(shape, device=device)


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    assert schedule == "linear"
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    return betas.numpy()


# ---- ldm.modules.attention ---------------------------------------------------------------------------------
class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0.):
        super().__init__()
        inner_dim = int(dim * mult)
        dim_out = default(dim_out, dim)
        project_in = nn.Sequential(nn.Linear(dim, inner_dim), nn.GELU()) if not glu else GEGLU(dim, inner_dim)
        self.net = nn.Sequential(project_in, nn.Dropout(dropout), nn.Linear(inner_dim, dim_out))

    def forward(self, x):
        return self.net(x)


# ---- ldm.modules.diffusionmodules.openaimodel --------------------------------------------------------------
class TimestepBlock(nn.Module):
    def forward(self, x, emb):
        raise NotImplementedError


class Upsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        self.channels, self.out_channels, self.use_conv, self.dims = channels, out_channels or channels, use_conv, dims
        if use_conv:
            self.conv = conv_nd(dims, self.channels, self.out_channels, 3, padding=padding)

    def forward(self, x):
        assert x.shape[1] == self.channels
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if self.use_conv:
            x = self.conv(x)
        return x


class Downsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        self.channels, self.out_channels, self.use_conv, self.dims = channels, out_channels or channels, use_conv, dims
        stride = 2 if dims != 3 else (1, 2, 2)
        if use_conv:
            self.op = conv_nd(dims, self.channels, self.out_channels, 3, stride=stride, padding=padding)
        else:
            assert self.channels == self.out_channels
            self.op = avg_pool_nd(dims, kernel_size=stride, stride=stride)

    def forward(self, x):
        assert x.shape[1] == self.channels
        return self.op(x)


class ResBlock(TimestepBlock):
    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False, use_scale_shift_norm=False, dims=2,
                 use_checkpoint=False, up=False, down=False):
        super().__init__()
        assert not (use_scale_shift_norm or up or down), "the shipped configs use plain ResBlocks"
        self.channels, self.emb_channels, self.dropout = channels, emb_channels, dropout
        self.out_channels = out_channels or channels
        self.in_layers = nn.Sequential(normalization(channels), nn.SiLU(), conv_nd(dims, channels, self.out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(normalization(self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        zero_module(conv_nd(dims, self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        elif use_conv:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 3, padding=1)
        else:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 1)

    def forward(self, x, emb):
        h = self.in_layers(x)
        emb_out = self.emb_layers(emb).type(h.dtype)
        while len(emb_out.shape) < len(h.shape):
            emb_out = emb_out[..., None]
        h = h + emb_out
        h = self.out_layers(h)
        return self.skip_connection(x) + h


class AttentionBlock(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("the shipped configs set use_spatial_transformer=True")


# ---- ldm.modules.ema / ldm.models.diffusion.ddpm (what MinimalRETRODiffusion's sampling methods use of its base class) ----------
class LitEma(nn.Module):
    def __init__(self, model, decay=0.9999, use_num_upates=True):
        super().__init__()
        self.m_name2s_name = {}
        self.register_buffer("decay", torch.tensor(decay, dtype=torch.float32))
        self.register_buffer("num_updates", torch.tensor(0, dtype=torch.int) if use_num_upates else torch.tensor(-1, dtype=torch.int))
        for name, p in model.named_parameters():
            if p.requires_grad:
                s_name = name.replace(".", "")                    # "remove as '.'-character is not allowed in buffers"
                self.m_name2s_name.update({name: s_name})
                self.register_buffer(s_name, p.clone().detach().data)
        self.collected_params = []

    def copy_to(self, model):
        m_param, shadow_params = dict(model.named_parameters()), dict(self.named_buffers())
        for key in m_param:
            if m_param[key].requires_grad:
                m_param[key].data.copy_(shadow_params[self.m_name2s_name[key]].data)

    def store(self, parameters):
        self.collected_params = [param.clone() for param in parameters]

    def restore(self, parameters):
        for c_param, param in zip(self.collected_params, parameters):
            param.data.copy_(c_param.data)


class LdmDiffusionWrapper(nn.Module):
    def __init__(self, diff_model_config, conditioning_key):
        super().__init__()
        self.diffusion_model = instantiate_from_config(diff_model_config)
        self.conditioning_key = conditioning_key


class IdentityFirstStage(nn.Module):
    def decode(self, x, *args, **kwargs):
        return x


class LatentDiffusion(nn.Module):
    """The slice of ldm's DDPM / LatentDiffusion that sampling touches: the eps-model wrapper, the float32 schedule buffers
    (register_schedule), `device`, `ema_scope` (store -> copy_to -> restore) and `decode_first_stage` over an identity first stage."""

    def __init__(self, unet_config, first_stage_config=None, cond_stage_config=None, timesteps=1000, beta_schedule="linear", linear_start=1e-4,
                 linear_end=2e-2, image_size=256, channels=3, conditioning_key=None, scale_factor=1.0, log_every_t=100, parameterization="eps",
                 first_stage_key="image", cond_stage_key="image", **unused):
        super().__init__()
        if cond_stage_config == "__is_unconditional__":
            conditioning_key = None
        self.parameterization, self.log_every_t, self.image_size, self.channels = parameterization, log_every_t, image_size, channels
        self.model = LdmDiffusionWrapper(unet_config, conditioning_key)
        betas = make_beta_schedule(beta_schedule, timesteps, linear_start=linear_start, linear_end=linear_end)
        alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
        self.num_timesteps = int(timesteps)
        to_torch = lambda a: torch.tensor(a, dtype=torch.float32)
        self.register_buffer("betas", to_torch(betas))
        self.register_buffer("alphas_cumprod", to_torch(alphas_cumprod))
        self.register_buffer("alphas_cumprod_prev", to_torch(np.append(1.0, alphas_cumprod[:-1])))
        self.first_stage_model = IdentityFirstStage()
        self.scale_factor = scale_factor

    @property
    def device(self):
        return self.betas.device

    def ema_scope(self, context=None):
        import contextlib

        @contextlib.contextmanager
        def scope():
            if self.use_ema:
                self.model_ema.store(self.model.parameters())
                self.model_ema.copy_to(self.model)
            try:
                yield None
            finally:
                if self.use_ema:
                    self.model_ema.restore(self.model.parameters())
        return scope()

    def decode_first_stage(self, z, predict_cids=False, force_not_quantize=False):
        return self.first_stage_model.decode(1.0 / self.scale_factor * z)


# ---- taming.models.cond_transformer (what LatentImageRETRO's sampling methods use of Net2NetTransformer) ---------------------------
class SOSProvider(nn.Module):
    def __init__(self, sos_token, quantize_interface=True):
        super().__init__()
        self.sos_token, self.quantize_interface = sos_token, quantize_interface

    def encode(self, x):
        c = torch.ones(x.shape[0], 1) * self.sos_token
        c = c.long().to(x.device)
        if self.quantize_interface:
            return c, None, [None, None, c]
        return c


class Net2NetTransformer(nn.Module):
    """The slice of taming's Net2NetTransformer that sampling touches (published taming-transformers source): first stage, SOS
    conditioning for `__is_unconditional__`, identity permuter, the transformer built from its config, `top_k_logits`, `encode_to_c`,
    `decode_to_img`."""

    def __init__(self, transformer_config, first_stage_config, cond_stage_config, permuter_config=None, ckpt_path=None, ignore_keys=[],
                 first_stage_key="image", cond_stage_key="depth", downsample_cond_size=-1, pkeep=1.0, sos_token=0, unconditional=False):
        super().__init__()
        self.be_unconditional, self.sos_token = unconditional, sos_token
        self.first_stage_key, self.cond_stage_key = first_stage_key, cond_stage_key
        self.first_stage_model = instantiate_from_config(first_stage_config).eval()
        if cond_stage_config == "__is_unconditional__" or self.be_unconditional:
            print(f"Using no cond stage. Assuming the training is intended to be unconditional. Prepending {self.sos_token} as a sos token.")
            self.be_unconditional, self.cond_stage_key = True, self.first_stage_key
            self.cond_stage_model = SOSProvider(self.sos_token)
        else:
            raise NotImplementedError
        assert permuter_config is None
        self.permuter = lambda x, reverse=False: x
        self.transformer = instantiate_from_config(transformer_config)
        self.downsample_cond_size, self.pkeep = downsample_cond_size, pkeep

    @property
    def device(self):                                            # pl.LightningModule.device
        return next(self.parameters()).device

    def top_k_logits(self, logits, k):
        v, ix = torch.topk(logits, k)
        out = logits.clone()
        out[out < v[..., [-1]]] = -float("Inf")
        return out

    @torch.no_grad()
    def encode_to_c(self, c):
        if self.downsample_cond_size > -1:
            c = F.interpolate(c, size=(self.downsample_cond_size, self.downsample_cond_size))
        quant_c, _, [_, _, indices] = self.cond_stage_model.encode(c)
        if len(indices.shape) > 2:
            indices = indices.view(c.shape[0], -1)
        return quant_c, indices

    @torch.no_grad()
    def decode_to_img(self, index, zshape):
        index = self.permuter(index, reverse=True)
        bhwc = (zshape[0], zshape[2], zshape[3], zshape[1])
        quant_z = self.first_stage_model.quantize.get_codebook_entry(index.reshape(-1), shape=bhwc)
        x = self.first_stage_model.decode(quant_z)
        return x


class AttrDict(dict):
    """Stand-in for an OmegaConf node: `cfg.params.context_dim` and `cfg["target"]` both work."""
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v


def install(reference_root="/root/reference"):
    """Registers the stand-ins in sys.modules and puts the reference checkout first on sys.path."""
    me = sys.modules[__name__]

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
        return m

    pick = lambda *names: {n: getattr(me, n) for n in names}
    mod("ldm")
    mod("ldm.util", **pick("exists", "default", "instantiate_from_config"), log_txt_as_img=None, isimage=None, ismap=None)
    mod("ldm.modules")
    mod("ldm.modules.attention", **pick("FeedForward", "GEGLU", "zero_module"))
    mod("ldm.modules.diffusionmodules")
    mod("ldm.modules.diffusionmodules.util", **pick("checkpoint", "conv_nd", "linear", "avg_pool_nd", "zero_module", "normalization",
                                                    "timestep_embedding", "make_ddim_timesteps", "make_ddim_sampling_parameters",
                                                    "noise_like", "make_beta_schedule"))
    mod("ldm.modules.diffusionmodules.openaimodel", **pick("TimestepBlock", "ResBlock", "Downsample", "Upsample", "AttentionBlock"))
    mod("ldm.modules.ema", **pick("LitEma"))
    mod("ldm.models")
    mod("ldm.models.autoencoder", AutoencoderKL=type("AutoencoderKL", (nn.Module,), {}), VQModelInterface=type("VQModelInterface", (nn.Module,), {}),
        VQModel=type("VQModel", (nn.Module,), {}), **pick("IdentityFirstStage"))
    cls = lambda n: type(n, (nn.Module,), {})                    # imported by rdm/modules/encoders/nn_encoders.py:6-9, never instantiated here
    mod("ldm.modules.x_transformer", AbsolutePositionalEmbedding=cls("AbsolutePositionalEmbedding"), Encoder=cls("Encoder"), always=None, **pick("exists"))
    mod("ldm.models.diffusion")
    mod("ldm.models.diffusion.ddpm", **pick("LatentDiffusion"))
    sys.modules["ldm.util"].get_obj_from_str = lambda s, reload=False: getattr(importlib.import_module(s.rsplit(".", 1)[0]), s.rsplit(".", 1)[1])
    sys.modules["ldm.util"].isimage = lambda x: isinstance(x, torch.Tensor) and x.ndim == 4 and x.shape[1] in (1, 3)
    mod("pytorch_lightning", LightningModule=nn.Module, seed_everything=lambda s: None)
    mod("pytorch_lightning.utilities")
    mod("pytorch_lightning.utilities.distributed", rank_zero_only=lambda f: f)
    mod("taming")
    mod("taming.models")
    mod("taming.models.cond_transformer", **pick("Net2NetTransformer", "SOSProvider"))
    sys.modules["ldm.util"].log_txt_as_img = None
    mod("main", **pick("instantiate_from_config"))
    mod("kornia")
    mod("omegaconf")
    mod("omegaconf.listconfig", ListConfig=type("ListConfig", (list,), {}))      # openaimodel.py:102 only type-checks against it
    for k in [k for k in sys.modules if k == "rdm" or k.startswith("rdm.")]:       # never mix with the repo's own mirror package
        del sys.modules[k]
    if reference_root in sys.path:
        sys.path.remove(reference_root)
    sys.path.insert(0, reference_root)
