"""`rdm.models.diffusion.ddpm.MinimalRETRODiffusion` -- the sampling surface of the reference model wrapper
(`rdm/models/diffusion/ddpm.py:134-1042`) on top of librdm_b200.

Kept: constructor keys of the shipped configs (`models/rdm/*/config.yaml`), checkpoint key layout (`model.diffusion_model.*`,
`model_ema.*`, `first_stage_model.*`), `load_state_dict(strict=False)`, `.eval()/.to()/.device`,
`get_qids`, `sample_from_rdata`, `sample_with_query`, `get_unconditional_conditioning`, `apply_model`, `sample_log`,
`decode_first_stage`, `ema_scope`, `q_sample`, `train_searcher`, `.retriever`.  Training (`shared_step`, `p_losses`,
`configure_optimizers`, `log_images`) is out of scope and raises.  Retrieval stays ON THE DEVICE: query rows, exact kNN
and the neighbour gather never leave HBM (the reference does them on the host, ddpm.py:897-921).
"""
import os
import pickle
from contextlib import contextmanager

import numpy as np
import torch
import torch.nn as nn

from ldm.util import instantiate_from_config
from rdm.models.diffusion.ddim import DDIMSampler
from rdm.util import SampleLogs, ischannellastimage, isimage
from rdm_b200 import sampler as _tables
from rdm_b200.knn import search_raw


def disabled_train(self, mode=True):
    return self


class DiffusionWrapper(nn.Module):
    """ldm DiffusionWrapper / ddpm.py:45-80: holds the U-Net at `.diffusion_model`."""

    def __init__(self, diff_model_config, conditioning_key, concat_dim=1):
        super().__init__()
        self.diffusion_model = instantiate_from_config(diff_model_config)
        self.retro_mode = conditioning_key == "retro_only"
        self.conditioning_key = None if self.retro_mode else conditioning_key
        assert self.conditioning_key in [None, 'crossattn'], "only cross-attention conditioning is implemented"
        self.wrapper_conditioning_key = self.conditioning_key
        self.concat_dim = concat_dim

    def forward(self, x, t, c_concat=None, c_crossattn=None):
        return self.diffusion_model(x, t, context=torch.cat(c_crossattn, self.concat_dim))


class RETRODiffusionWrapper(nn.Module):
    """ddpm.py:107-131: re-wraps the U-Net, conditionings chained in a list into the SpatialTransformers."""

    def __init__(self, diffusion_wrapper, concat=False):
        super().__init__()
        assert not concat, "retro_concat is not implemented"
        self.concat = concat
        self.diffusion_model = diffusion_wrapper.diffusion_model
        self.conditioning_key = diffusion_wrapper.conditioning_key
        self.wrapper_conditioning_key = diffusion_wrapper.conditioning_key

    def forward(self, x, t, c_crossattn=None):
        return self.diffusion_model(x, t, context=c_crossattn)


class LitEma(nn.Module):
    """ldm LitEma: shadow buffers named by parameter name without dots."""

    def __init__(self, model, decay=0.9999, use_num_upates=True):
        super().__init__()
        self.m_name2s_name = {}
        self.register_buffer('decay', torch.tensor(decay, dtype=torch.float32))
        self.register_buffer('num_updates', torch.tensor(0, dtype=torch.int) if use_num_upates else torch.tensor(-1, dtype=torch.int))
        for name, p in model.named_parameters():
            s_name = name.replace('.', '')
            self.m_name2s_name[name] = s_name
            self.register_buffer(s_name, p.clone().detach().data)

    def state_dict_for(self, prefix="diffusion_model."):
        """EMA weights keyed by the U-Net's own parameter names."""
        bufs = dict(self.named_buffers())
        return {name[len(prefix):]: bufs[s] for name, s in self.m_name2s_name.items() if name.startswith(prefix)}


class MinimalRETRODiffusion(nn.Module):
    def __init__(self, k_nn, query_key, retrieval_encoder_cfg, nn_encoder_cfg=None, query_encoder_cfg=None, nn_key='retro_conditioning',
                 retro_noise=False, retrieval_cfg=None, retro_conditioning_key=None, learn_nn_encoder=False, nn_memory=None,
                 n_patches_per_side=1, resize_patch_size=None, searcher_path=None, retro_concat=False, p_uncond=0., guidance_vex_shape=None,
                 # LatentDiffusion / DDPM keys (ldm)
                 unet_config=None, first_stage_config=None, cond_stage_config="__is_unconditional__", timesteps=1000, beta_schedule="linear",
                 linear_start=1e-4, linear_end=2e-2, image_size=256, channels=3, conditioning_key=None, scale_factor=1.0, scale_by_std=False,
                 parameterization="eps", use_ema=True, ckpt_path=None, ignore_keys=(), log_every_t=100, first_stage_key="image",
                 cond_stage_key="image", **unused):
        super().__init__()
        if nn_encoder_cfg or query_encoder_cfg or retro_conditioning_key is not None or retro_concat:
            raise NotImplementedError("nn_encoder / query_encoder / retro_conditioning_key / retro_concat are not used by the shipped RDM configs")
        assert beta_schedule == "linear" and parameterization == "eps" and not scale_by_std
        self.k_nn, self.query_key, self.nn_key = k_nn, query_key, nn_key
        self.image_size, self.channels, self.log_every_t = image_size, channels, log_every_t
        self.parameterization, self.scale_factor, self.num_timesteps = parameterization, scale_factor, int(timesteps)
        self.first_stage_key, self.cond_stage_key = first_stage_key, cond_stage_key
        self.n_patches_per_side, self.resize_nn_patch_size, self.retro_noise, self.p_uncond = n_patches_per_side, resize_patch_size, retro_noise, p_uncond
        if cond_stage_config == "__is_unconditional__":
            conditioning_key = None                                         # ldm LatentDiffusion
        self.model = RETRODiffusionWrapper(DiffusionWrapper(unet_config, conditioning_key), concat=False)
        self.use_ema = use_ema
        if self.use_ema:
            self.model_ema = LitEma(self.model)
            print(f"Keeping EMAs of {len(list(self.model_ema.buffers()))}.")
        # schedule buffers (ldm register_schedule, float64 -> float32)
        betas = _tables.make_beta_schedule(self.num_timesteps, linear_start, linear_end)
        ac = np.cumprod(1. - betas, axis=0)
        acp = np.append(1., ac[:-1])
        f32 = lambda a: torch.tensor(a, dtype=torch.float32)
        for name, val in dict(betas=betas, alphas_cumprod=ac, alphas_cumprod_prev=acp, sqrt_alphas_cumprod=np.sqrt(ac),
                              sqrt_one_minus_alphas_cumprod=np.sqrt(1. - ac), log_one_minus_alphas_cumprod=np.log(1. - ac),
                              sqrt_recip_alphas_cumprod=np.sqrt(1. / ac), sqrt_recipm1_alphas_cumprod=np.sqrt(1. / ac - 1)).items():
            self.register_buffer(name, f32(val))
        # first stage (decode only)
        self.first_stage_model = instantiate_from_config(first_stage_config) if first_stage_config else None
        if self.first_stage_model is not None:
            self.first_stage_model.eval(); self.first_stage_model.train = disabled_train.__get__(self.first_stage_model)
            for p in self.first_stage_model.parameters():
                p.requires_grad = False
        self.cond_stage_model = None
        # nn_memory (ddpm.py:166-176)
        self.searcher_path = searcher_path
        self.use_memory = nn_memory is not None and os.path.isfile(nn_memory)
        if self.use_memory:
            assert nn_memory.endswith('.p')
            with open(nn_memory, 'rb') as f:
                nn_data = pickle.load(f)
            self.register_buffer('nn_memory', torch.tensor(nn_data['nn_memory'], dtype=torch.int), persistent=False)
            self.id_count = nn_data['id_count']
        self.nn_encoder, self.resize_nn_patches, self.learn_nn_encoder = None, False, learn_nn_encoder
        self.retriever = None
        self.init_retriever(retrieval_cfg)
        self.conditional_retrieval_encoder = False
        self.retrieval_encoder = instantiate_from_config(retrieval_encoder_cfg)
        self.use_retriever_for_retro_cond = self.retriever is not None and not self.use_memory
        if guidance_vex_shape is None:
            guidance_vex_shape = (unet_config["params"]["context_dim"],)
        self.get_unconditional_guiding_vex(tuple(guidance_vex_shape))
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, list(ignore_keys))

    # ---- module plumbing -------------------------------------------------------------------------------------
    @property
    def device(self):
        return self.betas.device

    def _b200_unet(self):
        return self.model.diffusion_model

    def init_from_ckpt(self, path, ignore_keys=()):
        sd = torch.load(path, map_location="cpu")
        sd = sd.get("state_dict", sd)
        for k in list(sd.keys()):
            if any(k.startswith(ik) for ik in ignore_keys):
                del sd[k]
        missing, unexpected = self.load_state_dict(sd, strict=False)
        print(f"Restored from {path} with {len(missing)} missing and {len(unexpected)} unexpected keys")

    def init_retriever(self, cfg):
        if not cfg:
            self.retriever = None
            return
        self.retriever = instantiate_from_config(cfg)
        self.retriever.train = disabled_train

    def train_searcher(self):
        print("training searcher...")
        self.retriever.train_searcher(device=self.device if self.device.type == "cuda" else None)
        print("done training searcher")

    @contextmanager
    def ema_scope(self, context=None):
        """ldm ema_scope: sample with the EMA weights (ddpm.py:977).  The shadow tensors are handed to the engine directly;
        nothing is copied into the parameters and repeated calls do not re-upload."""
        unet = self.model.diffusion_model
        if self.use_ema:
            unet.use_weights(self.model_ema.state_dict_for("diffusion_model."), tag=("ema", int(self.model_ema.num_updates)) + tuple(
                b._version for b in self.model_ema.buffers()))
            if context is not None:
                print(f"{context}: Switched to EMA weights")
        try:
            yield None
        finally:
            if self.use_ema:
                unet.use_weights(None)
                if context is not None:
                    print(f"{context}: Restored training weights")

    # ---- diffusion pieces ----------------------------------------------------------------------------------------
    def q_sample(self, x_start, t, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        ext = lambda a: a[t].reshape(-1, *([1] * (x_start.ndim - 1)))
        return ext(self.sqrt_alphas_cumprod) * x_start + ext(self.sqrt_one_minus_alphas_cumprod) * noise

    def apply_model(self, x_noisy, t, cond, return_ids=False):
        if isinstance(cond, dict):
            return self.model(x_noisy, t, **cond)
        if not isinstance(cond, list):
            cond = [cond]
        return self.model(x_noisy, t, c_crossattn=cond)

    @torch.no_grad()
    def decode_first_stage(self, z, predict_cids=False, force_not_quantize=False):
        if self.first_stage_model is None:
            return z
        return self.first_stage_model.decode(1. / self.scale_factor * z, force_not_quantize=predict_cids or force_not_quantize)

    def get_unconditional_guiding_vex(self, vector_shape):
        print('Initializing unconditional guidance vector')
        self.register_buffer('unconditional_guidance_vex', torch.randn(vector_shape), persistent=True)

    @torch.no_grad()
    def get_unconditional_conditioning(self, shape, unconditional_guidance_label=None, k_nn=None, ignore_knn=False):
        k_nn = self.k_nn if k_nn is None else k_nn
        bs = shape[0]
        vex = self.unconditional_guidance_vex
        if unconditional_guidance_label is not None:
            sig = vex / torch.linalg.norm(vex.flatten()) * unconditional_guidance_label
            if sig.shape[0] != self.k_nn and not ignore_knn:
                sig = torch.stack([sig] * k_nn, dim=0)
            sig = torch.stack([sig] * bs, dim=0)
        else:
            sig = torch.stack([vex] * bs, dim=0)
        print(sig.shape)
        return sig

    # ---- retrieval (device resident) ---------------------------------------------------------------------------
    def _searcher(self):
        if self.retriever.searcher is None:
            self.train_searcher()
        return self.retriever.searcher

    @torch.no_grad()
    def get_nn_and_encoding(self, query, return_patches=False, k_nn=None, n_patches_per_side=None, return_query_patches=False):
        """`ddpm.py:263-316`, device resident: image batch in [-1, 1] -> n x n patches -> retriever (CLIP image tower) -> q / ||q|| ->
        exact kNN -> raw neighbour rows.  Returns {nn_key: float32 [b, n*n, k, d]} (+ 'query_patches').  Used by the per-step
        re-retrieval sampler (`ddim.py:355-380`, BASELINE cfg4)."""
        searcher = self._searcher()
        n_ptch = self.n_patches_per_side if n_patches_per_side is None else n_patches_per_side
        k_nn = self.k_nn if k_nn is None else k_nn
        query = query.to(self.device)
        if not isimage(query):
            query = query.permute(0, 3, 1, 2)                                              # 'b h w c -> b c h w'    ddpm.py:275
        side = query.shape[-1] // n_ptch
        queries = torch.stack([query[..., i * side:(i + 1) * side, j * side:(j + 1) * side] for i in range(n_ptch) for j in range(n_ptch)], dim=1)
        output = {}
        if return_query_patches:
            output['query_patches'] = queries[:, :, None]                                   # [b, n, 1, c, h, w] (not resized: no first-stage encoder here)
        queries = queries.reshape(-1, *queries.shape[2:]).contiguous().float()               # '(b n) c h w'           ddpm.py:292
        q_emb = self.retriever.retriever(queries).float()                                    # CLIP image encode       ddpm.py:294
        nns, _ = search_raw(searcher, q_emb, k_nn)                                           # q / ||q|| + search_batched   ddpm.py:297-298
        out = searcher.gather_device(nns)                                                    # data_pool['embedding'][nns]   ddpm.py:301
        output[self.nn_key] = out.reshape(query.shape[0], n_ptch ** 2, k_nn, out.shape[-1])  # '(b n) k d -> b n k d'
        output['nns'] = nns
        if return_patches or self.nn_encoder is not None:
            raise NotImplementedError("neighbour image patches need the patch dataset, which is outside the sampling hot path")
        return output

    def get_qids(self, memsize, N, qids=None, use_weights=False, verbose=False):
        if isinstance(memsize, float) and hasattr(self, 'nn_memory'):
            assert 0 < memsize <= 1., 'Require memsize in (0,1]'
            memsize = int(memsize * self.nn_memory.shape[0])
        if qids is None:
            if self.use_memory:
                memsize = min(memsize, self.nn_memory.shape[0])
                print(f'Top-M Sampling with memory size {memsize}')
                nn_mem = self.nn_memory.detach().cpu().numpy()[:memsize]
                ps = None
                if use_weights:
                    freqs = np.asarray([self.id_count[int(i)] for i in nn_mem])
                    ps = freqs / freqs.sum(keepdims=True)
                qids = np.random.choice(nn_mem, size=N, p=ps)
            else:
                print('Randomly sampling retrieval database entries')
                qids = np.random.choice(getattr(self.retriever, 'num_rows', None) or len(self.retriever.data_pool['embedding']), size=N)   # global row count, also when row-sharded
        else:
            assert qids.shape[0] == N
        if verbose:
            print(f'Sampled entries are {qids}')
        return qids

    @torch.no_grad()
    def sample_from_rdata(self, N, cond=None, return_nns=False, use_weights=False, qids=None, k_nn=None, memsize=100, verbose=False,
                          pre_loaded_patches=None, unconditional_guidance_scale=1., unconditional_guidance_label=None,
                          unconditional_retro_guidance_label=None, nn_embeddings=None, **kwargs):
        if cond is not None:
            raise NotImplementedError("extra conditionings next to the retrieved neighbours are not implemented")
        searcher = self._searcher()
        k_nn = self.k_nn if k_nn is None else k_nn
        qids = self.get_qids(memsize, N, qids=qids, use_weights=use_weights, verbose=verbose)
        out = SampleLogs()
        qd = torch.as_tensor(np.asarray(qids), dtype=torch.int64, device=self.device)
        if nn_embeddings is None:
            q = searcher.gather_device(qd)                                     # data_pool['embedding'][qids]      ddpm.py:897
            nns, _ = search_raw(searcher, q, k_nn)                             # q / ||q|| + search_batched(...)    ddpm.py:906-908
            retro_cond = searcher.gather_device(nns)                           # data_pool['embedding'][nns] fp32   ddpm.py:921
            out.extras['nns'] = nns
        else:
            retro_cond = nn_embeddings.to(self.device, torch.float32)
        if return_nns:
            raise NotImplementedError("return_nns needs the image patch dataset (load_patch_dataset), which is outside the sampling hot path")
        c = self.retrieval_encoder(retro_cond)
        uc = self.get_unconditional_conditioning(c.shape, unconditional_guidance_label=unconditional_retro_guidance_label, k_nn=k_nn)
        with self.ema_scope("Plotting"):
            samples, _ = self.sample_log(cond=c, batch_size=N, unconditional_guidance_scale=unconditional_guidance_scale,
                                         unconditional_conditioning=uc.to(self.device), **kwargs)
        out["samples_with_sampled_nns"] = self.decode_first_stage(samples)
        out.extras["latents"] = samples
        return out

    @torch.no_grad()
    def sample_with_query(self, query, cond=None, bs=None, k_nn=None, unconditional_guidance_scale=1., unconditional_guidance_label=None,
                          unconditional_retro_guidance_label=None, return_nns=False, n_reps=None, query_embedded=False, example_maps=None,
                          visualize_nns=True, omit_query=False, normalize=False, **kwargs):
        if cond is not None or example_maps is not None:
            raise NotImplementedError("extra conditionings / example maps are not implemented")
        searcher = self._searcher()
        if bs is None:
            bs = 1
        if isinstance(query, np.ndarray):
            query = torch.from_numpy(query)
        if not query_embedded:
            assert query.ndim in [3, 4], 'User defined query for sampling has to be an image or of batch of images'
            if query.ndim == 3:
                query = torch.stack([query] * bs, dim=0)
            elif query.shape[0] == 1 and bs > 1:
                query = query.expand(bs, *query.shape[1:]).contiguous()         # '1 h w c -> b h w c'   ddpm.py:715-716
            assert ischannellastimage(query) or isimage(query)
            q_emb = self.retriever.embed(query.to(self.device))                 # CLIP image encode (dsetbuilder.py:461-473)
        else:
            q_emb = query.to(self.device, torch.float32)
            if q_emb.shape[0] == 1 and bs > 1:
                q_emb = q_emb.expand(bs, -1).contiguous()
        k_nn = self.k_nn if k_nn is None else k_nn
        print(f'Query shape is {tuple(q_emb.shape)}')
        nns, _ = search_raw(searcher, q_emb, k_nn)                              # q / ||q|| + search_batched   dsetbuilder.py:487-490
        r_emb = searcher.gather_device(nns)                                     # dsetbuilder.py:493
        if normalize:
            q_emb = q_emb / q_emb.norm(dim=-1, keepdim=True)
            r_emb = r_emb / r_emb.norm(dim=-1, keepdim=True)
        if omit_query:
            retro_cond = r_emb
        else:
            retro_cond = torch.cat([q_emb[:, None], r_emb[:, :k_nn - 1]], dim=1)            # ddpm.py:775
        if n_reps is not None:
            retro_cond = torch.cat([retro_cond] * n_reps, dim=1)
        out = SampleLogs()
        out.extras['nns'] = nns
        if return_nns:
            raise NotImplementedError("return_nns needs the image patch dataset, which is outside the sampling hot path")
        c = self.retrieval_encoder(retro_cond.float())
        print(c.shape)
        uc = self.get_unconditional_conditioning(c.shape, unconditional_guidance_label=unconditional_retro_guidance_label, k_nn=k_nn)
        if n_reps is not None:
            uc = torch.cat([uc] * n_reps, dim=1)
        with self.ema_scope("Plotting"):
            samples, _ = self.sample_log(cond=c, batch_size=c.shape[0], unconditional_guidance_scale=unconditional_guidance_scale,
                                         unconditional_conditioning=uc.to(self.device), **kwargs)
        out["query_samples"] = self.decode_first_stage(samples)
        out.extras["latents"] = samples
        return out

    @torch.no_grad()
    def sample_log(self, cond, batch_size, ddim, ddim_steps, custom_shape=None, del_sampler=False, **kwargs):
        if not ddim:
            raise NotImplementedError("ancestral DDPM sampling is not implemented; the sampling scripts always pass ddim=True")
        ddim_sampler = DDIMSampler(self)
        shape = custom_shape if custom_shape is not None else (self.channels, self.image_size, self.image_size)
        ddim_steps = kwargs.pop('S', ddim_steps)
        verbose = kwargs.pop('verbose', False)
        for k in ('use_weights', 'memsize', 'return_nns', 'visualize_nns', 'omit_query', 'query_embedded'):
            kwargs.pop(k, None)                                                # caller-side options that flow through **kwargs
        return ddim_sampler.sample(S=ddim_steps, batch_size=batch_size, shape=shape, conditioning=cond, verbose=verbose, **kwargs)

    # ---- out of scope ------------------------------------------------------------------------------------------------
    def _training_only(self, *a, **k):
        raise NotImplementedError("training is outside the sampling hot path this package implements (SURVEY.md section 2)")

    shared_step = forward = p_losses = configure_optimizers = log_images = training_step = _training_only
