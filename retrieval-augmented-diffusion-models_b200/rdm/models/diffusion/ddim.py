"""`rdm.models.diffusion.ddim.DDIMSampler` -- the reference's sampler API (`rdm/models/diffusion/ddim.py:13-268`).

Fast path: when the eps-model is the B200 U-Net and no per-step host hook is requested, the whole loop runs as
`rdm_ddim_sample` (one captured CUDA graph per step: timestep fill, U-Net with CFG batch doubling, fused CFG+DDIM update;
cross-attention K/V of the step-invariant context projected once).  Anything else (masks, callbacks, score correctors,
foreign models) takes the generic per-step path: `model.apply_model` + the fused update kernel `rdm_ddim_step`.
"""
import numpy as np
import torch
from tqdm.auto import tqdm

from rdm_b200 import sampler as _tables
from rdm_b200.unet import ddim_step


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", **kwargs):
        super().__init__()
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule

    def register_buffer(self, name, attr):
        if isinstance(attr, torch.Tensor) and attr.device != self.model.device:
            attr = attr.to(self.model.device)          # the reference hard-codes "cuda" (ddim.py:21-25)
        setattr(self, name, attr)

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        assert ddim_discretize == "uniform", "only the uniform discretisation is used by the reference configs"
        alphas_cumprod = self.model.alphas_cumprod
        assert alphas_cumprod.shape[0] == self.ddpm_num_timesteps, 'alphas have to be defined for each timestep'
        to_torch = lambda x: x.clone().detach().to(torch.float32).to(self.model.device)
        ac = alphas_cumprod.detach().cpu().to(torch.float32)
        self.register_buffer('betas', to_torch(self.model.betas))
        self.register_buffer('alphas_cumprod', to_torch(alphas_cumprod))
        self.register_buffer('alphas_cumprod_prev', to_torch(self.model.alphas_cumprod_prev))
        self.register_buffer('sqrt_alphas_cumprod', to_torch(ac.sqrt()))
        self.register_buffer('sqrt_one_minus_alphas_cumprod', to_torch((1. - ac).sqrt()))
        self.register_buffer('log_one_minus_alphas_cumprod', to_torch((1. - ac).log()))
        self.register_buffer('sqrt_recip_alphas_cumprod', to_torch((1. / ac).sqrt()))
        self.register_buffer('sqrt_recipm1_alphas_cumprod', to_torch((1. / ac - 1).sqrt()))
        t = _tables.make_ddim_tables(ac, ddim_num_steps, ddim_eta, device=self.model.device)
        self.ddim_timesteps = t["ddim_timesteps"]
        self.register_buffer('ddim_sigmas', t["sigmas"])
        self.register_buffer('ddim_alphas', t["alphas"])
        self.register_buffer('ddim_alphas_prev', t["alphas_prev"])
        self.register_buffer('ddim_sqrt_one_minus_alphas', t["sqrt_one_minus_alphas"])
        self._t_order, self._coef = t["timesteps"], t["coef"]           # sampling order, for librdm_b200
        sig = ddim_eta * torch.sqrt((1 - self.alphas_cumprod_prev) / (1 - self.alphas_cumprod) * (1 - self.alphas_cumprod / self.alphas_cumprod_prev))
        self.register_buffer('ddim_sigmas_for_original_num_steps', sig)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None, quantize_x0=False,
               eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None, verbose=True,
               x_T=None, log_every_t=100, unconditional_guidance_scale=1., unconditional_conditioning=None, random_guiding='none',
               r_shape=None, retro_cond=None, return_neighbors=False, k_nn=None, ignore_noising=False, content_cond=None, style_cond=None,
               intermediates_to_cpu=False, **kwargs):
        if conditioning is not None:
            first = conditioning[list(conditioning.keys())[0]] if isinstance(conditioning, dict) else (conditioning[0] if isinstance(conditioning, list) else conditioning)
            if first.shape[0] != batch_size:
                print(f"Warning: Got {first.shape[0]} conditionings but batch-size is {batch_size}")
            if unconditional_guidance_scale > 1.:
                print(f'Using unconditonal diffusion guidance with scale {unconditional_guidance_scale}')
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        assert random_guiding in ['none', 'sampled', 'const'], f'Unknown random guidance option {random_guiding}'
        size = (batch_size,) + tuple(shape)
        print(f'Data shape for DDIM sampling is {size}, eta {eta}')
        samples, intermediates = self.ddim_sampling(
            cond=conditioning, shape=size, callback=callback, img_callback=img_callback, quantize_denoised=quantize_x0, mask=mask, x0=x0,
            ddim_use_original_steps=False, noise_dropout=noise_dropout, temperature=temperature, score_corrector=score_corrector,
            corrector_kwargs=corrector_kwargs, x_T=x_T, log_every_t=log_every_t, unconditional_guidance_scale=unconditional_guidance_scale,
            unconditional_conditioning=unconditional_conditioning, random_guiding=random_guiding, content_cond=content_cond, style_cond=style_cond,
            intermediates_to_cpu=intermediates_to_cpu)
        return samples.detach(), intermediates

    # ---- fused loop -----------------------------------------------------------------------------------------
    def _fused_engine(self, cond, unconditional_conditioning, scale):
        """The B200 U-Net engine with the context of this call set, or None when the generic path is required."""
        get = getattr(self.model, "_b200_unet", None)
        if get is None or cond is None or isinstance(cond, dict):
            return None
        c = cond[0] if isinstance(cond, (list, tuple)) and len(cond) == 1 else cond
        uc = unconditional_conditioning
        uc = uc[0] if isinstance(uc, (list, tuple)) and len(uc) == 1 else uc
        if not isinstance(c, torch.Tensor) or (scale > 1. and not isinstance(uc, torch.Tensor)):
            return None
        unet = get()
        ctx = torch.cat([c, uc], dim=0) if scale > 1. else c            # cat([c, uc]) ddim.py:232
        return unet.set_context(ctx.to(self.model.device, torch.float32), self.model.device)

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None, quantize_denoised=False,
                      mask=None, x0=None, img_callback=None, log_every_t=100, temperature=1., noise_dropout=0., score_corrector=None,
                      corrector_kwargs=None, unconditional_guidance_scale=1., unconditional_conditioning=None, random_guiding='none',
                      content_cond=None, style_cond=None, intermediates_to_cpu=False, **kwargs):
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T.to(device)
        if timesteps is None:
            timesteps = self.ddpm_num_timesteps if ddim_use_original_steps else self.ddim_timesteps
        elif not ddim_use_original_steps:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        intermediates = {'x_inter': [img], 'pred_x0': [img]}
        time_range = reversed(range(0, timesteps)) if ddim_use_original_steps else np.flip(timesteps)
        total_steps = timesteps if ddim_use_original_steps else timesteps.shape[0]
        print(f"Running DDIM Sampling with {total_steps} timesteps")
        keep = (lambda t: t.detach().cpu()) if intermediates_to_cpu else (lambda t: t)

        plain = (not ddim_use_original_steps and total_steps == self.ddim_timesteps.shape[0] and mask is None and callback is None
                 and img_callback is None and score_corrector is None and not quantize_denoised and noise_dropout == 0. and temperature == 1.
                 and random_guiding == 'none' and content_cond is None and style_cond is None)
        eng = self._fused_engine(cond, unconditional_conditioning, unconditional_guidance_scale) if plain else None
        if eng is not None:
            S, eta_on = total_steps, bool(np.any(np.asarray(self.ddim_sigmas) != 0))
            # the reference draws one randn per step even when sigma = 0 (ddim.py:226-227): keep the generator in step
            noise = torch.stack([torch.randn(shape, device=device) for _ in range(S)])
            img = img.to(torch.float32)
            marks = sorted({i for i in range(S) if (S - i - 1) % log_every_t == 0 or (S - i - 1) == S - 1})       # ddim.py:207
            done = 0
            for i in marks:
                img, p0 = eng.ddim_sample(img, self._t_order, self._coef, cfg_scale=float(unconditional_guidance_scale), first_step=done,
                                          num_steps=i + 1 - done, noise=noise.reshape(S, -1) if eta_on else None, want_pred_x0=True)
                done = i + 1
                intermediates['x_inter'].append(keep(img)); intermediates['pred_x0'].append(keep(p0))
            if done < S:
                img = eng.ddim_sample(img, self._t_order, self._coef, cfg_scale=float(unconditional_guidance_scale), first_step=done, num_steps=S - done,
                                      noise=noise.reshape(S, -1) if eta_on else None)
            return img, intermediates

        iterator = tqdm(time_range, desc='DDIM Sampler', total=total_steps)
        random_guider = None
        if random_guiding != 'none':
            random_guider = torch.clamp(torch.randn(shape, device=self.model.device), -1., 1.)
        for i, step in enumerate(iterator):
            index = total_steps - i - 1
            ts = torch.full((b,), int(step), device=device, dtype=torch.long)
            snr = self.ddim_alphas[index] / (1 - self.ddim_alphas[index])
            input_cond = cond
            if style_cond is not None and snr < 5.e-2:
                input_cond = style_cond
            if content_cond is not None and snr >= 5.e-2 and snr < 1.:
                input_cond = content_cond
            if mask is not None:
                assert x0 is not None
                img_orig = self.model.q_sample(x0, ts)
                img = img_orig * mask + (1. - mask) * img
            if random_guiding == 'sampled':
                random_guider = torch.clamp(torch.randn(shape, device=self.model.device), -1., 1.)
            img, pred_x0 = self.p_sample_ddim(img, input_cond, ts, index=index, use_original_steps=ddim_use_original_steps,
                                              quantize_denoised=quantize_denoised, temperature=temperature, noise_dropout=noise_dropout,
                                              score_corrector=score_corrector, corrector_kwargs=corrector_kwargs,
                                              unconditional_guidance_scale=unconditional_guidance_scale,
                                              unconditional_conditioning=unconditional_conditioning, random_guider=random_guider)
            if callback: callback(i)
            if img_callback: img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates['x_inter'].append(keep(img)); intermediates['pred_x0'].append(keep(pred_x0))
        return img, intermediates

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, use_original_steps=False, quantize_denoised=False, temperature=1., noise_dropout=0.,
                      score_corrector=None, corrector_kwargs=None, unconditional_guidance_scale=1., unconditional_conditioning=None, noise=None,
                      random_guider=None):
        b, device = x.shape[0], x.device
        assert unconditional_guidance_scale >= 1.
        if noise is None:
            noise = torch.randn(x.shape, device=device)
        cfg = unconditional_guidance_scale > 1.
        if cfg:
            assert unconditional_conditioning is not None
            if isinstance(c, (list, tuple)):
                combined_c = [torch.cat([a, u], dim=0) for a, u in zip(c, unconditional_conditioning)]
            else:
                combined_c = torch.cat([c, unconditional_conditioning], dim=0)
            eps = self.model.apply_model(torch.cat([x] * 2, dim=0), torch.cat([t] * 2, dim=0), combined_c)
        else:
            eps = self.model.apply_model(x, t, c)
        fused_ok = (score_corrector is None and not quantize_denoised and noise_dropout == 0. and x.is_cuda and x.dtype == torch.float32)
        if use_original_steps:
            alphas, alphas_prev = self.model.alphas_cumprod, self.model.alphas_cumprod_prev
            s1m_all, sigmas = self.model.sqrt_one_minus_alphas_cumprod, self.ddim_sigmas_for_original_num_steps
        else:
            alphas, alphas_prev, s1m_all, sigmas = self.ddim_alphas, self.ddim_alphas_prev, self.ddim_sqrt_one_minus_alphas, self.ddim_sigmas
        f32 = lambda v: torch.tensor(float(v), dtype=torch.float32)
        a_t, a_prev, sigma_t, s1m = f32(alphas[index]), f32(alphas_prev[index]), f32(sigmas[index]), f32(s1m_all[index])
        if fused_ok:
            coef = torch.stack([s1m, a_t.sqrt(), a_prev.sqrt(), (1. - a_prev - sigma_t ** 2).sqrt(), sigma_t * temperature]).to(device)
            return ddim_step(x, eps, coef, cfg_scale=float(unconditional_guidance_scale) if cfg else None,
                             noise=noise if float(sigma_t) != 0. else None)
        # generic tensor path (score correctors / quantisation / noise dropout): plain torch, as in the reference
        e_t = eps
        if cfg:
            e_t, e_u = eps[:b], eps[b:]
            e_t = e_u + unconditional_guidance_scale * (e_t - e_u)
        if score_corrector is not None:
            assert self.model.parameterization == "eps"
            e_t = score_corrector.modify_score(self.model, e_t, x, t, c, **corrector_kwargs)
        a_t, a_prev, sigma_t, s1m = (v.to(device) for v in (a_t, a_prev, sigma_t, s1m))
        pred_x0 = (x - s1m * e_t) / a_t.sqrt()
        if quantize_denoised:
            pred_x0, _, *_ = self.model.first_stage_model.quantize(pred_x0)
        dir_xt = (1. - a_prev - sigma_t ** 2).sqrt() * e_t
        noise = sigma_t * noise * temperature
        if noise_dropout > 0.:
            noise = torch.nn.functional.dropout(noise, p=noise_dropout)
        return a_prev.sqrt() * pred_x0 + dir_xt + noise, pred_x0


class DDIMRetroSampler(DDIMSampler):
    """Per-step re-retrieval sampler (`rdm/models/diffusion/ddim.py:270-415`, BASELINE cfg4): after every DDIM step the x0 prediction is
    decoded by the first stage, its patches are embedded by the CLIP image retriever, the k nearest database rows are looked up again and
    become the conditioning of the NEXT step.  Every stage runs on the device (U-Net, VQ decoder, CLIP tower, exact kNN, gather); the loop
    itself is host-driven because the context changes each step (cross-attention K/V are re-projected by `rdm_unet_set_context`).

    The reference class asserts a `PreNoiserRetroDiffusion` model that is not part of the repository (dead upstream), so the
    model hooks it reads are optional here: `pre_noise` (default False), `conditional_retrieval_encoder` (False), `adjust_support` (identity)."""

    def __init__(self, model, *args, **kwargs):
        super().__init__(model, *args, **kwargs)
        assert hasattr(model, "get_nn_and_encoding"), "the model must provide get_nn_and_encoding (MinimalRETRODiffusion)"

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, retro_cond=None, r_shape=None, eta=0., x_T=None, log_every_t=100, verbose=True,
               unconditional_guidance_scale=1., unconditional_conditioning=None, return_neighbors=False, k_nn=None, ignore_noising=False,
               callback=None, img_callback=None, mask=None, x0=None, **kwargs):
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        size = (batch_size,) + tuple(shape)
        samples, inter = self.ddim_sampling(conditioning, retro_cond, size, r_shape=r_shape, x_T=x_T, callback=callback, img_callback=img_callback,
                                            mask=mask, x0=x0, log_every_t=log_every_t, unconditional_guidance_scale=unconditional_guidance_scale,
                                            unconditional_conditioning=unconditional_conditioning, return_neighbors=return_neighbors, k_nn=k_nn,
                                            ignore_noising=ignore_noising)
        return samples.detach(), inter

    @torch.no_grad()
    def ddim_sampling(self, cond, retro_cond, shape, r_shape=None, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100, temperature=1., noise_dropout=0.,
                      score_corrector=None, corrector_kwargs=None, unconditional_guidance_scale=1., unconditional_conditioning=None,
                      return_neighbors=False, k_nn=None, ignore_noising=False):
        model, device = self.model, self.model.betas.device
        if return_neighbors:
            raise NotImplementedError("neighbour image patches need the patch dataset, which is outside the sampling hot path")
        if cond is not None:
            raise NotImplementedError("extra conditionings next to the retrieved neighbours are not implemented")
        pre_noise = bool(getattr(model, "pre_noise", False))
        assert not pre_noise and not getattr(model, "conditional_retrieval_encoder", False), "pre-noised / query-conditioned retrieval encoders are not part of the shipped models"
        adjust_support = getattr(model, "adjust_support", lambda t: t)
        b = shape[0]
        k_nn = model.k_nn if k_nn is None else k_nn
        img = torch.randn(shape, device=device) if x_T is None else x_T.to(device)
        if r_shape is None:
            r_shape = shape                                         # ddim.py:293-294
        if retro_cond is None:                                      # ddim.py:296-316 (pre_noise False): two draws, like the reference --
            rc = torch.randn(r_shape, device=device)                # one that only gives the encoder something to shape the conditioning from,
            r_enc = torch.randn_like(model.retrieval_encoder(rc))   # and the noise that IS the first conditioning
        else:
            retro_cond = retro_cond.to(device, torch.float32)
            model.retrieval_encoder(retro_cond)                     # evaluated and discarded by the reference (:309-316):
            r_enc = retro_cond                                      # the first step is conditioned on the RAW retro_cond
        if timesteps is None:
            timesteps = self.ddpm_num_timesteps if ddim_use_original_steps else self.ddim_timesteps
        elif not ddim_use_original_steps:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        intermediates = {'x_inter': [], 'pred_x0': [], 'nns': []}
        time_range = reversed(range(0, timesteps)) if ddim_use_original_steps else np.flip(timesteps)
        total_steps = timesteps if ddim_use_original_steps else timesteps.shape[0]
        print(f"Running DDIM Sampling with {total_steps} timesteps")
        for i, step in enumerate(tqdm(time_range, desc='DDIM Sampler', total=total_steps)):
            index = total_steps - i - 1
            ts = torch.full((b,), int(step), device=device, dtype=torch.long)
            if mask is not None:
                assert x0 is not None
                img = model.q_sample(x0, ts) * mask + (1. - mask) * img
            noise = torch.randn(shape, device=device)                                       # ddim.py:340
            img, pred_x0 = self.p_sample_ddim(img, [r_enc], ts, index=index, use_original_steps=ddim_use_original_steps,
                                              quantize_denoised=quantize_denoised, temperature=temperature, noise_dropout=noise_dropout,
                                              score_corrector=score_corrector, corrector_kwargs=corrector_kwargs,
                                              unconditional_guidance_scale=unconditional_guidance_scale,
                                              unconditional_conditioning=unconditional_conditioning, noise=noise)
            if retro_cond is None:
                px0 = model.decode_first_stage(pred_x0)                                     # ddim.py:357
                found = model.get_nn_and_encoding(px0, k_nn=k_nn)                           # ddim.py:358-361
                rc = found[model.nn_key]
                intermediates['nns'].append(found['nns'])
                if rc.ndim == 4:
                    rc = rc.reshape(rc.shape[0], rc.shape[1] * rc.shape[2], rc.shape[3])    # 'b n k d -> b (n k) d'   ddim.py:374
            else:
                rc = retro_cond
            r_enc = model.retrieval_encoder(rc)                                             # ddim.py:396
            r_enc = adjust_support(r_enc)                                                   # ddim.py:398-400
            if not ignore_noising:
                r_enc = model.q_sample(r_enc, ts)                                           # ddim.py:401-402
            if callback: callback(i)
            if img_callback: img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates['x_inter'].append(img)
                intermediates['pred_x0'].append(pred_x0)
        return img, intermediates
