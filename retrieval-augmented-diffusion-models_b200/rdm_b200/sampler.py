"""DDIM schedule tables for the fused sampling loop.

Restates ``DDIMSampler.make_schedule`` (``rdm/models/diffusion/ddim.py:27-56``) with the ldm helpers of SURVEY.md
Appendix A, keeping the reference's numerics: ``alphas_cumprod`` is the float32 buffer of ``LatentDiffusion``; every
per-step coefficient is rounded to float32 once (``torch.full_like(e_t, v)``, ddim.py:253-256) and the derived values
(``a_t.sqrt()``, ``(1 - a_prev - sigma**2).sqrt()``) are float32 tensor arithmetic.  Host-side, tiny, NumPy/torch CPU.
"""
import numpy as np
import torch


def make_beta_schedule(n_timestep=1000, linear_start=1e-4, linear_end=2e-2):
    return np.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=np.float64) ** 2


def make_ddim_timesteps(num_ddim_timesteps, num_ddpm_timesteps=1000):
    c = num_ddpm_timesteps // num_ddim_timesteps
    return np.asarray(list(range(0, num_ddpm_timesteps, c))) + 1


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta):
    alphacums = np.asarray(alphacums, dtype=np.float32)
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


def make_ddim_tables(alphas_cumprod, S, eta=0.0, device="cpu"):
    """-> dict(ddim_timesteps, sigmas, alphas, alphas_prev, sqrt_one_minus_alphas [reference order, ascending t],
    timesteps int64 [S] and coef float32 [S, 8] IN SAMPLING ORDER (descending t) for librdm_b200)."""
    ac = np.asarray(alphas_cumprod.detach().cpu().numpy() if isinstance(alphas_cumprod, torch.Tensor) else alphas_cumprod, dtype=np.float32)
    ts = make_ddim_timesteps(S, ac.shape[0])
    sigmas, alphas, alphas_prev = make_ddim_sampling_parameters(ac, ts, eta)
    s1m = np.sqrt(1.0 - alphas)
    f32 = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), dtype=torch.float32)
    a_t, a_prev, sg, s1 = f32(alphas), f32(alphas_prev), f32(sigmas), f32(s1m)
    S = len(ts)                                           # range(0, T, T // S) has MORE than S entries when S does not divide T (e.g. 6 -> 7, 30 -> 31): the reference runs them all (ddim.py:164)
    coef = torch.zeros(S, 8, dtype=torch.float32)
    coef[:, 0], coef[:, 1], coef[:, 2] = s1, a_t.sqrt(), a_prev.sqrt()
    coef[:, 3], coef[:, 4] = (1.0 - a_prev - sg ** 2).sqrt(), sg
    order = np.arange(S)[::-1].copy()                     # index = total_steps - i - 1 (ddim.py:175)
    return dict(ddim_timesteps=ts, sigmas=sigmas, alphas=alphas, alphas_prev=alphas_prev, sqrt_one_minus_alphas=s1m,
                timesteps=torch.from_numpy(ts[order].astype(np.int64)).to(device), coef=coef[order].contiguous().to(device))


def alphas_cumprod_linear(timesteps=1000, linear_start=0.0015, linear_end=0.0195):
    """The float32 ``alphas_cumprod`` buffer for the shipped RDM configs (models/rdm/imagenet/config.yaml:7-11)."""
    betas = make_beta_schedule(timesteps, linear_start, linear_end)
    return torch.tensor(np.cumprod(1.0 - betas, axis=0), dtype=torch.float32)
