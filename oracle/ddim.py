"""Oracle DDIM sampler: torch-CPU fp32 restatement.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
``rdm/models/diffusion/ddim.py:27-56`` (make_schedule), ``:143-215`` (ddim_sampling),
``:218-268`` (p_sample_ddim with classifier-free guidance by batch doubling) and the
ldm helpers of SURVEY.md Appendix A (``make_beta_schedule("linear")``,
``make_ddim_timesteps("uniform")``, ``make_ddim_sampling_parameters``).
"""
import numpy as np
import torch


def make_beta_schedule(n_timestep=1000, linear_start=0.0015, linear_end=0.0195):
    """ldm linear schedule: betas = linspace(sqrt(s), sqrt(e), T, float64)**2."""
    return np.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=np.float64) ** 2


def alphas_cumprod_f32(n_timestep=1000, linear_start=0.0015, linear_end=0.0195):
    """The float32 ``alphas_cumprod`` buffer LatentDiffusion registers (float64 cumprod, cast)."""
    betas = make_beta_schedule(n_timestep, linear_start, linear_end)
    return np.cumprod(1.0 - betas, axis=0).astype(np.float32)


def make_ddim_timesteps(num_ddim, num_ddpm=1000):
    c = num_ddpm // num_ddim
    return np.asarray(list(range(0, num_ddpm, c))) + 1


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta):
    """float32 in → (sigmas, alphas, alphas_prev) exactly as numpy evaluates them on the
    float32 ``alphas_cumprod`` (ddim.py:44-46)."""
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


class Schedule:
    """Per-step coefficient table.  The reference keeps ``ddim_alphas`` (float32 ndarray),
    ``ddim_alphas_prev`` (float64 ndarray built from a python list), ``ddim_sigmas`` and
    ``ddim_sqrt_one_minus_alphas`` as numpy arrays and feeds single elements through
    ``torch.full_like(e_t, v)`` (ddim.py:253-256), i.e. every coefficient is rounded to
    float32 once and all later arithmetic (``.sqrt()``, ``1 - a_prev - sigma**2``) is float32
    tensor arithmetic.  ``coeffs(i)`` reproduces that."""

    def __init__(self, S, eta=0.0, alphas_cumprod=None):
        ac = alphas_cumprod_f32() if alphas_cumprod is None else np.asarray(alphas_cumprod, dtype=np.float32)
        self.timesteps = make_ddim_timesteps(S, ac.shape[0])
        self.sigmas, self.alphas, self.alphas_prev = make_ddim_sampling_parameters(ac, self.timesteps, eta)
        self.sqrt_one_minus_alphas = np.sqrt(1.0 - self.alphas)

    def coeffs(self, index):
        f = lambda v: torch.tensor(float(v), dtype=torch.float32)
        a_t, a_prev = f(self.alphas[index]), f(self.alphas_prev[index])
        sigma, s1m = f(self.sigmas[index]), f(self.sqrt_one_minus_alphas[index])
        return a_t, a_prev, sigma, s1m


def ddim_update(x, e_t, a_t, a_prev, sigma, s1m, noise=None, temperature=1.0):
    """ddim.py:258-267 in float32 tensor arithmetic."""
    pred_x0 = (x - s1m * e_t) / a_t.sqrt()
    dir_xt = (1.0 - a_prev - sigma ** 2).sqrt() * e_t
    x_prev = a_prev.sqrt() * pred_x0 + dir_xt
    if noise is not None:
        x_prev = x_prev + sigma * noise * temperature
    return x_prev, pred_x0


@torch.no_grad()
def ddim_sample(unet, x_T, cond, uncond=None, S=100, scale=1.0, eta=0.0, schedule=None, return_all=False, draw_noise_always=False):
    """ddim.py:143-215 + :218-268: eps-model = ``unet(x, t, context)``; CFG by batch doubling
    with the *conditional half first* (``cat([c, uc])``, ddim.py:232-238).  ``draw_noise_always`` consumes the global torch RNG
    like the reference does -- one ``randn(x.shape)`` per step BEFORE the model call, even when sigma = 0 (ddim.py:226-227)."""
    sch = schedule or Schedule(S, eta)
    x, b = x_T.clone(), x_T.shape[0]
    traj = []
    for i, step in enumerate(np.flip(sch.timesteps)):
        index = len(sch.timesteps) - i - 1
        ts = torch.full((b,), int(step), dtype=torch.long)
        noise = torch.randn(x.shape) if (eta > 0 or draw_noise_always) else None
        if scale > 1.0:
            out = unet(torch.cat([x] * 2), torch.cat([ts] * 2), torch.cat([cond, uncond]))
            e_c, e_u = out[:b], out[b:]
            e_t = e_u + scale * (e_c - e_u)
        else:
            e_t = unet(x, ts, cond)
        x, pred_x0 = ddim_update(x, e_t, *sch.coeffs(index), noise=noise if eta > 0 else None)
        if return_all:
            traj.append((x.clone(), pred_x0.clone()))
    return (x, traj) if return_all else x
