"""CPU: control flow of the per-step re-retrieval sampler (rdm/models/diffusion/ddim.py DDIMRetroSampler; reference ddim.py:270-415) with a
stub model -- step i must be conditioned on what was retrieved from the decoded x0 prediction of step i-1, the first step on r_shape noise,
a fixed retro_cond must never trigger retrieval, and the DDIM arithmetic must equal the oracle's."""
import numpy as np
import torch

from oracle import ddim as oddim


class StubModel:
    """eps = 0.1 * x + mean(context) (broadcast); decode = identity; retrieval = a deterministic function of the decoded image."""
    num_timesteps, k_nn, nn_key, device = 1000, 2, "nn_embeddings", torch.device("cpu")

    def __init__(self):
        ac = torch.from_numpy(oddim.alphas_cumprod_f32())
        self.alphas_cumprod, self.betas = ac, torch.zeros(1000)
        self.alphas_cumprod_prev = torch.cat([torch.ones(1), ac[:-1]])
        self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod = ac.sqrt(), (1 - ac).sqrt()
        self.contexts, self.queries = [], []
        self.retrieval_encoder = lambda t, **kw: t

    def apply_model(self, x, t, cond):
        c = cond[0]
        self.contexts.append(c.clone())
        return 0.1 * x + c.mean(dim=(1, 2)).reshape(-1, 1, 1, 1)

    def decode_first_stage(self, z):
        return z

    def get_nn_and_encoding(self, img, k_nn=None, **kw):
        self.queries.append(img.clone())
        rc = img.mean(dim=(1, 2, 3)).reshape(-1, 1, 1, 1) * torch.ones(img.shape[0], 1, k_nn, 4)
        return {self.nn_key: rc, "nns": torch.zeros(img.shape[0], k_nn, dtype=torch.long)}

    def q_sample(self, x0, t, noise=None):
        raise AssertionError("ignore_noising=True must not noise the conditioning")


def test_each_step_uses_the_previous_steps_retrieval():
    from rdm.models.diffusion.ddim import DDIMRetroSampler
    m = StubModel()
    S, x_T = 4, torch.randn(2, 3, 4, 4, generator=torch.Generator().manual_seed(0))
    torch.manual_seed(5)
    r0 = torch.randn(2, 2, 4)
    torch.manual_seed(5)
    img, inter = DDIMRetroSampler(m).sample(S, 2, (3, 4, 4), r_shape=(2, 2, 4), x_T=x_T, log_every_t=1, k_nn=2, ignore_noising=True, verbose=False)
    assert len(m.contexts) == S and len(m.queries) == S and len(inter["nns"]) == S
    assert torch.equal(m.contexts[0], r0)                                                      # first conditioning: noise of shape r_shape
    sch = oddim.Schedule(S)
    x, ctx = x_T, r0
    for i in range(S):
        assert torch.allclose(m.contexts[i], ctx)
        e = 0.1 * x + ctx.mean(dim=(1, 2)).reshape(-1, 1, 1, 1)
        x, p0 = oddim.ddim_update(x, e, *sch.coeffs(S - i - 1))
        assert torch.allclose(m.queries[i], p0, atol=1e-6) and torch.allclose(inter["pred_x0"][i], p0, atol=1e-6)
        ctx = p0.mean(dim=(1, 2, 3)).reshape(-1, 1, 1) * torch.ones(2, 2, 4)                   # 'b n k d -> b (n k) d' of the stub's retrieval
    assert torch.allclose(img, x, atol=1e-6)


def test_fixed_retro_cond_never_retrieves():
    from rdm.models.diffusion.ddim import DDIMRetroSampler
    m = StubModel()
    rc = torch.randn(2, 2, 4, generator=torch.Generator().manual_seed(1))
    DDIMRetroSampler(m).sample(4, 2, (3, 4, 4), retro_cond=rc, x_T=torch.zeros(2, 3, 4, 4), ignore_noising=True, verbose=False)
    assert not m.queries and all(torch.equal(c, rc) for c in m.contexts)
