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
    torch.randn(2, 2, 4)                                                                       # ddim.py:297: drawn, used for its shape only
    r0 = torch.randn(2, 2, 4)                                                                  # ddim.py:316: the noise that is the first conditioning
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


def test_matches_the_reference_retro_sampler_including_rng_consumption():
    """tests/golden/ref_retro_sampler.npz: the REFERENCE's DDIMRetroSampler.ddim_sampling (ddim.py:270-415) over the closed-form model of
    tests/golden/retro_stub.py.  Same model, same seed of the global torch generator -> the product's sampler must route the same tensors
    (contexts seen by the eps-model, queries handed to retrieval, x0 predictions, final sample) and draw the same random numbers in the
    same order (x_T, the two draws of the first conditioning, one noise tensor per step, q_sample of the new conditioning)."""
    import os
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import retro_stub
    from rdm.models.diffusion.ddim import DDIMRetroSampler
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_retro_sampler.npz"))
    for tag in ("retrieve", "retrieve_quiet", "fixed"):
        m = retro_stub.RetroStub().setup()
        s = DDIMRetroSampler(m)
        s.make_schedule(ddim_num_steps=4, ddim_eta=float(g[f"{tag}:eta"]), verbose=False)
        rc = torch.from_numpy(g[f"{tag}:retro_cond"]) if f"{tag}:retro_cond" in g.files else None
        torch.manual_seed(int(g[f"{tag}:seed"]))
        img, inter = s.ddim_sampling(None, rc, (2, 3, 4, 4), r_shape=(2, 2, 4), x_T=None, log_every_t=1, k_nn=2, ignore_noising=bool(g[f"{tag}:ignore_noising"]))
        close = lambda a, b: torch.allclose(a, torch.from_numpy(b), rtol=1e-5, atol=1e-6)
        assert close(torch.stack(m.contexts), g[f"{tag}:contexts"]), tag
        if rc is None:
            assert close(torch.stack(m.queries), g[f"{tag}:queries"]), tag
        else:
            assert not m.queries
        assert close(torch.stack(inter["pred_x0"]), g[f"{tag}:pred_x0"]) and close(img, g[f"{tag}:img"]), tag
