"""A deterministic stand-in MODEL for the per-step re-retrieval sampler (`DDIMRetroSampler`, reference ddim.py:270-415), shared by the
golden generator (which feeds it to the REFERENCE's sampler) and tests/test_retro_sampler_host.py (which feeds it to the product's):
a closed-form eps-model, first stage, retrieval and q_sample, so that every tensor the sampler routes between them -- and every random
number it draws from the global torch generator -- is checked, independent of any U-Net."""
import numpy as np
import torch


class RetroStub:
    num_timesteps, k_nn, nn_key, n_patches_per_side = 1000, 2, "nn_embeddings", 1
    pre_noise, conditional_retrieval_encoder, parameterization = False, False, "eps"
    device = torch.device("cpu")

    def setup(self):
        betas = np.linspace(0.0015 ** 0.5, 0.0195 ** 0.5, 1000, dtype=np.float64) ** 2
        ac = np.cumprod(1.0 - betas, axis=0)
        f32 = lambda a: torch.tensor(a, dtype=torch.float32)
        self.betas, self.alphas_cumprod, self.alphas_cumprod_prev = f32(betas), f32(ac), f32(np.append(1.0, ac[:-1]))
        self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod = f32(np.sqrt(ac)), f32(np.sqrt(1.0 - ac))
        self.contexts, self.queries = [], []
        return self

    def retrieval_encoder(self, t, **kw):                     # not the identity, so a missing / doubled application shows
        return 0.5 * t + 0.1

    def adjust_support(self, t):
        return torch.clamp(t, -1.0, 1.0)

    def apply_model(self, x, t, cond):
        c = cond[0] if isinstance(cond, (list, tuple)) else cond
        self.contexts.append(c.clone())
        return 0.1 * x + c.mean(dim=(1, 2)).reshape(-1, 1, 1, 1) + 0.001 * t.float().reshape(-1, 1, 1, 1)

    def decode_first_stage(self, z):
        return 2.0 * z

    def get_nn_and_encoding(self, img, k_nn=None, **kw):
        k = self.k_nn if k_nn is None else k_nn
        self.queries.append(img.clone())
        base = img.mean(dim=(1, 2, 3)).reshape(-1, 1, 1, 1)
        rc = base * torch.arange(1, k * 4 + 1, dtype=torch.float32).reshape(1, 1, k, 4) / 10.0           # [b, n=1, k, d=4]
        return {self.nn_key: rc, "nns": torch.zeros(img.shape[0], k, dtype=torch.long)}

    def q_sample(self, x_start, t, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        shp = (-1,) + (1,) * (x_start.ndim - 1)
        return self.sqrt_alphas_cumprod[t].reshape(shp) * x_start + self.sqrt_one_minus_alphas_cumprod[t].reshape(shp) * noise


class PatchEmbedStub(torch.nn.Module):
    """Stand-in for the image retriever (`retriever.retriever`, CLIP in the reference): 4x4 average pooling of every channel, then a fixed
    linear map to 512 dimensions.  Shared by the golden generator and the tests of `get_nn_and_encoding`."""

    def __init__(self):
        super().__init__()
        import ref_weights
        self.register_buffer("w", torch.from_numpy(ref_weights.tensor_for("patch_embed_stub", (48, 512), 51)) * 30.0)

    def forward(self, x):
        pooled = torch.nn.functional.adaptive_avg_pool2d(x.float(), 4).reshape(x.shape[0], -1)
        return pooled @ self.w.to(pooled.device)
