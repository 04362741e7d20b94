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


class FakeBuilder:
    """Stand-in DatasetBuilder for the offline neighbour precompute (scripts/search_neighbors.py:380-450): `search_k_nearest` with the
    reference's result keys; the neighbours of query i are rows i, i+1, ... shifted per call (deterministic)."""
    k, searcher = 3, object()

    def __init__(self):
        self.pool = np.arange(50 * 4, dtype=np.float32).reshape(50, 4)
        self.calls = 0

    def search_k_nearest(self, queries, visualize=False, is_caption=False):
        n = len(queries)
        nns = (np.arange(n)[:, None] + np.arange(self.k)[None] + 7 * self.calls) % 50
        self.calls += 1
        return {"embeddings": self.pool[nns], "nns": nns, "img_ids": nns * 10, "patch_coords": np.zeros((n, self.k, 4), np.int32), "queries": queries}


class Loader(list):
    batch_size = 2


def precompute_scenario(search_nns, base):
    """The call sequence both implementations are put through; returns everything observable: return values and every file written."""
    import os
    import pickle
    os.makedirs(os.path.join(base, "embeddings"), exist_ok=True)
    b = FakeBuilder()
    grid2 = Loader({"patches": torch.zeros(2, 4, 8, 8, 3)} for _ in range(3))
    paths = search_nns(b, grid2, device="cpu", save=True, npatches_perside=2, base_savedir=base, start_id=10)
    with open(os.path.join(base, paths[12]), "wb") as f:               # a truncated file: the 1 x 1 pass must replace it
        f.write(b"garbage")
    grid1 = Loader({"patches": torch.zeros(2, 1, 16, 16, 3)} for _ in range(3))
    paths1 = search_nns(b, grid1, device="cpu", save=True, npatches_perside=1, base_savedir=base, start_id=10, nn_paths=dict(paths))
    counts = search_nns(b, Loader({"caption": ["a", "b"]} for _ in range(5)), device="cpu", mode="text", save=False, max_its=2)
    files = {}
    for name in sorted(os.listdir(os.path.join(base, "embeddings"))):
        with open(os.path.join(base, "embeddings", name), "rb") as f:
            files[name] = pickle.load(f)
    return {"paths": dict(paths), "paths_after_second_pass": dict(paths1), "counts": {int(k): int(v) for k, v in counts.items()}, "files": files}


class StubFirstStage(torch.nn.Module):
    """Stand-in VQGAN for the RARM flow: `quantize.get_codebook_entry(indices, shape)` (taming VectorQuantizer2, lookup side) over a fixed
    codebook and a closed-form `decode` (first three channels of the code image), so decoded "images" depend on every sampled id."""

    class _Quantize(torch.nn.Module):
        def __init__(self, n_e, e_dim):
            super().__init__()
            import ref_weights
            self.embedding = torch.nn.Embedding(n_e, e_dim)
            with torch.no_grad():
                self.embedding.weight.copy_(torch.from_numpy(ref_weights.tensor_for("stub_codebook", (n_e, e_dim), 61)) * float(e_dim) ** 0.5)

        def get_codebook_entry(self, indices, shape):
            z_q = self.embedding(indices)
            if shape is not None:
                z_q = z_q.view(shape).permute(0, 3, 1, 2).contiguous()
            return z_q

    def __init__(self, n_embed=48, embed_dim=8, **ignored):
        super().__init__()
        self.quantize = StubFirstStage._Quantize(n_embed, embed_dim)

    def decode(self, quant, *a, **k):
        return torch.tanh(quant[:, :3])
