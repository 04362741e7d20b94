"""Golden vectors from the REFERENCE's own hot-path code, run on CPU in the build container.

    python tests/golden/make_golden_ref.py        (writes tests/golden/ref_*.npz; needs /root/reference)

The reference's modules are imported UNMODIFIED from /root/reference; the un-vendored packages they import are replaced by the
stand-ins of tests/golden/ref_stubs.py (see its header for what that covers).  What the fixtures pin:

  ref_unet_tiny.npz   rdm/modules/diffusionmodules/openaimodel.py  UNetModel.__init__/forward (:66-317, :335-371),
                      TimestepEmbedSequential (:17-33); rdm/modules/attention.py SpatialTransformer (:122-196),
                      BasicTransformerBlock (:77-96), CrossAttention (:20-74)                          -> oracle/unet.py
  ref_ddim_tiny.npz   rdm/models/diffusion/ddim.py DDIMSampler.make_schedule/sample/ddim_sampling/p_sample_ddim (:27-268) with
                      classifier-free guidance, eta = 0 and eta = 0.5, over the same U-Net                -> oracle/ddim.py
  ref_pipeline_tiny.npz  rdm/models/diffusion/ddpm.py MinimalRETRODiffusion.sample_from_rdata / sample_with_query / get_qids /
                      get_unconditional_conditioning / apply_model / sample_log (:445-458, :647-686, :689-844, :847-875, :878-1011) with
                      EMA weights, over an exact brute-force searcher (what the reference builds for pools < 2e4 rows)   -> the mirror
  ref_sampler_options.npz  rdm/models/diffusion/ddim.py DDIMSampler.sample with inpainting mask / eta + temperature / noise dropout / style-content
                      switching / callbacks / sampler-drawn x_T / log_every_t, over the closed-form model                 -> the product's DDIMSampler (generic path)
  ref_retro_sampler.npz  rdm/models/diffusion/ddim.py DDIMRetroSampler.ddim_sampling (:270-415: per-step re-retrieval, BASELINE cfg4) over a
                      closed-form model: every tensor routed between eps-model / first stage / retrieval / q_sample and every draw from
                      the global torch generator                                                          -> the product's DDIMRetroSampler
  ref_dsetbuilder.npz rdm/data/retrieval_dataset/dsetbuilder.py DatasetBuilder.load_embeddings / train_searcher / embed / search_k_nearest over an
                      exact stand-in for scann's brute-force scorer                                       -> the product's DatasetBuilder
  ref_search_nns.p    scripts/search_neighbors.py search_nns / save_pkl (:355-450): file names, per-example pickle layout, merging of patch
                      grids, corrupt-file policy, neighbour histogram                                     -> rdm_b200/nn_precompute.py
  ref_rarm_small.npz  rdm/modules/attention.py RetrievalPatchTransformer (:199-272; discrete tokens, positional encodings, causal
                      self-attention, cross-attention to the retrieved vectors) and the sampling arithmetic of
                      rdm/models/autoregression/transformer.py LatentImageRETRO.sample (:224-270)         -> oracle/rarm.py

The only reference method replaced is DDIMSampler.register_buffer (ddim.py:21-25), which force-moves every buffer to "cuda"
(SURVEY F6); the override keeps them on the CPU and changes no arithmetic.  tests/test_oracle_ref_golden.py checks the oracle
against these files; the GPU tests check the CUDA path against the same files.  Nothing on the GPU box reads /root/reference.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_stubs  # noqa: E402
import ref_weights  # noqa: E402

UNET_CFG = dict(image_size=16, in_channels=4, out_channels=4, model_channels=64, attention_resolutions=[2, 4], num_res_blocks=1,
                channel_mult=[1, 2, 3], num_head_channels=32, use_spatial_transformer=True, transformer_depth=1, context_dim=512)
# = oracle.unet.TINY_UNET (every layer kind, non-aligned concat GroupNorm groups) + the reference's use_spatial_transformer switch
RARM_CFG = dict(in_channels=50, n_heads=2, d_head=64, depth=2, context_dim=128, positional_encodings=True, sequence_length=12,
                out_channels=48, cross_attend=True, causal=True, continuous=False)


def save(name, out):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def unet_and_ddim():
    from rdm.modules.diffusionmodules.openaimodel import UNetModel          # the reference's
    from rdm.models.diffusion.ddim import DDIMSampler                        # the reference's
    assert UNetModel.__module__.startswith("rdm.") and "/root/reference" in sys.modules[UNetModel.__module__].__file__
    net = ref_weights.fill_(UNetModel(**UNET_CFG), 11).eval()
    g = torch.Generator().manual_seed(12)
    x = torch.randn(4, 4, 16, 16, generator=g)
    t = torch.tensor([1, 501, 991, 250])
    ctx = torch.randn(4, 3, 512, generator=g)
    with torch.no_grad():
        y = net(x, t, context=[ctx])                                         # the list form RETRODiffusionWrapper passes (ddpm.py:128-131)
    xs, ts_, cs = torch.randn(2, 4, 8, 8, generator=g), torch.tensor([991, 42]), torch.randn(2, 2, 512, generator=g)       # a small case for the emulated CUDA path
    with torch.no_grad():
        ys = net(xs, ts_, context=[cs])
    out = {"cfg_json": np.array(repr(UNET_CFG)), "x": x.numpy(), "t": t.numpy(), "context": ctx.numpy(), "out": y.numpy(), "weight_seed": np.int64(11),
           "small:x": xs.numpy(), "small:t": ts_.numpy(), "small:context": cs.numpy(), "small:out": ys.numpy(),
           "sd_keys": np.array(list(net.state_dict().keys())), "n_params": np.int64(sum(p.numel() for p in net.parameters()))}
    save("ref_unet_tiny.npz", out)

    class CpuDDIMSampler(DDIMSampler):
        def register_buffer(self, name, attr):                               # ddim.py:21-25 without the forced .to("cuda")
            setattr(self, name, attr)

    class Model:                                                             # what DDIMSampler touches of LatentDiffusion
        num_timesteps = 1000
        device = torch.device("cpu")
        parameterization = "eps"

        def __init__(self):
            betas = ref_stubs.make_beta_schedule("linear", 1000, linear_start=0.0015, linear_end=0.0195)      # config.yaml:7-11
            ac = np.cumprod(1.0 - betas, axis=0)
            f32 = lambda a: torch.tensor(a, dtype=torch.float32)             # ldm DDPM.register_schedule: to_torch = float32
            self.betas, self.alphas_cumprod, self.alphas_cumprod_prev = f32(betas), f32(ac), f32(np.append(1.0, ac[:-1]))

        def apply_model(self, x_noisy, t, cond):                             # ddpm.py:445-458 -> :128-131
            return net(x_noisy, t, context=[cond])

    xT = torch.randn(2, 4, 16, 16, generator=g)
    c, uc = torch.randn(2, 3, 512, generator=g), torch.zeros(2, 3, 512)
    out = {"x_T": xT.numpy(), "cond": c.numpy(), "uncond": uc.numpy(), "scale": np.float32(2.0), "S": np.int64(5)}
    for tag, eta, seed in (("eta0", 0.0, 0), ("eta05", 0.5, 7)):
        torch.manual_seed(seed)
        s = CpuDDIMSampler(Model())
        samples, inter = s.sample(5, 2, (4, 16, 16), conditioning=c, eta=eta, x_T=xT.clone(), verbose=False, log_every_t=1,
                                  unconditional_guidance_scale=2.0, unconditional_conditioning=uc)
        out[f"{tag}:samples"] = samples.numpy()
        out[f"{tag}:x_inter"] = torch.stack(inter["x_inter"][1:]).numpy()
        out[f"{tag}:pred_x0"] = torch.stack(inter["pred_x0"][1:]).numpy()
        out[f"{tag}:seed"] = np.int64(seed)
        if eta == 0.0:
            out["ddim_timesteps"] = np.asarray(s.ddim_timesteps)
            out["ddim_alphas"] = np.asarray(s.ddim_alphas)
            out["ddim_alphas_prev"] = np.asarray(s.ddim_alphas_prev)
        else:
            out["eta05:ddim_sigmas"] = np.asarray(s.ddim_sigmas)
    # plain conditional sampling (scale 1: no batch doubling, ddim.py:239-240)
    torch.manual_seed(0)
    s = CpuDDIMSampler(Model())
    samples, _ = s.sample(4, 2, (4, 16, 16), conditioning=c, eta=0.0, x_T=xT.clone(), verbose=False)
    out["noguid:samples"] = samples.numpy()
    save("ref_ddim_tiny.npz", out)


def rarm():
    from rdm.modules.attention import RetrievalPatchTransformer                # the reference's
    net = ref_weights.fill_(RetrievalPatchTransformer(**RARM_CFG), 21).eval()
    g = torch.Generator().manual_seed(22)
    tok = torch.randint(0, 50, (3, 12), generator=g)
    tok[:, 0] = 49                                                             # sos id = last vocabulary entry (config.yaml: sos_token 16385 of 16386)
    ctx = torch.randn(3, 4, 128, generator=g)
    with torch.no_grad():
        logits = net(tok, context=ctx)
        logits_short = net(tok[:, :5], context=ctx)                            # a prefix: what step 4 of the sampling loop evaluates
        logits_uncond = net(tok, context=torch.zeros_like(ctx))
    out = {"cfg_json": np.array(repr(RARM_CFG)), "tokens": tok.numpy(), "context": ctx.numpy(), "logits": logits.numpy(),
           "logits_prefix5": logits_short.numpy(), "logits_uncond": logits_uncond.numpy(), "weight_seed": np.int64(21),
           "sd_keys": np.array(list(net.state_dict().keys())), "n_params": np.int64(sum(p.numel() for p in net.parameters()))}
    save("ref_rarm_small.npz", out)


N_DB, K_NN = 600, 4


class ExactSearcher:
    """What `scann.scann_ops_pybind.builder(db / |db|, k, "dot_product").score_brute_force().build()` computes for pools < 2e4
    (dsetbuilder.py:574,590-592): exact top-k by dot product on unit rows, best first."""

    def __init__(self, emb):
        e = emb.astype(np.float64)
        self.unit = e / np.linalg.norm(e, axis=1)[:, None]

    def search_batched(self, q, final_num_neighbors=None):
        s = q.astype(np.float64) @ self.unit.T
        idx = np.argsort(-s, axis=1, kind="stable")[:, :final_num_neighbors]
        return idx.astype(np.uint32), np.take_along_axis(s, idx, 1).astype(np.float32)


class Retriever:
    """The DatasetBuilder surface `sample_from_rdata` / `sample_with_query` touch; `search_k_nearest` follows dsetbuilder.py:478-518
    for already-embedded queries (the reference's own dsetbuilder.py cannot be imported: scann, streamlit, the image datasets)."""
    load_patch_dataset = False

    def __init__(self, emb):
        self.data_pool = {"embedding": emb, "img_id": np.arange(len(emb)), "patch_coords": np.zeros((len(emb), 4), np.int32)}
        self.searcher = None
        self.retriever = torch.nn.Identity()

    def train_searcher(self):
        self.searcher = ExactSearcher(self.data_pool["embedding"])

    def search_k_nearest(self, queries, k=None, is_caption=False, visualize=None, query_embedded=False):
        assert query_embedded
        q_emb_ = queries
        query_embeddings = q_emb_ / np.linalg.norm(q_emb_, axis=1)[:, np.newaxis]
        nns, distances = self.searcher.search_batched(query_embeddings, final_num_neighbors=k)
        return {"embeddings": self.data_pool["embedding"][nns], "img_ids": self.data_pool["img_id"][nns], "patch_coords": self.data_pool["patch_coords"][nns],
                "queries": queries, "exec_time": 0.0, "nns": nns, "q_embeddings": q_emb_}


def pipeline():
    """MinimalRETRODiffusion.sample_from_rdata / sample_with_query / get_qids / get_unconditional_conditioning / apply_model / sample_log
    (ddpm.py:445-458,647-686,689-844,847-875,878-1011) run end to end on CPU."""
    import pickle
    import tempfile
    import rdm.models.diffusion.ddpm as ref_ddpm                              # the reference's
    from rdm.models.diffusion.ddim import DDIMSampler

    class CpuDDIMSampler(DDIMSampler):
        def register_buffer(self, name, attr):
            setattr(self, name, attr)
    ref_ddpm.DDIMSampler = CpuDDIMSampler                                      # sample_log constructs DDIMSampler(self) (ddpm.py:993)

    db, mem, id_count = ref_weights.make_db(N_DB)
    with tempfile.TemporaryDirectory() as td:
        mem_path = os.path.join(td, "nn_memory.p")
        with open(mem_path, "wb") as f:
            pickle.dump({"nn_memory": mem, "id_count": id_count}, f)
        A = ref_stubs.AttrDict
        unet = dict(UNET_CFG)
        model = ref_ddpm.MinimalRETRODiffusion(
            k_nn=K_NN, query_key="clip_img_emb", retrieval_encoder_cfg=A(target="torch.nn.Identity"), nn_memory=mem_path, retrieval_cfg=None,
            unet_config=A(target="rdm.modules.diffusionmodules.openaimodel.UNetModel", params=unet), first_stage_config=None,
            cond_stage_config="__is_unconditional__", timesteps=1000, linear_start=0.0015, linear_end=0.0195, image_size=16, channels=4,
            conditioning_key="crossattn", log_every_t=100).eval()
    model.retriever = Retriever(db)
    # checkpoint layout: live weights under model.diffusion_model.*, EMA shadows under model_ema.<name without dots> (ddpm.py:162-164)
    live = ref_weights.state_dict_for(((k, v.shape) for k, v in model.model.diffusion_model.state_dict().items()), 11)
    ema = ref_weights.state_dict_for(((k, v.shape) for k, v in model.model.diffusion_model.state_dict().items()), 12)
    sd = {"model.diffusion_model." + k: v for k, v in live.items()}
    sd.update({"model_ema." + ("diffusion_model." + k).replace(".", ""): v for k, v in ema.items()})
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    out = {"n_db": np.int64(N_DB), "k_nn": np.int64(K_NN), "live_seed": np.int64(11), "ema_seed": np.int64(12), "missing_keys": np.array(sorted(missing)),
           "nn_memory": mem, "id_count_keys": np.array(list(id_count.keys())), "id_count_vals": np.array(list(id_count.values()))}
    g = torch.Generator().manual_seed(43)
    xT = torch.randn(3, 4, 16, 16, generator=g)
    common = dict(unconditional_guidance_scale=2.0, ddim_steps=4, ddim=True, unconditional_retro_guidance_label=0.)
    # (1) scripts/rdm_sample.py:226-262: top-m pseudo-queries from the NumPy global RNG, then sample_from_rdata
    np.random.seed(44)
    qids = model.get_qids(50, 3, use_weights=False)
    np.random.seed(44)
    logs = model.sample_from_rdata(3, qids=None, k_nn=K_NN, use_weights=False, memsize=50, x_T=xT.clone(), **common)
    out["rdata:qids"], out["rdata:x_T"], out["rdata:samples"] = qids, xT.numpy(), logs["samples_with_sampled_nns"].numpy()
    np.random.seed(45)
    out["qids_weighted"] = model.get_qids(0.4, 5, use_weights=True)
    # (1b) a small case for the product's emulated CUDA path: one image, two steps, an 8 x 8 latent through `custom_shape` (ddpm.py:994-1001)
    xs = torch.randn(1, 4, 8, 8, generator=g)
    logs = model.sample_from_rdata(1, qids=np.array([123]), k_nn=K_NN, x_T=xs.clone(), custom_shape=(4, 8, 8), unconditional_guidance_scale=2.0,
                                   ddim_steps=2, ddim=True, unconditional_retro_guidance_label=0.)
    out["rdata_small:x_T"], out["rdata_small:samples"] = xs.numpy(), logs["samples_with_sampled_nns"].numpy()
    # (2) scripts/rdm_sample.py:272-299: an embedded (text) query, prepended as neighbour 0; and omit_query
    q = torch.from_numpy(ref_weights.tensor_for("query", (2, 512), 46) * 22.0)
    for tag, extra in (("query", dict(omit_query=False)), ("query_omit", dict(omit_query=True)), ("query_normalize", dict(normalize=True)),
                       ("query_reps", dict(n_reps=2)), ("query_single", dict(bs=2, single=True))):
        extra = dict(extra)
        qq = q[:1] if extra.pop("single", False) else q                       # one embedded query repeated over the batch (ddpm.py:717-718)
        logs = model.sample_with_query(query=qq, query_embedded=True, k_nn=K_NN, visualize_nns=False, x_T=xT[:2].clone(), **common, **extra)
        out[f"{tag}:samples"] = logs["query_samples"].numpy()
    out["query:q"] = q.numpy()
    # (2b) get_nn_and_encoding (ddpm.py:263-316): image -> n x n patches -> retriever -> q / |q| -> kNN -> RAW rows [b, n*n, k, d]
    import retro_stub
    model.retriever.retriever = retro_stub.PatchEmbedStub()
    imgs = torch.from_numpy(ref_weights.tensor_for("nn_enc_images", (2, 3, 16, 16), 48) * 8.0)
    for tag, x, n in (("nnenc_2x2", imgs, 2), ("nnenc_1x1_channels_last", imgs.permute(0, 2, 3, 1).contiguous(), 1)):
        r = model.get_nn_and_encoding(x, k_nn=3, n_patches_per_side=n)
        out[f"{tag}:nn_embeddings"] = r[model.nn_key].numpy()
    out["nnenc:images"] = imgs.numpy()
    # (3) unconditional conditioning for a non-zero label (ddpm.py:663-686): vex / |vex| * label, stacked [bs, k, d]
    model.unconditional_guidance_vex.copy_(torch.from_numpy(ref_weights.tensor_for("vex", (512,), 47)))
    out["uncond_label_1.5"] = model.get_unconditional_conditioning((2, K_NN, 512), unconditional_guidance_label=1.5, k_nn=K_NN).numpy()
    out["vex"] = model.unconditional_guidance_vex.numpy().copy()
    save("ref_pipeline_tiny.npz", out)


def retro_sampler():
    """DDIMRetroSampler.ddim_sampling (ddim.py:270-415: per-step re-retrieval) over the closed-form model of tests/golden/retro_stub.py.
    `sample()` cannot be used: the base class passes keyword arguments this subclass does not accept (dead upstream).  The class
    asserts a `PreNoiserRetroDiffusion` model from ldm that the repository does not define: the stand-in below is that marker type."""
    import ldm.models.diffusion.ddpm as ldm_ddpm
    import retro_stub
    from rdm.models.diffusion.ddim import DDIMRetroSampler

    class PreNoiserRetroDiffusion(object):
        pass
    ldm_ddpm.PreNoiserRetroDiffusion = PreNoiserRetroDiffusion

    class Model(PreNoiserRetroDiffusion, retro_stub.RetroStub):
        pass

    class CpuRetroSampler(DDIMRetroSampler):
        def register_buffer(self, name, attr):
            setattr(self, name, attr)

    out = {}
    for tag, kw in (("retrieve", dict(retro_cond=None, ignore_noising=False, eta=0.3, seed=3)),
                    ("retrieve_quiet", dict(retro_cond=None, ignore_noising=True, eta=0.0, seed=4)),
                    ("fixed", dict(retro_cond=torch.randn(2, 2, 4, generator=torch.Generator().manual_seed(9)), ignore_noising=True, eta=0.0, seed=5))):
        m = Model().setup()
        s = CpuRetroSampler(m)
        s.make_schedule(ddim_num_steps=4, ddim_eta=kw["eta"], verbose=False)
        torch.manual_seed(kw["seed"])
        img, inter = s.ddim_sampling(None, kw["retro_cond"], (2, 3, 4, 4), r_shape=(2, 2, 4), x_T=None, log_every_t=1, k_nn=2, ignore_noising=kw["ignore_noising"])
        out[f"{tag}:img"] = img.numpy()
        out[f"{tag}:pred_x0"] = torch.stack(inter["pred_x0"]).numpy()
        out[f"{tag}:contexts"] = torch.stack(m.contexts).numpy()
        out[f"{tag}:queries"] = torch.stack(m.queries).numpy() if m.queries else np.zeros((0,))
        out[f"{tag}:eta"], out[f"{tag}:seed"], out[f"{tag}:ignore_noising"] = np.float32(kw["eta"]), np.int64(kw["seed"]), np.bool_(kw["ignore_noising"])
        if kw["retro_cond"] is not None:
            out[f"{tag}:retro_cond"] = kw["retro_cond"].numpy()
    save("ref_retro_sampler.npz", out)


SAMPLER_CASES = {       # keyword arguments of DDIMSampler.sample (ddim.py:59-91) beyond the plain guided loop
    "mask_eta": dict(S=5, eta=0.5, mask=True, temperature=0.8, log_every_t=2, seed=11),
    "noise_dropout": dict(S=4, eta=1.0, noise_dropout=0.25, log_every_t=100, seed=12),
    "style_content": dict(S=10, eta=0.0, style=True, log_every_t=3, seed=13, guidance=1.0),
    "callbacks_xT_drawn": dict(S=4, eta=0.2, callbacks=True, draw_xT=True, log_every_t=1, seed=14),
}


def run_sampler_case(sampler_cls, model, kw, retro_stub):
    """One DDIMSampler.sample call described by a SAMPLER_CASES entry -> dict of results (shared by the generator and the test)."""
    g = torch.Generator().manual_seed(kw["seed"] + 100)
    shape = (3, 4, 4)
    c, uc = torch.randn(2, 2, 4, generator=g), torch.zeros(2, 2, 4)
    xT = None if kw.get("draw_xT") else torch.randn(2, *shape, generator=g)
    extra, seen = {}, []
    if kw.get("mask"):
        extra["mask"] = (torch.rand(2, 1, 4, 4, generator=g) > 0.5).float()
        extra["x0"] = torch.randn(2, *shape, generator=g)
    if kw.get("style"):
        extra["style_cond"], extra["content_cond"] = torch.randn(2, 2, 4, generator=g), torch.randn(2, 2, 4, generator=g)
    if kw.get("callbacks"):
        extra["callback"] = lambda i: seen.append(("cb", int(i)))
        extra["img_callback"] = lambda p0, i: seen.append(("img", int(i), float(p0.sum())))
    scale = kw.get("guidance", 2.0)
    torch.manual_seed(kw["seed"])
    s = sampler_cls(model)
    samples, inter = s.sample(kw["S"], 2, shape, conditioning=c, eta=kw["eta"], x_T=xT, verbose=False, log_every_t=kw["log_every_t"],
                              temperature=kw.get("temperature", 1.0), noise_dropout=kw.get("noise_dropout", 0.0), unconditional_guidance_scale=scale,
                              unconditional_conditioning=uc if scale > 1.0 else None, **extra)
    return {"samples": samples.numpy(), "x_inter": torch.stack([t for t in inter["x_inter"]]).numpy(), "pred_x0": torch.stack([t for t in inter["pred_x0"]]).numpy(),
            "contexts": torch.stack(model.contexts).numpy(), "after": torch.rand(3).numpy(),       # `after`: the generator state the call leaves behind
            "seen": np.array([list(map(float, e[1:])) + [0.0] * (3 - len(e)) for e in seen]) if seen else np.zeros((0, 2))}


def sampler_options():
    """DDIMSampler.sample / ddim_sampling / p_sample_ddim (ddim.py:59-268) with the options outside the plain guided loop -- inpainting mask,
    eta > 0 with temperature, noise dropout, style / content conditioning by SNR, callbacks, x_T drawn by the sampler, intermediates every
    log_every_t -- over the closed-form model of tests/golden/retro_stub.py; includes what the call leaves in the global torch generator."""
    import retro_stub
    from rdm.models.diffusion.ddim import DDIMSampler

    class CpuDDIMSampler(DDIMSampler):
        def register_buffer(self, name, attr):
            setattr(self, name, attr)
    out = {}
    for tag, kw in SAMPLER_CASES.items():
        for k, v in run_sampler_case(CpuDDIMSampler, retro_stub.RetroStub().setup(), kw, retro_stub).items():
            out[f"{tag}:{k}"] = v
    save("ref_sampler_options.npz", out)


def neighbour_precompute():
    """scripts/search_neighbors.py `search_nns` / `save_pkl` (:355-450) -- the writer of the per-example neighbour pickles that
    QueryDataset.load_nns reads -- over the stand-in builder of tests/golden/retro_stub.py.  The script's module-level imports of the
    database builder (scann, streamlit, image datasets) and kornia are satisfied by empty stand-ins; the two functions run unmodified."""
    import importlib.util
    import pickle
    import tempfile
    import types
    import retro_stub
    m = types.ModuleType("rdm.data.retrieval_dataset.dsetbuilder"); m.DatasetBuilder = object
    sys.modules["rdm.data.retrieval_dataset.dsetbuilder"] = m
    for name, attrs in (("kornia.geometry", {}), ("kornia.geometry.transform", {"crop_by_boxes": None})):
        k = types.ModuleType(name); k.__dict__.update(attrs); sys.modules[name] = k
    sys.modules["ldm.util"].parallel_data_prefetch = None
    sys.modules["omegaconf"].OmegaConf = object
    spec = importlib.util.spec_from_file_location("ref_search_neighbors", "/root/reference/scripts/search_neighbors.py")
    script = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(script)
    with tempfile.TemporaryDirectory() as td:
        res = retro_stub.precompute_scenario(script.search_nns, td)
    path = os.path.join(HERE, "ref_search_nns.p")
    with open(path, "wb") as f:
        pickle.dump(res, f, protocol=4)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


RARM_MODEL_CFG = dict(mask_token=48, sos_token=49, p_mask_max=0.0, nn_key="nn_embeddings", nn_memory=None, retrieval_cfg=None,
                      cond_stage_config="__is_unconditional__")
RARM_SAMPLING_CASES = {"guided": dict(guidance_scale=2.5, temperature=0.9, top_k=8, seed=71), "plain": dict(guidance_scale=1.0, temperature=1.3, top_k=20, seed=72)}


def rarm_model_cfg(A=dict):
    return dict(RARM_MODEL_CFG, nn_reshaper_cfg=A(target="rdm.modules.encoders.nn_encoders.CLIPEmbeddingReshaper"),
                nn_encoder_cfg=A(target="rdm.modules.encoders.nn_encoders.IdentityEncoder"),
                transformer_config=A(target="rdm.modules.attention.RetrievalPatchTransformer", params=dict(RARM_CFG)),
                first_stage_config=A(target="retro_stub.StubFirstStage", params=dict(n_embed=48, embed_dim=8)))


def rarm_sampling():
    """LatentImageRETRO.sample / sampling_util / sample_from_rdata (rdm/models/autoregression/transformer.py:224-391) run end to end over the
    reference's own RetrievalPatchTransformer: guidance on the logits, temperature, top-k filter, softmax, prefix handling, SOS conditioning,
    decode_to_img.  The one replaced call is `torch.multinomial` (its draws cannot be reproduced by a device kernel): it is swapped for the
    inverse-CDF draw on recorded uniforms -- the definition the oracle and the CUDA sampler use -- and the probabilities it is handed are
    recorded, so the fixture holds, per step, exactly what the reference computed."""
    import retro_stub  # noqa: F401
    from rdm.models.autoregression.transformer import LatentImageRETRO           # the reference's
    A = ref_stubs.AttrDict
    model = LatentImageRETRO(**rarm_model_cfg(A)).eval()
    ref_weights.fill_(model.transformer, 21)
    db, _, _ = ref_weights.make_db(N_DB)
    db128 = db[:, :128].copy()                                                   # context_dim of the small decoder
    model.retriever = Retriever(db128)
    out = {"n_db": np.int64(N_DB), "sd_keys": np.array([k for k in model.state_dict().keys()])}
    real_multinomial = torch.multinomial
    for tag, kw in RARM_SAMPLING_CASES.items():
        torch.manual_seed(kw["seed"])
        u = torch.rand((9, 2))                                                   # what the product's sampler draws (torch.rand((steps, B)))
        rec = {"probs": [], "step": 0}

        def fake_multinomial(probs, num_samples=1, **k):
            rec["probs"].append(probs.clone())
            cdf = probs.double().cumsum(-1)
            ix = (cdf > (u[rec["step"]].double() * cdf[:, -1])[:, None]).float().argmax(-1)
            rec["step"] += 1
            return ix[:, None]
        torch.multinomial = fake_multinomial
        try:
            np.random.seed(kw["seed"])
            logs = model.sample_from_rdata(2, qids=None, k_nn=4, memsize=100, top_k=kw["top_k"], temperature=kw["temperature"], code_side_len=3,
                                           z_dimensionality=8, guidance_scale=kw["guidance_scale"])
        finally:
            torch.multinomial = real_multinomial
        out[f"{tag}:qids"], out[f"{tag}:uniforms"] = np.asarray(logs["qids"]), u.numpy()
        out[f"{tag}:probs"] = torch.stack(rec["probs"]).numpy()
        out[f"{tag}:images"] = logs["samples_with_sampled_nns"].numpy()
    # greedy decoding of a given start prefix through `sample` itself (sample=False, transformer.py:265-266)
    g = torch.Generator().manual_seed(73)
    r = torch.randn(2, 4, 128, generator=g)
    _, c = model.encode_to_c(torch.zeros((2, 0)))
    start = torch.randint(0, 48, (2, 3), generator=g)
    out["greedy:r"], out["greedy:start"] = r.numpy(), start.numpy()
    out["greedy:tokens"] = model.sample(start, r, c, steps=6, sample=False, top_k=None, guidance_scale=3.0).numpy()
    save("ref_rarm_sampling.npz", out)


def dataset_builder():
    """rdm/data/retrieval_dataset/dsetbuilder.py DatasetBuilder.load_embeddings / train_searcher / embed / search_k_nearest (:181-236,
    :461-518, :534-619), unmodified.  The module's imports of scann, streamlit and the image datasets are satisfied by stand-ins; the scann
    stand-in implements what the reference asks of it for pools below 2e4 rows -- `builder(db_normalised, k, "dot_product")
    .score_brute_force().build()` = exact top-k by dot product over the rows AS THE REFERENCE NORMALISED THEM (:574).  The constructor
    (datasets, CLIP download) is bypassed; the methods run on an instance carrying the attributes they read."""
    import tempfile
    import types
    import retro_stub

    class ExactScann:
        def __init__(self, db, k, metric):
            assert metric == "dot_product"
            self.db = np.asarray(db, dtype=np.float32)               # ScaNN holds float32 copies of the rows it is given

        def score_brute_force(self):
            return self

        def build(self):
            return self

        def search_batched(self, q, final_num_neighbors=None):
            s = np.asarray(q, dtype=np.float32).astype(np.float64) @ self.db.astype(np.float64).T
            idx = np.argsort(-s, axis=1, kind="stable")[:, :final_num_neighbors]
            return idx.astype(np.uint32), np.take_along_axis(s, idx, 1).astype(np.float32)
    ops = types.ModuleType("scann.scann_ops_pybind"); ops.builder = ExactScann
    sc = types.ModuleType("scann"); sc.scann_ops_pybind = ops
    sys.modules.update({"scann": sc, "scann.scann_ops_pybind": ops, "streamlit": types.ModuleType("streamlit")})
    base = types.ModuleType("rdm.data.base"); base.PatcherDataset = object
    sys.modules["rdm.data.base"] = base
    sys.modules["ldm.util"].parallel_data_prefetch = None
    sys.modules["omegaconf"].OmegaConf = object
    sys.modules["pytorch_lightning"].seed_everything = lambda s: None
    sys.modules.pop("rdm.data.retrieval_dataset.dsetbuilder", None)                # (neighbour_precompute() registered an empty stand-in)
    import importlib
    dsb = importlib.import_module("rdm.data.retrieval_dataset.dsetbuilder")       # the reference's
    assert "/root/reference" in dsb.__file__
    db, _, _ = ref_weights.make_db(N_DB)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "2022-01-01T00-00-00-db.npz")
        np.savez(path, embedding=db, img_id=np.arange(N_DB) * 3, patch_coords=np.stack([np.arange(N_DB)] * 4, 1).astype(np.int32))
        b = object.__new__(dsb.DatasetBuilder)
        b.data_pool = {"embedding": [], "img_id": [], "patch_coords": []}
        b.saved_embeddings, b.max_pool_size, b.k, b.distance_metric = path, 1000, 5, "dot_product"
        b.searcher, b.searcher_savedir, b.save_searcher, b.visualize, b.gpu = None, None, False, False, False
        b.retriever = retro_stub.PatchEmbedStub()
        b.load_embeddings()
        b.train_searcher()
    out = {"n_db": np.int64(N_DB)}
    q = ref_weights.tensor_for("dsb_queries", (3, 512), 81) * 22.0
    q[0] = db[17].astype(np.float32)                                               # a database row as query
    r = b.search_k_nearest(q, k=5, query_embedded=True)
    out["embedded:queries"] = q
    for key in ("embeddings", "img_ids", "patch_coords", "nns", "q_embeddings"):
        out[f"embedded:{key}"] = np.asarray(r[key])
    imgs = ref_weights.tensor_for("dsb_images", (2, 2, 8, 8, 3), 82)               # [b, n, h, w, c] patches in channel-last layout (:464-467)
    r = b.search_k_nearest(torch.from_numpy(imgs), k=4, is_caption=False)
    out["images:queries"] = imgs
    for key in ("embeddings", "img_ids", "patch_coords", "nns", "q_embeddings"):
        out[f"images:{key}"] = np.asarray(r[key])
    save("ref_dsetbuilder.npz", out)


if __name__ == "__main__":
    ref_stubs.install()
    unet_and_ddim()
    rarm()
    pipeline()
    retro_sampler()
    sampler_options()
    neighbour_precompute()
    rarm_sampling()
    dataset_builder()
