"""GPU: the reference-facing API end to end -- YAML-style config -> MinimalRETRODiffusion -> sample_from_rdata /
sample_with_query -> DDIMSampler -> librdm_b200, checked against the oracle pipeline on identical inputs and RNG."""
import numpy as np
import pytest
import torch

from oracle import ddim as oddim
from oracle import knn as oknn
from oracle import unet as ounet
from test_mirror_host import TINY_CFG

pytestmark = pytest.mark.gpu


def _setup(tmp_path, cuda, seed=4):
    import copy
    import rdm  # noqa: F401
    from ldm.util import instantiate_from_config
    from omegaconf import OmegaConf
    rng = np.random.default_rng(seed)
    db = (rng.standard_normal((30_000, 512)) * rng.uniform(0.5, 8, (30_000, 1))).astype(np.float16)
    np.savez(tmp_path / "db.npz", embedding=db, img_id=np.arange(30_000), patch_coords=np.zeros((30_000, 4), np.int32))
    cfg = copy.deepcopy(TINY_CFG)
    cfg["params"]["retrieval_cfg"]["params"]["saved_embeddings"] = str(tmp_path / "db.npz")
    model = instantiate_from_config(OmegaConf.create(cfg))
    ref = ounet.randomize_(ounet.UNetModel(**ounet.TINY_UNET), seed).eval()
    ema = ounet.randomize_(ounet.UNetModel(**ounet.TINY_UNET), seed + 1).eval()          # sampling must use THESE (ddpm.py:977)
    ck = {"model.diffusion_model." + k: v for k, v in ref.state_dict().items()}
    ck.update({"model_ema." + ("diffusion_model." + k).replace(".", ""): v for k, v in ema.state_dict().items()})
    missing, unexpected = model.load_state_dict(ck, strict=False)
    sched = {"betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod",
             "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod"}            # present in real checkpoints, not in this synthetic one
    assert not unexpected and set(missing) <= {"unconditional_guidance_vex", "model_ema.decay", "model_ema.num_updates"} | sched
    model = model.eval().to(cuda)
    # strict split mode for the oracle comparison: this tiny random net has activations ~100 and amplifies operand rounding; the
    # default fp16x2 mode is validated at full architecture size (tests/test_unet_gpu.py::test_full_arch_ddim20_tensor_core, tools/ddim_error.py)
    model.model.diffusion_model.engine_mode = "bf16x3"
    return model, db, ema


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


def test_sample_from_rdata_matches_the_oracle_pipeline(tmp_path, cuda):
    model, db, ema = _setup(tmp_path, cuda)
    qids = np.array([5, 777, 12345])
    x_T = torch.randn(3, 4, 16, 16, generator=torch.Generator().manual_seed(0))
    logs = model.sample_from_rdata(3, qids=qids, k_nn=4, use_weights=False, memsize=100, unconditional_guidance_scale=2.0, ddim_steps=10,
                                   ddim=True, unconditional_retro_guidance_label=0., x_T=x_T.to(cuda))             # scripts/rdm_sample.py:241-252
    # oracle: exact kNN on normalised DB-row queries, raw neighbour rows as context, zeros uncond, EMA weights
    qh = oknn.normalize_queries(db[qids].astype(np.float32))
    nns, _ = oknn.search(db, qh, 4)
    assert np.array_equal(logs["nns"].cpu().numpy(), nns) and list(nns[:, 0]) == list(qids)
    cond = torch.from_numpy(db[nns].astype(np.float32))
    want = oddim.ddim_sample(ema, x_T, cond, torch.zeros_like(cond), S=10, scale=2.0)
    assert logs["samples_with_sampled_nns"].shape == (3, 4, 16, 16)
    assert rel_l2(logs["samples_with_sampled_nns"], want) < 1e-3


def test_sample_with_query_embedded_prepends_the_query(tmp_path, cuda):
    model, db, ema = _setup(tmp_path, cuda, seed=6)
    g = torch.Generator().manual_seed(1)
    q = torch.randn(2, 512, generator=g) * 4
    x_T = torch.randn(2, 4, 16, 16, generator=g)
    logs = model.sample_with_query(query=q, query_embedded=True, k_nn=4, visualize_nns=False, use_weights=False, unconditional_guidance_scale=3.0,
                                   ddim_steps=5, ddim=True, unconditional_retro_guidance_label=0., omit_query=False, x_T=x_T.to(cuda))   # rdm_sample.py:287-299
    qh = oknn.normalize_queries(q.numpy())
    nns, _ = oknn.search(db, qh, 4)
    cond = torch.cat([q[:, None], torch.from_numpy(db[nns].astype(np.float32))[:, :3]], dim=1)                         # ddpm.py:775
    want = oddim.ddim_sample(ema, x_T, cond, torch.zeros_like(cond), S=5, scale=3.0)
    assert rel_l2(logs["query_samples"], want) < 1e-3


def test_ddim_sampler_generic_path_with_callbacks_and_intermediates(tmp_path, cuda):
    from rdm.models.diffusion.ddim import DDIMSampler
    model, db, ema = _setup(tmp_path, cuda, seed=8)
    g = torch.Generator().manual_seed(2)
    cond = torch.randn(2, 4, 512, generator=g)
    x_T = torch.randn(2, 4, 16, 16, generator=g)
    seen = []
    with model.ema_scope():
        s1, inter1 = DDIMSampler(model).sample(S=10, batch_size=2, shape=(4, 16, 16), conditioning=cond.to(cuda), x_T=x_T.to(cuda), verbose=False,
                                               unconditional_guidance_scale=2.0, unconditional_conditioning=torch.zeros_like(cond).to(cuda), log_every_t=5)
        s2, inter2 = DDIMSampler(model).sample(S=10, batch_size=2, shape=(4, 16, 16), conditioning=cond.to(cuda), x_T=x_T.to(cuda), verbose=False,
                                               unconditional_guidance_scale=2.0, unconditional_conditioning=torch.zeros_like(cond).to(cuda), log_every_t=5,
                                               callback=lambda i: seen.append(i))
    want, traj = oddim.ddim_sample(ema, x_T, cond, torch.zeros_like(cond), S=10, scale=2.0, return_all=True)
    assert seen == list(range(10))
    assert rel_l2(s1, want) < 1e-3 and rel_l2(s2, want) < 1e-3
    # intermediates logged at index % 5 == 0 or index == 9  ->  loop positions i = 0, 4, 9 (ddim.py:207-213), after the initial x_T entry
    assert len(inter1["x_inter"]) == 4 and len(inter2["x_inter"]) == 4
    for k, i in enumerate((0, 4, 9)):
        assert rel_l2(inter1["x_inter"][k + 1], traj[i][0]) < 1e-3 and rel_l2(inter1["pred_x0"][k + 1], traj[i][1]) < 1e-3
        assert rel_l2(inter2["x_inter"][k + 1], traj[i][0]) < 1e-3


def test_default_engine_mode_runs_end_to_end(tmp_path, cuda):
    model, db, ema = _setup(tmp_path, cuda, seed=9)
    model.model.diffusion_model.engine_mode = "fp16x2"
    x_T = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(3))
    logs = model.sample_from_rdata(2, qids=np.array([1, 2]), k_nn=4, unconditional_guidance_scale=2.0, ddim_steps=5, ddim=True,
                                   unconditional_retro_guidance_label=0., x_T=x_T.to(cuda))
    cond = torch.from_numpy(db[logs["nns"].cpu().numpy()].astype(np.float32))
    want = oddim.ddim_sample(ema, x_T, cond, torch.zeros_like(cond), S=5, scale=2.0)
    assert torch.isfinite(logs["latents"]).all() and rel_l2(logs["latents"], want) < 2e-2


def test_first_stage_decode_runs_on_the_device(tmp_path, cuda):
    """`decode_first_stage` (ddpm.py:840,981): with a first_stage_config the sampler returns IMAGES decoded by the CUDA VQ decoder
    (librdm_b200 rdm_vqdec_decode behind ldm.models.autoencoder.VQModelInterface); checked against the oracle decoder on the returned latents."""
    import copy
    import rdm  # noqa: F401
    from ldm.util import instantiate_from_config
    from omegaconf import OmegaConf
    from oracle import vqdecoder as ovq
    rng = np.random.default_rng(3)
    db = rng.standard_normal((20_000, 512)).astype(np.float16)
    np.savez(tmp_path / "db.npz", embedding=db, img_id=np.arange(20_000), patch_coords=np.zeros((20_000, 4), np.int32))
    vq = dict(embed_dim=4, n_embed=256, ddconfig=dict(ovq.TINY_VQ["ddconfig"], z_channels=4, resolution=32))
    cfg = copy.deepcopy(TINY_CFG)
    cfg["params"]["retrieval_cfg"]["params"]["saved_embeddings"] = str(tmp_path / "db.npz")
    cfg["params"]["first_stage_config"] = {"target": "ldm.models.autoencoder.VQModelInterface", "params": dict(vq, lossconfig={"target": "torch.nn.Identity"})}
    model = instantiate_from_config(OmegaConf.create(cfg))
    unet = ounet.randomize_(ounet.UNetModel(**ounet.TINY_UNET), 5).eval()
    fs = ovq.randomize_(ovq.VQModelInterface(**vq), 6).eval()
    ck = {"model.diffusion_model." + k: v for k, v in unet.state_dict().items()}
    ck.update({"model_ema." + ("diffusion_model." + k).replace(".", ""): v for k, v in unet.state_dict().items()})
    ck.update({"first_stage_model." + k: v for k, v in fs.state_dict().items()})
    missing, unexpected = model.load_state_dict(ck, strict=False)
    assert not unexpected and not [m for m in missing if m.startswith("first_stage_model.")]
    model = model.eval().to(cuda)
    x_T = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(1))
    logs = model.sample_from_rdata(2, qids=np.array([10, 20]), k_nn=4, unconditional_guidance_scale=2.0, ddim_steps=5, ddim=True,
                                   unconditional_retro_guidance_label=0., x_T=x_T.to(cuda))
    imgs = logs["samples_with_sampled_nns"]
    assert imgs.is_cuda and imgs.shape == (2, 3, 32, 32)
    with torch.no_grad():
        want = fs.decode(logs["latents"].float().cpu() / float(model.scale_factor))
    assert rel_l2(imgs, want) < 5e-3


def _retro_model(tmp_path, cuda):
    """Tiny RDM with a first stage and a (random-weight, one-layer, full-width) CLIP image retriever: everything cfg4's re-retrieval needs."""
    import copy
    import rdm  # noqa: F401
    from ldm.util import instantiate_from_config
    from omegaconf import OmegaConf
    from oracle import clip as oclip
    from oracle import vqdecoder as ovq
    from rdm_b200.clip import VIT_B32
    rng = np.random.default_rng(11)
    db = rng.standard_normal((20_000, 512)).astype(np.float16)
    np.savez(tmp_path / "db.npz", embedding=db, img_id=np.arange(20_000), patch_coords=np.zeros((20_000, 4), np.int32))
    vq = dict(embed_dim=4, n_embed=256, ddconfig=dict(ovq.TINY_VQ["ddconfig"], z_channels=4, resolution=32))
    cfg = copy.deepcopy(TINY_CFG)
    cfg["params"]["retrieval_cfg"]["params"]["saved_embeddings"] = str(tmp_path / "db.npz")
    cfg["params"]["first_stage_config"] = {"target": "ldm.models.autoencoder.VQModelInterface", "params": dict(vq, lossconfig={"target": "torch.nn.Identity"})}
    model = instantiate_from_config(OmegaConf.create(cfg))
    unet = ounet.randomize_(ounet.UNetModel(**ounet.TINY_UNET), 5).eval()
    fs = ovq.randomize_(ovq.VQModelInterface(**vq), 6).eval()
    ck = {"model.diffusion_model." + k: v for k, v in unet.state_dict().items()}
    ck.update({"model_ema." + ("diffusion_model." + k).replace(".", ""): v for k, v in unet.state_dict().items()})
    ck.update({"first_stage_model." + k: v for k, v in fs.state_dict().items()})
    model.load_state_dict(ck, strict=False)
    model = model.eval().to(cuda)
    model.model.diffusion_model.engine_mode = "bf16x3"
    clip_sd = oclip.random_state_dict(**dict(VIT_B32, vocab_size=64, vision_layers=1, transformer_layers=1), seed=7)
    model.retriever.retriever.model.load_state_dict(clip_sd)
    model.retriever.retriever.to(cuda)                                                          # scripts/rdm_sample.py:185
    return model, db, unet


def _oracle_nns(model, images, db, k):
    emb = model.retriever.retriever(images.float()).float().cpu().numpy()
    qh = emb / np.linalg.norm(emb, axis=1)[:, np.newaxis]                                       # ddpm.py:297 (numpy, as the reference)
    return oknn.search(db, qh.astype(np.float32), k)[0]


def test_get_nn_and_encoding_on_device(tmp_path, cuda):
    model, db, _ = _retro_model(tmp_path, cuda)
    imgs = (torch.rand(3, 3, 32, 32, generator=torch.Generator().manual_seed(2)) * 2 - 1).to(cuda)
    out = model.get_nn_and_encoding(imgs, k_nn=4)
    enc = out[model.nn_key]
    assert enc.shape == (3, 1, 4, 512) and enc.is_cuda
    nns = _oracle_nns(model, imgs, db, 4)
    assert np.array_equal(out["nns"].cpu().numpy(), nns)
    assert np.array_equal(enc.cpu().numpy()[:, 0], db[nns].astype(np.float32))
    four = model.get_nn_and_encoding(imgs, k_nn=2, n_patches_per_side=2)                        # 2 x 2 patches of 16 x 16
    assert four[model.nn_key].shape == (3, 4, 2, 512)
    patches = torch.stack([imgs[..., i * 16:(i + 1) * 16, j * 16:(j + 1) * 16] for i in range(2) for j in range(2)], dim=1).reshape(12, 3, 16, 16)
    assert np.array_equal(four["nns"].cpu().numpy(), _oracle_nns(model, patches, db, 2))


def test_per_step_re_retrieval_sampler(tmp_path, cuda):
    """DDIMRetroSampler (ddim.py:270-415, BASELINE cfg4): step i is conditioned on the neighbours retrieved from the decoded x0 prediction of
    step i-1.  Checked (a) step by step: the recorded neighbours are the exact kNN of the CLIP embedding of decode(pred_x0); (b) end to end:
    the oracle U-Net replaying the trajectory with the recorded contexts reproduces the final latent."""
    from rdm.models.diffusion.ddim import DDIMRetroSampler
    model, db, unet = _retro_model(tmp_path, cuda)
    S, k, scale = 5, 4, 2.0
    x_T = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(3))
    torch.manual_seed(77)
    r0 = torch.randn_like(torch.randn((2, k, 512), device=cuda))                                # the sampler's two draws (ddim.py:297,316): the second is the first conditioning
    uc = torch.zeros(2, k, 512, device=cuda)
    torch.manual_seed(77)
    with model.ema_scope():
        img, inter = DDIMRetroSampler(model).sample(S, 2, (4, 16, 16), r_shape=(2, k, 512), x_T=x_T.to(cuda), log_every_t=1, k_nn=k,
                                                    unconditional_guidance_scale=scale, unconditional_conditioning=[uc], ignore_noising=True, verbose=False)
    assert len(inter["nns"]) == S and len(inter["pred_x0"]) == S
    sch = oddim.Schedule(S)
    x, ctx = x_T, r0.cpu()
    for i in range(S):
        step = int(np.flip(sch.timesteps)[i])
        with torch.no_grad():
            e = unet(torch.cat([x] * 2), torch.full((4,), step), torch.cat([ctx, torch.zeros_like(ctx)]))
        x, p0 = oddim.ddim_update(x, e[2:] + scale * (e[:2] - e[2:]), *sch.coeffs(S - i - 1))
        assert rel_l2(inter["pred_x0"][i], p0) < 2e-3, f"step {i}"
        px0 = model.decode_first_stage(inter["pred_x0"][i])                                      # deterministic device decode of the device's own prediction
        nns = _oracle_nns(model, px0, db, k)
        assert np.array_equal(inter["nns"][i].cpu().numpy(), nns), f"step {i}"
        ctx = torch.from_numpy(db[nns].astype(np.float32))                                       # conditioning of the NEXT step
    assert rel_l2(img, x) < 2e-3


def test_offline_neighbour_precompute_writes_the_reference_file_format(tmp_path, cuda):
    """scripts/search_neighbors.py:381-450 over the device searcher: per-example pickles in the layout QueryDataset.load_nns reads
    (rdm/data/base.py:925-939), and the neighbour histogram behind nn_memory."""
    import pickle
    from rdm_b200.nn_precompute import build_nn_memory, search_nns
    model, db, _ = _retro_model(tmp_path, cuda)
    builder = model.retriever
    builder.train_searcher()

    class Loader(list):
        batch_size = 3
    g = torch.Generator().manual_seed(4)
    batches = Loader({'patches': torch.rand(3, 4, 16, 16, 3, generator=g) * 2 - 1} for _ in range(2))      # 2 x 2 patch grid per example, channel-last
    (tmp_path / "embeddings").mkdir()
    paths = search_nns(builder, batches, device=cuda, save=True, npatches_perside=2, base_savedir=str(tmp_path), start_id=100)
    assert sorted(paths) == list(range(100, 106)) and paths[104] == f"embeddings/{builder.k}_nns-img000000104.p"
    with open(tmp_path / paths[104], "rb") as f:
        entry = pickle.load(f)[2]
    assert entry['nn_ids'].shape == (4, builder.k) and entry['embeddings'].shape == (4, builder.k, 512) and entry['img_ids'].shape == (4, builder.k)
    q = batches[1]['patches'][1].permute(0, 3, 1, 2).to(cuda)                                          # example 104 = batch 1, item 1
    assert np.array_equal(entry['nn_ids'], _oracle_nns(model, q, db, builder.k))
    assert np.array_equal(entry['embeddings'], db[entry['nn_ids']])
    hist = search_nns(builder, batches, device=cuda, save=False)
    assert sum(hist.values()) == 2 * 3 * 4 * builder.k
    mem = build_nn_memory(hist)
    assert mem['nn_memory'].shape[0] == len(hist) and hist[int(mem['nn_memory'][0])] == max(hist.values())
