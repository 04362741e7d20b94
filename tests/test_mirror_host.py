"""CPU: host-side logic of the `rdm.*` mirror (no CUDA calls): construction from the shipped config layout, state-dict
key compatibility with the oracle / checkpoint layout, schedule buffers, qid sampling, DB loading."""
import numpy as np
import pytest
import torch

from oracle import ddim as oddim
from oracle import unet as ounet

TINY_CFG = {
    "target": "rdm.models.diffusion.ddpm.MinimalRETRODiffusion",
    "params": {
        "k_nn": 4, "query_key": "clip_img_emb", "linear_start": 0.0015, "linear_end": 0.0195, "num_timesteps_cond": 1, "log_every_t": 200,
        "timesteps": 1000, "first_stage_key": "image", "cond_stage_key": "nixda", "image_size": 16, "channels": 4, "cond_stage_trainable": False,
        "nn_key": "nn_embeddings", "conditioning_key": "retro_only", "monitor": "val/loss_simple_ema", "scale_by_std": False,
        "ignore_keys": ["unconditional_guidance_vex"],
        "unet_config": {"target": "rdm.modules.diffusionmodules.openaimodel.UNetModel",
                        "params": dict(ounet.TINY_UNET, use_spatial_transformer=True, use_checkpoint=True, use_scale_shift_norm=False, resblock_updown=False)},
        "first_stage_config": None,
        "retrieval_cfg": {"target": "rdm.data.retrieval_dataset.dsetbuilder.DatasetBuilder",
                          "params": {"patch_size": 256, "batch_size": 100, "k": 20, "max_pool_size": 1000, "gpu": False, "load_patch_dataset": False,
                                     "retriever_config": {"target": "rdm.modules.retrievers.ClipImageRetriever", "params": {"model": "ViT-B/32", "device": "cpu"}},
                                     "data": {"target": "rdm.data.openimages.FullOpenImagesTrain", "params": {}}}},
        "retrieval_encoder_cfg": {"target": "torch.nn.Identity"},
        "cond_stage_config": "__is_unconditional__",
    },
}


def _model():
    import rdm  # noqa: F401  (installs the shims)
    from ldm.util import instantiate_from_config
    from omegaconf import OmegaConf
    return instantiate_from_config(OmegaConf.create(TINY_CFG))


def test_instantiates_from_the_reference_config_layout_and_exposes_the_checkpoint_keys():
    m = _model()
    sd = m.state_dict()
    ref = ounet.UNetModel(**ounet.TINY_UNET).state_dict()
    for k, v in ref.items():
        assert sd["model.diffusion_model." + k].shape == v.shape, k
    # LitEma naming: parameter name without dots + decay/num_updates
    assert "model_ema.diffusion_modeltime_embed0weight" in sd and "model_ema.decay" in sd and "model_ema.num_updates" in sd
    n_unet = len(ref)
    assert len([k for k in sd if k.startswith("model_ema.")]) == n_unet + 2          # 'Keeping EMAs of 690' pattern (demo_rdm.ipynb:129)
    assert "unconditional_guidance_vex" in sd and sd["unconditional_guidance_vex"].shape == (512,)
    assert np.array_equal(sd["alphas_cumprod"].numpy(), oddim.alphas_cumprod_f32())


def test_load_state_dict_non_strict_like_the_sampling_script():
    m = _model()
    ref = ounet.randomize_(ounet.UNetModel(**ounet.TINY_UNET), 4)
    ck = {"model.diffusion_model." + k: v for k, v in ref.state_dict().items()}
    ck.update({"model_ema." + ("diffusion_model." + k).replace(".", ""): v * 0.5 for k, v in ref.state_dict().items()})
    ck["first_stage_model.decoder.conv_in.weight"] = torch.zeros(1)
    missing, unexpected = m.load_state_dict(ck, strict=False)                       # scripts/rdm_sample.py:170
    assert "unconditional_guidance_vex" in missing and "first_stage_model.decoder.conv_in.weight" in unexpected
    ema = m.model_ema.state_dict_for("diffusion_model.")
    assert torch.equal(ema["out.2.weight"], ref.state_dict()["out.2.weight"] * 0.5)
    assert torch.equal(m.model.diffusion_model.state_dict()["out.2.weight"], ref.state_dict()["out.2.weight"])


def test_unconditional_conditioning_label_zero_is_zeros():
    m = _model()
    uc = m.get_unconditional_conditioning((3, 4, 512), unconditional_guidance_label=0., k_nn=4)     # ddpm.py:673-680
    assert uc.shape == (3, 4, 512) and float(uc.abs().max()) == 0.0


def test_get_qids_follows_numpy_global_rng(tmp_path):
    m = _model()
    m.retriever.data_pool = {"embedding": np.zeros((1000, 512), np.float16)}
    np.random.seed(5); want = np.random.choice(1000, size=7)
    np.random.seed(5); got = m.get_qids(100, 7)                                         # ddpm.py:867
    assert np.array_equal(got, want)


def test_dataset_builder_loads_single_and_multi_part_npz(tmp_path):
    import rdm  # noqa: F401
    from rdm.data.retrieval_dataset.dsetbuilder import DatasetBuilder
    rng = np.random.default_rng(0)
    parts = [rng.standard_normal((n, 512)).astype(np.float16) for n in (30, 50)]
    d = tmp_path / "db"; d.mkdir()
    for i, p in enumerate(parts):
        np.savez_compressed(d / f"part{i}.npz", embedding=p, img_id=np.arange(len(p)), patch_coords=np.zeros((len(p), 4)))
    b = DatasetBuilder({"target": "rdm.modules.retrievers.ClipImageRetriever", "params": {"model": "ViT-B/32"}}, saved_embeddings=str(d),
                       load_patch_dataset=False, gpu=False, max_pool_size=10)
    assert b.data_pool["embedding"].shape == (80, 512) and b.data_pool["embedding"].dtype == np.float16
    assert np.array_equal(b.data_pool["embedding"][:30], parts[0]) and b.searcher is None
    b1 = DatasetBuilder(None, saved_embeddings=str(d / "part1.npz"), load_patch_dataset=False, gpu=False)
    assert b1.data_pool["embedding"].shape == (50, 512) and b1.max_pool_size == 50


def test_out_of_scope_surfaces_fail_loudly():
    m = _model()
    with pytest.raises(NotImplementedError):
        m.shared_step({})
    with pytest.raises(NotImplementedError):
        m.retriever.build_data_pool()
    with pytest.raises(RuntimeError):
        m.model.diffusion_model(torch.zeros(1, 4, 16, 16), torch.zeros(1, dtype=torch.long), torch.zeros(1, 4, 512))    # no CPU path


def test_ddim_sampler_schedule_buffers_match_the_oracle():
    import rdm  # noqa: F401
    from rdm.models.diffusion.ddim import DDIMSampler

    class Stub:
        num_timesteps = 1000
        device = torch.device("cpu")
        alphas_cumprod = torch.from_numpy(oddim.alphas_cumprod_f32())
        betas = torch.from_numpy(oddim.make_beta_schedule().astype(np.float32))
        alphas_cumprod_prev = torch.cat([torch.ones(1), alphas_cumprod[:-1]])
    s = DDIMSampler(Stub())
    s.make_schedule(100, ddim_eta=0.0, verbose=False)
    ref = oddim.Schedule(100)
    assert np.array_equal(s.ddim_timesteps, ref.timesteps) and np.array_equal(np.asarray(s.ddim_alphas), ref.alphas)
    assert np.array_equal(np.asarray(s.ddim_alphas_prev), ref.alphas_prev) and np.array_equal(np.asarray(s.ddim_sqrt_one_minus_alphas), ref.sqrt_one_minus_alphas)


def test_ddim_tables_for_step_counts_that_do_not_divide_the_schedule():
    """`make_ddim_timesteps` is `range(0, T, T // S)` (ldm util): 6 requested steps give 7, 30 give 31 -- the reference runs them all
    (ddim.py:164, `total_steps = timesteps.shape[0]`).  Found by tests/test_script_flow_gpu.py (`--steps 6`)."""
    from oracle import ddim as oddim
    from rdm_b200 import sampler
    for S, n in ((6, 7), (30, 31), (100, 100), (250, 250)):
        tb = sampler.make_ddim_tables(sampler.alphas_cumprod_linear(), S, 0.0)
        sch = oddim.Schedule(S, 0.0)
        assert tb["timesteps"].shape == (n,) and tb["coef"].shape == (n, 8)
        assert np.array_equal(tb["timesteps"].numpy(), np.flip(sch.timesteps).astype(np.int64))
