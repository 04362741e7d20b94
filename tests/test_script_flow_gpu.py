"""GPU: the sampling script of the reference on the REAL library.

Where the reference checkout exists next to a GPU (`RDM_REFERENCE_DIR`, default /root/reference) the UNCHANGED `scripts/rdm_sample.py` is
loaded and its `parse_args` -> `load_model` -> `sample_unconditional` / `sample_conditional` run with `--gpu 0` against librdm_b200.so.
On the GPU boxes of this build the checkout does not exist (nothing there may read it), so the same test then drives the identical call
sequence -- the statements of `load_model` (rdm_sample.py:141-187) and `sample_unconditional` (:226-252) restated below with their line
numbers -- over a model directory in the reference's layout: `config.yaml` with every key of the shipped `models/rdm/imagenet/config.yaml`
(tests/golden/shipped_config.py, equality with the shipped file is asserted in the CPU tier), a Lightning checkpoint (live + EMA U-Net,
first stage), a two-part `.npz` database and the `nn_memory` pickle.  No stand-ins: exact kNN, U-Net, DDIM and the VQ decoder are the
CUDA executors; the images must equal the oracle pipeline (EMA weights, retrieval conditioning, CFG 2, first-stage decode)."""
import argparse
import importlib.util
import os
import pathlib
import pickle
import sys

import numpy as np
import pytest
import torch
import yaml

from conftest import ROOT
from oracle import ddim as oddim, knn as oknn, unet as ounet, vqdecoder as ovq

GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLD)
import ref_weights  # noqa: E402
from shipped_config import RDM_IMAGENET_MODEL  # noqa: E402

REF = os.environ.get("RDM_REFERENCE_DIR", "/root/reference")
HAVE_REF = os.path.isfile(os.path.join(REF, "scripts", "rdm_sample.py"))


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference checkout (build container only)")
def test_restated_config_equals_the_shipped_file():
    shipped = yaml.safe_load(open(os.path.join(REF, "models", "rdm", "imagenet", "config.yaml")))["model"]
    assert shipped == RDM_IMAGENET_MODEL


def make_model_dir(tmp_path, n_db=400):
    """config.yaml (shipped keys, sizes reduced), model.ckpt (Lightning layout), database/ parts, nn_memory.p"""
    import copy
    db, mem, id_count = ref_weights.make_db(n_db)
    os.makedirs(tmp_path / "database", exist_ok=True)
    np.savez(tmp_path / "database" / "part0.npz", embedding=db[:150], img_id=np.arange(150), patch_coords=np.zeros((150, 4), np.int32))
    np.savez(tmp_path / "database" / "part1.npz", embedding=db[150:], img_id=np.arange(150, n_db), patch_coords=np.zeros((n_db - 150, 4), np.int32))
    with open(tmp_path / "nn_memory.p", "wb") as f:
        pickle.dump({"nn_memory": mem, "id_count": id_count}, f)
    cfg = {"model": copy.deepcopy(RDM_IMAGENET_MODEL)}
    p = cfg["model"]["params"]
    p["image_size"], p["nn_memory"] = 16, str(tmp_path / "nn_memory.p")
    p["unet_config"]["params"].update(dict(image_size=16, model_channels=64, attention_resolutions=[2, 4], num_res_blocks=1, channel_mult=[1, 2, 3]))
    p["first_stage_config"]["params"].update(dict(n_embed=64))
    p["first_stage_config"]["params"]["ddconfig"].update(dict(resolution=32, ch=64, ch_mult=[1, 2], num_res_blocks=1))      # (the CUDA decoder works on 64-channel K blocks)
    p["retrieval_cfg"]["params"]["saved_embeddings"] = str(tmp_path / "database")
    model_dir = tmp_path / "model"
    model_dir.mkdir()
    with open(model_dir / "config.yaml", "w") as f:
        yaml.safe_dump(cfg, f)
    ucfg = dict(p["unet_config"]["params"])
    unet = ounet.randomize_(ounet.UNetModel(**ucfg), 1)
    ema = ounet.randomize_(ounet.UNetModel(**ucfg), 2).eval()
    fs = ovq.randomize_(ovq.VQModelInterface(**p["first_stage_config"]["params"]), 3).eval()
    sd = {"model.diffusion_model." + k: v for k, v in unet.state_dict().items()}
    sd.update({"model_ema." + ("diffusion_model." + k).replace(".", ""): v for k, v in ema.state_dict().items()})
    sd.update({"model_ema.decay": torch.tensor(0.9999), "model_ema.num_updates": torch.tensor(10, dtype=torch.int)})
    sd.update({"first_stage_model." + k: v for k, v in fs.state_dict().items()})
    torch.save({"state_dict": sd, "global_step": 1}, model_dir / "model.ckpt")
    return model_dir, db, ema, fs


def restated_load_model(opt):
    """rdm_sample.py:141-187, statement by statement"""
    import rdm  # noqa: F401
    from ldm.util import instantiate_from_config
    from omegaconf import OmegaConf
    from rdm.models.diffusion.ddpm import MinimalRETRODiffusion
    config_path, ckpt_path = pathlib.Path(opt.model_path) / "config.yaml", pathlib.Path(opt.model_path) / "model.ckpt"      # :146-151
    config = OmegaConf.load(config_path)                                                                                # :155
    config.model.params.retrieval_cfg.params.load_patch_dataset = opt.save_nns                                          # :156
    config.model.params.retrieval_cfg.params.gpu = False                                                                # :159
    config.model.params.retrieval_cfg.params.retriever_config.params.device = "cpu"                                     # :160
    pl_sd = torch.load(ckpt_path, map_location="cpu")                                                                   # :163
    model = instantiate_from_config(config.model)                                                                       # :166
    assert isinstance(model, MinimalRETRODiffusion)                                                                     # :167
    m, u = model.load_state_dict(pl_sd["state_dict"], strict=False)                                                     # :170
    assert len(u) == 0
    model = model.eval()                                                                                                # :178
    if opt.gpu >= 0:                                                                                                    # :180-185
        device = torch.device(f"cuda:{opt.gpu}")
        model = model.to(device)
        model.retriever.retriever.to(device)
    return model


def restated_sample_unconditional(model, opt):
    """rdm_sample.py:226-252 (the PNG writing of :253-262 is torchvision's and stays with the caller)"""
    qids = model.get_qids(opt.top_m, opt.batch_size, use_weights=opt.use_weights) if opt.keep_qids else None            # :227-230
    return model.sample_from_rdata(opt.batch_size, qids=qids, k_nn=opt.k_nn, return_nns=opt.save_nns, use_weights=opt.use_weights,          # :241-252
                                   memsize=opt.top_m, unconditional_guidance_scale=opt.guidance_scale, ddim_steps=opt.steps, ddim=True,
                                   unconditional_retro_guidance_label=0.)


@pytest.mark.gpu
def test_rdm_sample_flow_on_the_device(tmp_path, cuda, monkeypatch):
    model_dir, db, ema, fs = make_model_dir(tmp_path)
    bs, steps = 3, 6
    out = tmp_path / "out"
    monkeypatch.setenv("RDM_B200_MODE", "bf16x3")                    # strict split mode: this tiny random net amplifies operand rounding (see tests/test_mirror_gpu.py)
    argv = ["rdm_sample.py", "-s", str(out), "--model_path", str(model_dir), "-bs", str(bs), "--gpu", str(cuda.index or 0), "--n_runs", "1",
            "--steps", str(steps), "--k_nn", "4", "--guidance_scale", "2.0", "--top_m", "0.5"]
    if HAVE_REF:
        import rdm  # noqa: F401
        spec = importlib.util.spec_from_file_location("ref_script_rdm_sample", os.path.join(REF, "scripts", "rdm_sample.py"))
        script = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(script)
        monkeypatch.setattr(sys, "argv", argv)
        opt = script.parse_args()
        load_model, sample_unconditional = script.load_model, script.sample_unconditional
    else:
        opt = argparse.Namespace(model_path=str(model_dir), savepath=out, batch_size=bs, gpu=int(cuda.index or 0), n_runs=1, steps=steps, k_nn=4, guidance_scale=2.0,
                                 top_m=0.5, save_nns=False, use_weights=False, keep_qids=False, seed=None, increase_guidance=False)
        load_model, sample_unconditional = restated_load_model, restated_sample_unconditional
    out.mkdir(parents=True, exist_ok=True)
    model = load_model(opt)
    assert type(model).__module__ == "rdm.models.diffusion.ddpm" and "retrieval-augmented-diffusion-models_b200" in sys.modules[type(model).__module__].__file__
    assert next(model.model.diffusion_model.parameters()).is_cuda
    # the sampler draws x_T (and one noise tensor per step) from the DEVICE generator; route the draws through the CPU generator so that the
    # oracle below can replay them (same order: x_T, then S per-step draws, ddim.py:156,226-227)
    orig_randn = torch.randn
    monkeypatch.setattr(torch, "randn", lambda *a, device=None, **k: orig_randn(*a, **k).to(device) if device is not None else orig_randn(*a, **k))
    np.random.seed(0); torch.manual_seed(0)
    logs = sample_unconditional(model, opt)
    if HAVE_REF:                                                       # the script writes PNG files and returns nothing: repeat the call for the tensors
        assert len(os.listdir(out)) == bs
        np.random.seed(0); torch.manual_seed(0)
        logs = restated_sample_unconditional(model, argparse.Namespace(batch_size=bs, k_nn=4, save_nns=False, use_weights=False, top_m=0.5, guidance_scale=2.0,
                                                                       steps=steps, keep_qids=False))
    assert list(logs.keys()) == ["samples_with_sampled_nns"]
    imgs = logs["samples_with_sampled_nns"]
    assert imgs.is_cuda and tuple(imgs.shape) == (bs, 3, 32, 32)
    # the oracle pipeline on the same RNG: pseudo-queries (NumPy global RNG), exact neighbours, x_T (torch global RNG), EMA U-Net, CFG 2, VQ decode
    nns = logs.extras["nns"].cpu().numpy()
    qh = oknn.normalize_queries(db[nns[:, 0]].astype(np.float32))
    assert np.array_equal(oknn.search(db, qh, 4)[0], nns)              # kNN indices bit-exact (the query row is its own nearest neighbour)
    np.random.seed(0); torch.manual_seed(0)
    qids = model.get_qids(0.5, bs, use_weights=False)
    assert np.array_equal(np.asarray(qids), nns[:, 0])
    x_T = orig_randn(bs, 3, 16, 16)                                    # DDIMSampler.ddim_sampling draws x_T first (ddim.py:156)
    cond = torch.from_numpy(db[nns].astype(np.float32))
    want_z = oddim.ddim_sample(ema, x_T, cond, torch.zeros_like(cond), S=steps, scale=2.0, draw_noise_always=True)
    with torch.no_grad():
        want = fs.decode(want_z)
    err = float((imgs.double().cpu() - want.double()).norm() / want.double().norm())
    print(f"rdm_sample flow on the device ({'unchanged script' if HAVE_REF else 'restated call sequence'}): image rel-L2 {err:.2e}")
    assert err < 2e-3
