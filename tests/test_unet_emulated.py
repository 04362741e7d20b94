"""CPU: the strict fp32 mode of the U-Net executor -- csrc/unet.cu (plan, weight re-packing, arena, CUDA-graph capture of the forward and
of the DDIM step), csrc/kernels.cu (GroupNorm, LayerNorm, attention, layout, fused CFG + DDIM update) and csrc/gemm_simt.cu (implicit-GEMM
convolutions), as written -- compiled against the host emulation of CUDA in tests/emu/ and driven through the same C ABI and Python
wrapper as on the GPU, against the reference-pinned oracle.  The tensor-core modes need hardware (tests/test_unet_gpu.py); this checks,
without a GPU, the part of the product that defines its results in strict mode, and puts guard zones around every buffer it writes."""
import contextlib
import ctypes
import os
import shutil
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import ddim as oddim, unet as ounet

sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import ref_weights  # noqa: E402

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ to build the emulated library")


@pytest.fixture()
def emulated(monkeypatch):
    import build_emu
    from rdm_b200 import _lib
    os.environ["RDM_KNN_NO_TC"] = "1"                      # the tensor-core kNN scan needs hardware; every batch goes through the SIMT kernels
    names = [n for n in _lib.SIGNATURES if n.startswith(("rdm_unet_", "rdm_ddim_", "rdm_knn_"))] + ["rdm_last_error", "rdm_launch_count", "rdm_abi_version"]
    L = _lib.bind(ctypes.CDLL(build_emu.build()), names)
    monkeypatch.setattr(_lib, "_lib", L)
    monkeypatch.setattr(_lib, "resolve_device", lambda d: torch.device("cpu"))
    monkeypatch.setattr(_lib, "device_ctx", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(_lib, "stream_ptr", lambda d=None: None)
    return L


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def _pair(seed):
    from rdm_b200.unet import B200UNet
    ref = ounet.randomize_(ounet.UNetModel(**ounet.TINY_UNET), seed).eval()
    net = B200UNet("cpu", **ounet.TINY_UNET)
    net.load_state_dict(ref.state_dict())
    assert net.missing() == 0
    net.set_mode(0)
    return ref, net


def test_strict_forward_matches_the_pinned_oracle(emulated):
    ref, net = _pair(1)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 4, 8, 8, generator=g)
    t = torch.tensor([501, 12])
    c = torch.randn(2, 3, 512, generator=g) * 3
    with torch.no_grad():
        want = ref(x, t, c)
    net.set_context(c)
    got = net.forward(x, t)                                   # eager pass (warm-up) + graph capture
    assert rel(got, want) < 1e-5
    assert rel(net.forward(x, t), want) < 1e-5                # graph replay
    # classifier-free-guidance doubling without materialising the batch: x [1] against a [2]-row context
    net.set_context(torch.cat([c[:1], torch.zeros_like(c[:1])]))
    with torch.no_grad():
        want2 = ref(torch.cat([x[:1]] * 2), t[:1].repeat(2), torch.cat([c[:1], torch.zeros_like(c[:1])]))
    assert rel(net.forward(x[:1], t[:1].repeat(2)), want2) < 1e-5


def test_fused_guided_ddim_loop_matches_the_pinned_oracle(emulated):
    from rdm_b200 import sampler
    from rdm_b200.unet import ddim_step
    ref, net = _pair(2)
    g = torch.Generator().manual_seed(4)
    x_T = torch.randn(1, 4, 8, 8, generator=g)
    c, uc = torch.randn(1, 2, 512, generator=g), torch.zeros(1, 2, 512)
    S = 2
    want, traj = oddim.ddim_sample(ref, x_T, c, uc, S=S, scale=2.0, return_all=True)
    tb = sampler.make_ddim_tables(sampler.alphas_cumprod_linear(), S, 0.0)
    net.set_context(torch.cat([c, uc]))
    got, p0 = net.ddim_sample(x_T, tb["timesteps"], tb["coef"], cfg_scale=2.0, want_pred_x0=True)      # step 1 eager, step 2 from the captured graph
    assert rel(got, want) < 1e-5 and rel(p0, traj[-1][1]) < 1e-5
    a = net.ddim_sample(x_T, tb["timesteps"], tb["coef"], cfg_scale=2.0, first_step=0, num_steps=1)
    b = net.ddim_sample(a, tb["timesteps"], tb["coef"], cfg_scale=2.0, first_step=1, num_steps=1)
    assert rel(b, want) < 1e-5
    # the stand-alone update kernel is bit-exact against the reference's float32 tensor expressions
    eps = torch.randn(2, 4, 8, 8, generator=g)
    xp, pp = ddim_step(x_T, eps, tb["coef"][1], cfg_scale=2.0)
    e = eps[1:] + 2.0 * (eps[:1] - eps[1:])
    sch = oddim.Schedule(S)
    wx, wp = oddim.ddim_update(x_T, e, *sch.coeffs(S - 2))
    assert torch.equal(xp, wx) and torch.equal(pp, wp)


def test_strict_forward_matches_reference_code(emulated):
    """The product's CUDA source (strict mode, emulated) against the output of the REFERENCE's own UNetModel.forward
    (tests/golden/ref_unet_tiny.npz, small case): no oracle in between."""
    import ast
    from rdm_b200.unet import B200UNet
    d = np.load(os.path.join(ROOT, "tests", "golden", "ref_unet_tiny.npz"))
    cfg = ast.literal_eval(str(d["cfg_json"]))
    cfg.pop("use_spatial_transformer")
    net = B200UNet("cpu", **cfg)
    assert list(net.shapes) == [str(k) for k in d["sd_keys"]]
    net.load_state_dict(ref_weights.state_dict_for(net.shapes.items(), int(d["weight_seed"])))
    net.set_mode(0)
    net.set_context(torch.from_numpy(d["small:context"]))
    got = net.forward(torch.from_numpy(d["small:x"]), torch.from_numpy(d["small:t"]))
    assert rel(got, torch.from_numpy(d["small:out"])) < 1e-5


def test_product_pipeline_on_the_emulated_engine_matches_reference_code(emulated, monkeypatch, tmp_path):
    """`MinimalRETRODiffusion.sample_from_rdata` of this repository end to end -- host orchestration, EMA weights handed to the engine,
    DDIMSampler's fused loop, and the product's U-Net / DDIM kernels (strict mode, emulated) -- against the latents the REFERENCE's own
    MinimalRETRODiffusion produced for the same checkpoint, database, query id and x_T (tests/golden/ref_pipeline_tiny.npz, small case:
    one image, two guided steps, an 8 x 8 latent via `custom_shape`).  Every device stage is the product's own source under emulation: the exact
    kNN scan and gather (csrc/knn.cu), the cross-attention K/V projection, the U-Net and the fused guided DDIM step (csrc/unet.cu, kernels.cu,
    gemm_simt.cu); no oracle anywhere on the path."""
    import copy
    import rdm  # noqa: F401
    from ldm.util import instantiate_from_config
    from omegaconf import OmegaConf
    from rdm_b200.unet import unet_param_shapes
    from test_mirror_host import TINY_CFG
    p = np.load(os.path.join(ROOT, "tests", "golden", "ref_pipeline_tiny.npz"))
    db, _, _ = ref_weights.make_db(int(p["n_db"]))
    np.savez(tmp_path / "db.npz", embedding=db, img_id=np.arange(len(db)), patch_coords=np.zeros((len(db), 4), np.int32))
    cfg = copy.deepcopy(TINY_CFG)
    cfg["params"]["retrieval_cfg"]["params"]["saved_embeddings"] = str(tmp_path / "db.npz")
    model = instantiate_from_config(OmegaConf.create(cfg))
    shapes = unet_param_shapes(**ounet.TINY_UNET)
    live, ema = (ref_weights.state_dict_for(shapes.items(), int(p[s])) for s in ("live_seed", "ema_seed"))
    ck = {"model.diffusion_model." + k: v for k, v in live.items()}
    ck.update({"model_ema." + ("diffusion_model." + k).replace(".", ""): v for k, v in ema.items()})
    model.load_state_dict(ck, strict=False)
    model = model.eval()
    model.model.diffusion_model.engine_mode = "fp32"
    logs = model.sample_from_rdata(1, qids=np.array([123]), k_nn=4, x_T=torch.from_numpy(p["rdata_small:x_T"]), custom_shape=(4, 8, 8),
                                   unconditional_guidance_scale=2.0, ddim_steps=2, ddim=True, unconditional_retro_guidance_label=0.)
    assert rel(logs["samples_with_sampled_nns"], torch.from_numpy(p["rdata_small:samples"])) < 1e-5


def test_batch_growth_with_smaller_images_does_not_overrun_the_timestep_buffer(emulated):
    """Regression (found by the emulator's guard zones): the staging buffers were sized by B2*C*H*W only, so a larger batch of smaller
    images reused a timestep vector allocated for fewer samples and wrote past it.  Shapes in that order, results checked as well."""
    ref, net = _pair(7)
    g = torch.Generator().manual_seed(0)
    for B2, H, W, k in [(3, 12, 12, 1), (7, 8, 4, 2)]:
        x, t, c = torch.randn(B2, 4, H, W, generator=g), torch.randint(0, 1000, (B2,), generator=g), torch.randn(B2, k, 512, generator=g) * 2
        with torch.no_grad():
            want = ref(x, t, c)
        net.set_context(c)
        assert rel(net.forward(x, t), want) < 1e-5


@pytest.mark.parametrize("chains", [2, 3])
def test_batch_chains_partition_is_invisible(emulated, chains):
    """rdm_unet_set_chains under emulation: the host-side partition (arena slices with guard zones, statistics slices, context / time-embedding /
    input / output row offsets, CFG doubling across the chain boundary) gives the single-chain result; the streams run in issue order here,
    the concurrency itself is covered by tests/test_unet_gpu.py::test_batch_chains_do_not_change_the_result."""
    from rdm_b200 import sampler
    ref, net = _pair(5)
    g = torch.Generator().manual_seed(9)
    B = 3
    x_T = torch.randn(B, 4, 8, 8, generator=g)
    c, uc = torch.randn(B, 2, 512, generator=g) * 2, torch.zeros(B, 2, 512)
    t = torch.tensor([501, 12, 700, 3, 999, 250])
    with torch.no_grad():
        want_f = ref(torch.cat([x_T] * 2), t, torch.cat([c, uc]))
    want = oddim.ddim_sample(ref, x_T, c, uc, S=4, scale=2.0)
    tb = sampler.make_ddim_tables(sampler.alphas_cumprod_linear(), 4, 0.0, device="cpu")
    net.set_chains(chains)
    net.set_context(torch.cat([c, uc]))
    assert rel(net.forward(x_T, t), want_f) < 1e-5            # eager + capture
    assert rel(net.forward(x_T, t), want_f) < 1e-5            # replay
    assert rel(net.ddim_sample(x_T, tb["timesteps"], tb["coef"], cfg_scale=2.0), want) < 1e-5
    net.set_chains(1)
    assert rel(net.forward(x_T, t), want_f) < 1e-5
