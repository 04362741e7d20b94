"""The oracle against outputs of the REFERENCE's own code (tests/golden/ref_*.npz, written by tests/golden/make_golden_ref.py from
/root/reference in the build container): U-Net forward (openaimodel.py + attention.py), the DDIM loop with classifier-free guidance
(ddim.py) and the state-dict key layout.  This is what pins oracle/unet.py and oracle/ddim.py; the GPU tests then hold
the CUDA path to the same files (tests/test_ref_golden_gpu.py)."""
import ast
import os
import sys

import numpy as np
import torch

from conftest import ROOT
from oracle import ddim as oddim, unet as ounet

GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLD)
import ref_weights  # noqa: E402


def load(name):
    return np.load(os.path.join(GOLD, name))


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm())


def oracle_unet(d):
    cfg = ast.literal_eval(str(d["cfg_json"]))
    net = ounet.UNetModel(**cfg).eval()
    assert list(net.state_dict().keys()) == [str(k) for k in d["sd_keys"]]            # key layout AND registration order of the reference
    assert sum(p.numel() for p in net.parameters()) == int(d["n_params"])
    return ref_weights.fill_(net, int(d["weight_seed"]))


def test_unet_forward_matches_reference_code():
    d = load("ref_unet_tiny.npz")
    net = oracle_unet(d)
    with torch.no_grad():
        y = net(torch.from_numpy(d["x"]), torch.from_numpy(d["t"]), torch.from_numpy(d["context"]))
    assert rel(y, d["out"]) < 2e-6


def test_ddim_schedule_matches_reference_code():
    g = load("ref_ddim_tiny.npz")
    sch = oddim.Schedule(int(g["S"]), 0.0)
    assert np.array_equal(sch.timesteps, g["ddim_timesteps"])
    assert np.array_equal(np.asarray(sch.alphas), g["ddim_alphas"]) and np.array_equal(np.asarray(sch.alphas_prev), g["ddim_alphas_prev"])
    assert np.allclose(oddim.Schedule(int(g["S"]), 0.5).sigmas, g["eta05:ddim_sigmas"], rtol=1e-7, atol=0)


def test_ddim_loop_with_guidance_matches_reference_code():
    d, g = load("ref_unet_tiny.npz"), load("ref_ddim_tiny.npz")
    net = oracle_unet(d)
    xT, c, uc = (torch.from_numpy(g[k]) for k in ("x_T", "cond", "uncond"))
    for tag, eta in (("eta0", 0.0), ("eta05", 0.5)):
        torch.manual_seed(int(g[f"{tag}:seed"]))                  # the reference draws one randn(x.shape) per step (ddim.py:226-227)
        x, traj = oddim.ddim_sample(net, xT, c, uc, S=int(g["S"]), scale=float(g["scale"]), eta=eta, return_all=True, draw_noise_always=True)
        assert rel(x, g[f"{tag}:samples"]) < 5e-6, tag
        for i, (xi, p0) in enumerate(traj):
            assert rel(xi, g[f"{tag}:x_inter"][i]) < 5e-6 and rel(p0, g[f"{tag}:pred_x0"][i]) < 5e-6, (tag, i)
    x = oddim.ddim_sample(net, xT, c, None, S=4, scale=1.0)
    assert rel(x, g["noguid:samples"]) < 5e-6


def test_product_parameter_inventory_is_the_reference_state_dict():
    """Host side of the C ABI: the names/shapes the U-Net executor registers == the keys the reference's UNetModel creates."""
    from rdm_b200.unet import unet_param_shapes
    d = load("ref_unet_tiny.npz")
    cfg = ast.literal_eval(str(d["cfg_json"]))
    cfg.pop("use_spatial_transformer")
    shapes = unet_param_shapes(**cfg)
    assert list(shapes) == [str(k) for k in d["sd_keys"]]
    assert sum(int(np.prod(s)) for s in shapes.values()) == int(d["n_params"])
