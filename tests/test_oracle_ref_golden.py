"""The oracle against outputs of the REFERENCE's own code (tests/golden/ref_*.npz, written by tests/golden/make_golden_ref.py from
/root/reference in the build container): U-Net forward (openaimodel.py + attention.py), the DDIM loop with classifier-free guidance
(ddim.py) and the state-dict key layout.  This is what pins oracle/unet.py and oracle/ddim.py; the GPU tests then hold
the CUDA path to the same files (tests/test_zy_ref_golden_gpu.py)."""
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


# ---- the reference's MinimalRETRODiffusion.sample_from_rdata / sample_with_query run end to end (ref_pipeline_tiny.npz) --------------
def _pipeline():
    p = load("ref_pipeline_tiny.npz")
    db, mem, id_count = ref_weights.make_db(int(p["n_db"]))
    assert np.array_equal(mem, p["nn_memory"]) and list(id_count.values()) == [int(v) for v in p["id_count_vals"]]
    d = load("ref_unet_tiny.npz")
    cfg = ast.literal_eval(str(d["cfg_json"]))
    ema = ref_weights.fill_(ounet.UNetModel(**cfg).eval(), int(p["ema_seed"]))         # sampling runs under ema_scope (ddpm.py:977)
    return p, db, mem, id_count, ema


def test_sample_from_rdata_pipeline_matches_reference_code():
    """qids -> raw DB rows -> q / |q| -> exact top-k -> RAW neighbour rows as context, zeros as the unconditional context (label 0),
    EMA weights, guided DDIM: the oracle pipeline the GPU mirror tests compare against == the reference's own orchestration."""
    from oracle import knn as oknn
    p, db, mem, id_count, ema = _pipeline()
    qids, k = p["rdata:qids"], int(p["k_nn"])
    nns, _ = oknn.search(db, oknn.normalize_queries(db[qids].astype(np.float32)), k)
    assert list(nns[:, 0]) == list(qids)                                                # a DB row retrieves itself first (ddpm.py:897)
    cond = torch.from_numpy(db[nns].astype(np.float32))
    x = oddim.ddim_sample(ema, torch.from_numpy(p["rdata:x_T"]), cond, torch.zeros_like(cond), S=4, scale=2.0)
    assert rel(x, p["rdata:samples"]) < 5e-6


def test_sample_with_query_pipeline_matches_reference_code():
    from oracle import knn as oknn
    p, db, mem, id_count, ema = _pipeline()
    q, k = p["query:q"], int(p["k_nn"])
    nns, _ = oknn.search(db, oknn.normalize_queries(q), k)
    r = torch.from_numpy(db[nns].astype(np.float32))
    xT = torch.from_numpy(p["rdata:x_T"])[:2]
    cond = torch.cat([torch.from_numpy(q)[:, None], r[:, :k - 1]], 1)                   # the query itself is neighbour 0 (ddpm.py:775)
    assert rel(oddim.ddim_sample(ema, xT, cond, torch.zeros_like(cond), S=4, scale=2.0), p["query:samples"]) < 5e-6
    assert rel(oddim.ddim_sample(ema, xT, r, torch.zeros_like(r), S=4, scale=2.0), p["query_omit:samples"]) < 5e-6      # omit_query (:772-773)


def test_mirror_host_functions_match_reference_code(tmp_path):
    """get_qids (NumPy global RNG, top-m memory, frequency weights) and get_unconditional_conditioning of the mirror vs the reference's."""
    import copy
    import pickle
    import rdm  # noqa: F401
    from ldm.util import instantiate_from_config
    from omegaconf import OmegaConf
    from test_mirror_host import TINY_CFG
    p, db, mem, id_count, _ = _pipeline()
    with open(tmp_path / "nn_memory.p", "wb") as f:
        pickle.dump({"nn_memory": mem, "id_count": id_count}, f)
    cfg = copy.deepcopy(TINY_CFG)
    cfg["params"]["nn_memory"] = str(tmp_path / "nn_memory.p")
    m = instantiate_from_config(OmegaConf.create(cfg))
    np.random.seed(44)
    assert np.array_equal(m.get_qids(50, 3, use_weights=False), p["rdata:qids"])
    np.random.seed(45)
    assert np.array_equal(m.get_qids(0.4, 5, use_weights=True), p["qids_weighted"])
    m.unconditional_guidance_vex.copy_(torch.from_numpy(p["vex"]))
    uc = m.get_unconditional_conditioning((2, 4, 512), unconditional_guidance_label=1.5, k_nn=4)
    assert uc.shape == (2, 4, 512) and float((uc - torch.from_numpy(p["uncond_label_1.5"])).abs().max()) < 1e-6
    # the checkpoint keys the reference model reports missing for this synthetic checkpoint are the ones the mirror reports
    live = {"model.diffusion_model." + k: torch.zeros(s) for k, s in __import__("rdm_b200.unet", fromlist=["x"]).unet_param_shapes(**ounet.TINY_UNET).items()}
    live.update({"model_ema." + k[len("model."):].replace(".", ""): v for k, v in live.items()})
    missing, unexpected = m.load_state_dict(live, strict=False)
    sched = {"betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod",
             "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod"}            # ldm's schedule buffers (the generator's base-class stand-in keeps three)
    assert not unexpected and set(missing) - sched == {str(k) for k in p["missing_keys"]} - sched
