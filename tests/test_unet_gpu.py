"""GPU parity: librdm_b200 U-Net forward / DDIM (through the C ABI) vs the torch-CPU fp32 oracle."""
import numpy as np
import pytest
import torch

from oracle import ddim as oddim
from oracle import unet as ounet

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


def _pair(cfg, seed, cuda):
    from rdm_b200.unet import B200UNet
    ref = ounet.randomize_(ounet.UNetModel(**cfg), seed).eval()
    net = B200UNet(cuda, **cfg)
    net.load_state_dict(ref.state_dict())
    assert net.missing() == 0
    return ref, net


@pytest.mark.parametrize("B2,H,k", [(2, 16, 4), (3, 8, 1), (4, 24, 8)])
def test_tiny_unet_forward_strict(cuda, B2, H, k):
    ref, net = _pair(ounet.TINY_UNET, 1, cuda)
    g = torch.Generator().manual_seed(B2 * 100 + H)
    x = torch.randn(B2, 4, H, H, generator=g)
    t = torch.randint(0, 1000, (B2,), generator=g)
    c = torch.randn(B2, k, 512, generator=g) * 3
    with torch.no_grad():
        want = ref(x, t, c)
    net.set_context(c.to(cuda))
    got = net.forward(x.to(cuda), t.to(cuda))
    err = rel_l2(got, want)
    assert err < 1e-4, f"strict fp32 forward rel-L2 {err:.2e} (tolerance 1e-4, SURVEY 8d)"


def test_cfg_doubling_without_materialising_the_batch(cuda):
    ref, net = _pair(ounet.TINY_UNET, 2, cuda)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 4, 16, 16, generator=g)
    t = torch.full((4,), 501)
    c = torch.cat([torch.randn(2, 4, 512, generator=g), torch.zeros(2, 4, 512)])      # [cond | uncond=0] (ddpm.py:673-680, label 0)
    with torch.no_grad():
        want = ref(torch.cat([x] * 2), t, c)
    net.set_context(c.to(cuda))
    got = net.forward(x.to(cuda), t.to(cuda))          # Bx = B2/2
    assert rel_l2(got, want) < 1e-4


def test_full_arch_forward_strict(cuda):
    """BASELINE cfg2 architecture (mc 192, mult 1-2-3-5, 16 SpatialTransformers) on the 32x32x4 latent, B2 = 2."""
    ref, net = _pair(ounet.BASELINE_UNET, 3, cuda)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 4, 32, 32, generator=g)
    t = torch.tensor([991, 991])
    c = torch.cat([torch.randn(1, 4, 512, generator=g) * 3, torch.zeros(1, 4, 512)])
    with torch.no_grad():
        want = ref(x, t, c)
    net.set_context(c.to(cuda))
    got = net.forward(x.to(cuda), t.to(cuda))
    err = rel_l2(got, want)
    assert err < 1e-4, f"rel-L2 {err:.2e}"


@pytest.mark.parametrize("mode,tol", [(1, 1e-4), (2, 3e-2), (3, 3e-3), (4, 4e-3)])
@pytest.mark.parametrize("B2,H", [(2, 16), (4, 32), (3, 8)])
def test_tiny_unet_forward_tensor_core(cuda, mode, tol, B2, H):
    """tcgen05 engine: mode 1 = bf16 hi/lo split (3 MMAs per product, fp32-grade), mode 2 = plain bf16."""
    ref, net = _pair(ounet.TINY_UNET, 1, cuda)
    net.set_mode(mode)
    g = torch.Generator().manual_seed(B2 * 100 + H)
    x = torch.randn(B2, 4, H, H, generator=g)
    t = torch.randint(0, 1000, (B2,), generator=g)
    c = torch.randn(B2, 4, 512, generator=g) * 3
    with torch.no_grad():
        want = ref(x, t, c)
    net.set_context(c.to(cuda))
    got = net.forward(x.to(cuda), t.to(cuda))
    err = rel_l2(got, want)
    assert err < tol, f"mode {mode}: rel-L2 {err:.2e} (tolerance {tol})"


@pytest.mark.parametrize("mode,tol", [(1, 1e-4), (2, 3e-2), (3, 3e-3), (4, 4e-3)])
def test_full_arch_forward_tensor_core(cuda, mode, tol):
    ref, net = _pair(ounet.BASELINE_UNET, 3, cuda)
    net.set_mode(mode)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 4, 32, 32, generator=g)
    t = torch.tensor([991, 991])
    c = torch.cat([torch.randn(1, 4, 512, generator=g) * 3, torch.zeros(1, 4, 512)])
    with torch.no_grad():
        want = ref(x, t, c)
    net.set_context(c.to(cuda))
    got = net.forward(x.to(cuda), t.to(cuda))
    err = rel_l2(got, want)
    assert err < tol, f"mode {mode}: rel-L2 {err:.2e}"


def test_ddim_step_is_bit_exact(cuda):
    from rdm_b200.unet import ddim_step
    sch = oddim.Schedule(100)
    g = torch.Generator().manual_seed(3)
    x, e = torch.randn(3, 4, 8, 8, generator=g), torch.randn(6, 4, 8, 8, generator=g)
    for index in (0, 50, 99):
        a_t, a_prev, sigma, s1m = sch.coeffs(index)
        e_t = e[3:] + 2.0 * (e[:3] - e[3:])
        xp_ref, p0_ref = oddim.ddim_update(x, e_t, a_t, a_prev, sigma, s1m)
        coef = torch.stack([s1m, a_t.sqrt(), a_prev.sqrt(), (1.0 - a_prev - sigma ** 2).sqrt(), sigma]).to(cuda)
        xp, p0 = ddim_step(x.to(cuda), e.to(cuda), coef, cfg_scale=2.0)
        assert torch.equal(xp.cpu(), xp_ref) and torch.equal(p0.cpu(), p0_ref)


def test_ddim_trajectory_tiny(cuda):
    """10 chained CFG steps with the tiny U-Net: final latents within 1e-3 rel of the fp32 oracle (north_star tolerance)."""
    from rdm_b200.unet import ddim_step
    ref, net = _pair(ounet.TINY_UNET, 5, cuda)
    g = torch.Generator().manual_seed(21)
    B, S, scale = 2, 10, 2.0
    x_T = torch.randn(B, 4, 16, 16, generator=g)
    cond, unc = torch.randn(B, 4, 512, generator=g) * 3, torch.zeros(B, 4, 512)
    want = oddim.ddim_sample(ref, x_T, cond, unc, S=S, scale=scale)
    sch = oddim.Schedule(S)
    net.set_context(torch.cat([cond, unc]).to(cuda))
    x = x_T.to(cuda)
    for i, step in enumerate(np.flip(sch.timesteps)):
        index = S - i - 1
        a_t, a_prev, sigma, s1m = sch.coeffs(index)
        coef = torch.stack([s1m, a_t.sqrt(), a_prev.sqrt(), (1.0 - a_prev - sigma ** 2).sqrt(), sigma]).to(cuda)
        eps = net.forward(x, torch.full((2 * B,), int(step), device=cuda))
        x, _ = ddim_step(x, eps, coef, cfg_scale=scale)
    err = rel_l2(x, want)
    assert err < 1e-3, f"DDIM-10 final latent rel-L2 {err:.2e}"


def _tables(S, cuda):
    from rdm_b200 import sampler
    return sampler.make_ddim_tables(sampler.alphas_cumprod_linear(), S, 0.0, device=cuda)


@pytest.mark.parametrize("mode", [0, 1])
def test_fused_ddim_loop_matches_oracle_tiny(cuda, mode):
    """rdm_ddim_sample (one captured CUDA graph replayed per step) vs the oracle sampler: DDIM-10, CFG 2.0, zeros uncond."""
    ref, net = _pair(ounet.TINY_UNET, 5, cuda)
    net.set_mode(mode)
    g = torch.Generator().manual_seed(21)
    B, S = 3, 10
    x_T = torch.randn(B, 4, 16, 16, generator=g)
    cond, unc = torch.randn(B, 4, 512, generator=g) * 3, torch.zeros(B, 4, 512)
    want = oddim.ddim_sample(ref, x_T, cond, unc, S=S, scale=2.0)
    tb = _tables(S, cuda)
    net.set_context(torch.cat([cond, unc]).to(cuda))
    got, p0 = net.ddim_sample(x_T.to(cuda), tb["timesteps"], tb["coef"], cfg_scale=2.0, want_pred_x0=True)
    assert rel_l2(got, want) < 1e-3
    # a second call replays the captured graph with a new x_T and a new context of the same shape
    x2 = torch.randn(B, 4, 16, 16, generator=g)
    cond2 = torch.randn(B, 4, 512, generator=g)
    want2 = oddim.ddim_sample(ref, x2, cond2, unc, S=S, scale=2.0)
    net.set_context(torch.cat([cond2, unc]).to(cuda))
    got2 = net.ddim_sample(x2.to(cuda), tb["timesteps"], tb["coef"], cfg_scale=2.0)
    assert rel_l2(got2, want2) < 1e-3
    # split ranges (what DDIMSampler does around intermediates) give the same trajectory
    net.set_context(torch.cat([cond, unc]).to(cuda))
    a = net.ddim_sample(x_T.to(cuda), tb["timesteps"], tb["coef"], cfg_scale=2.0, first_step=0, num_steps=4)
    b = net.ddim_sample(a, tb["timesteps"], tb["coef"], cfg_scale=2.0, first_step=4, num_steps=6)
    assert rel_l2(b, got) < 1e-4          # not bit-equal: GroupNorm statistics are accumulated with fp64 atomics (order varies)


def test_no_cfg_path(cuda):
    ref, net = _pair(ounet.TINY_UNET, 6, cuda)
    g = torch.Generator().manual_seed(5)
    x_T, cond = torch.randn(2, 4, 16, 16, generator=g), torch.randn(2, 2, 512, generator=g)
    want = oddim.ddim_sample(ref, x_T, cond, None, S=5, scale=1.0)
    tb = _tables(5, cuda)
    net.set_context(cond.to(cuda))
    got = net.ddim_sample(x_T.to(cuda), tb["timesteps"], tb["coef"], cfg_scale=1.0)
    assert rel_l2(got, want) < 1e-3


@pytest.mark.parametrize("mode,tol", [(1, 1e-4), (3, 1e-3)])
def test_full_arch_ddim20_tensor_core(cuda, mode, tol):
    """BASELINE cfg2 architecture, 20 chained CFG steps, one image: denoised latents within 1e-3 rel of the fp32 oracle."""
    ref, net = _pair(ounet.BASELINE_UNET, 3, cuda)
    net.set_mode(mode)
    g = torch.Generator().manual_seed(31)
    x_T = torch.randn(1, 4, 32, 32, generator=g)
    cond, unc = torch.randn(1, 4, 512, generator=g) * 3, torch.zeros(1, 4, 512)
    want = oddim.ddim_sample(ref, x_T, cond, unc, S=20, scale=2.0)
    tb = _tables(20, cuda)
    net.set_context(torch.cat([cond, unc]).to(cuda))
    got = net.ddim_sample(x_T.to(cuda), tb["timesteps"], tb["coef"], cfg_scale=2.0)
    err = rel_l2(got, want)
    print(f"full-arch DDIM-20 mode {mode}: rel-L2 {err:.3e}")
    assert err < tol


@pytest.mark.parametrize("mode,tol", [(0, 1e-5), (3, 3e-3)])
@pytest.mark.parametrize("chains", [2, 3, 8])
def test_batch_chains_do_not_change_the_result(cuda, mode, tol, chains):
    """rdm_unet_set_chains: sub-batches of a forward run concurrently on their own streams (graph branches).  GroupNorm / LayerNorm /
    attention are per sample, so the split is invisible: forward and the fused DDIM loop agree with the single-chain run (to the
    summation-order noise of atomics / split-K) and with the oracle, for batch sizes that do not divide evenly and for CFG doubling."""
    ref, net = _pair(ounet.TINY_UNET, 5, cuda)
    net.set_mode(mode)
    g = torch.Generator().manual_seed(77)
    B = 5
    x_T = torch.randn(B, 4, 16, 16, generator=g)
    cond, unc = torch.randn(B, 4, 512, generator=g) * 3, torch.zeros(B, 4, 512)
    t = torch.randint(0, 1000, (2 * B,), generator=g)
    with torch.no_grad():
        want_f = ref(torch.cat([x_T] * 2), t, torch.cat([cond, unc]))
    want = oddim.ddim_sample(ref, x_T, cond, unc, S=5, scale=2.0)
    tb = _tables(5, cuda)
    outs = []
    for ch in (1, chains):
        net.set_chains(ch)
        net.set_context(torch.cat([cond, unc]).to(cuda))
        f = net.forward(x_T.to(cuda), t.to(cuda))              # Bx = B2 / 2: chains cross the cond / uncond boundary
        f2 = net.forward(x_T.to(cuda), t.to(cuda))             # graph replay
        d = net.ddim_sample(x_T.to(cuda), tb["timesteps"], tb["coef"], cfg_scale=2.0)
        assert rel_l2(f, want_f) < tol and rel_l2(f2, f) < 1e-5 and rel_l2(d, want) < max(tol, 1e-3)
        outs.append((f, d))
    assert rel_l2(outs[1][0], outs[0][0]) < (1e-5 if mode == 0 else 1e-3)
    assert rel_l2(outs[1][1], outs[0][1]) < (1e-5 if mode == 0 else 2e-3)
    net.set_chains(1)
