"""GPU parity at the configurations bench.py actually measures and at the reference's shipped shape ("do this" 1b).

* cfg2-B: bench.UNET / bench.make_weights() on the 32x32x4 latent, DDIM-100, CFG 2.0, batch 16, the DEFAULT engine mode and the other
  tensor-core modes, 3+ seeds: final latents within the north-star 1e-3 rel-L2 of the fp32 oracle (tests/golden/ddim100_cfg2.npz, written by
  tests/golden/make_golden_ddim100.py from oracle/unet.py + oracle/ddim.py).
* R: models/rdm/imagenet/config.yaml:14-59 (64x64x3 latent, in_channels 3): one forward strict <= 1e-4, tensor-core modes, DDIM-20.
* kNN at BASELINE sizes: 1,281,167 x 512 fp16 with 16 queries, and Q in {64, 65, 256, 1024} with k = 20 (multi-pass path): indices and
  fp64 scores bit-identical to oracle/knn_ref.c.
"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


@pytest.fixture(scope="module")
def bench_net(cuda):
    import bench
    from rdm_b200.unet import B200UNet
    net = B200UNet(cuda, **bench.UNET)
    net.load_state_dict(bench.make_weights())
    return net


# mode ids: rdm_unet_set_mode (1 bf16x3, 3 fp16x2, 4 fp16).  Every mode must meet the north-star tolerance on EVERY seed.
@pytest.mark.parametrize("mode,tol", [(3, 1e-3), (4, 1e-3), (1, 1e-4)])
def test_cfg2_ddim100_batch16_final_latents(cuda, bench_net, mode, tol):
    import make_golden_ddim100 as gen
    from rdm_b200 import sampler
    want_all = torch.from_numpy(np.load(os.path.join(GOLD, "ddim100_cfg2.npz"))["latents"])
    assert want_all.shape[0] >= 3 and want_all.shape[1:] == (gen.BATCH, 4, 32, 32)
    bench_net.set_mode(mode)
    tb = sampler.make_ddim_tables(sampler.alphas_cumprod_linear(), gen.S_DDIM, 0.0, device=cuda)
    errs = []
    for seed in range(want_all.shape[0]):
        x_T, cond, unc = gen.inputs_cfg2(seed)
        bench_net.set_context(torch.cat([cond, unc]).to(cuda))
        got = bench_net.ddim_sample(x_T.to(cuda), tb["timesteps"], tb["coef"], cfg_scale=gen.SCALE)
        errs.append(rel_l2(got, want_all[seed]))
        per_image = [rel_l2(got[i], want_all[seed][i]) for i in range(gen.BATCH)]
        assert max(per_image) < 3 * tol, f"mode {mode} seed {seed}: worst single image {max(per_image):.2e}"
    print(f"cfg2 DDIM-100 batch 16, mode {mode}: rel-L2 per seed {['%.2e' % e for e in errs]}")
    assert max(errs) < tol, f"mode {mode}: {errs}"


@pytest.fixture(scope="module")
def rshape(cuda):
    from oracle import unet as ounet
    from rdm_b200.unet import B200UNet
    ref = ounet.randomize_(ounet.UNetModel(**ounet.IMAGENET_UNET), 3).eval()
    net = B200UNet(cuda, **ounet.IMAGENET_UNET)
    net.load_state_dict(ref.state_dict())
    assert net.missing() == 0
    return net, np.load(os.path.join(GOLD, "rshape_imagenet.npz"))


@pytest.mark.parametrize("mode,tol", [(0, 1e-4), (1, 1e-4), (3, 3e-3), (4, 4e-3)])
def test_reference_shape_forward(cuda, rshape, mode, tol):
    """64x64x3 latent, in_channels 3 (the shape every shipped RDM checkpoint uses)."""
    import make_golden_ddim100 as gen
    net, g = rshape
    x, t, c, _, _, _ = gen.inputs_rshape()
    net.set_mode(mode)
    net.set_context(c.to(cuda))
    got = net.forward(x.to(cuda), t.to(cuda))
    err = rel_l2(got, torch.from_numpy(g["forward"]))
    print(f"R-shape forward mode {mode}: rel-L2 {err:.2e}")
    assert err < tol


@pytest.mark.parametrize("mode,tol", [(1, 1e-4), (3, 1e-3), (4, 1e-3)])
def test_reference_shape_ddim20(cuda, rshape, mode, tol):
    import make_golden_ddim100 as gen
    from rdm_b200 import sampler
    net, g = rshape
    _, _, _, x_T, cond, unc = gen.inputs_rshape()
    net.set_mode(mode)
    tb = sampler.make_ddim_tables(sampler.alphas_cumprod_linear(), 20, 0.0, device=cuda)
    net.set_context(torch.cat([cond, unc]).to(cuda))
    got = net.ddim_sample(x_T.to(cuda), tb["timesteps"], tb["coef"], cfg_scale=gen.SCALE)
    err = rel_l2(got, torch.from_numpy(g["ddim20"]))
    print(f"R-shape DDIM-20 mode {mode}: rel-L2 {err:.2e}")
    assert err < tol


@pytest.mark.parametrize("mode,tol", [(4, 1e-3), (3, 1e-3)])
def test_reference_shape_ddim100(cuda, rshape, mode, tol):
    """The shipped shape at the headline step count (DDIM-100, CFG 2.0, two images) in the default single-MMA fp16 mode and in fp16x2."""
    import make_golden_ddim100 as gen
    from rdm_b200 import sampler
    net, _ = rshape
    want = torch.from_numpy(np.load(os.path.join(GOLD, "rshape_imagenet_ddim100.npz"))["ddim100"])
    x_T, cond, unc = gen.inputs_rshape100()
    net.set_mode(mode)
    tb = sampler.make_ddim_tables(sampler.alphas_cumprod_linear(), gen.S_DDIM, 0.0, device=cuda)
    net.set_context(torch.cat([cond, unc]).to(cuda))
    got = net.ddim_sample(x_T.to(cuda), tb["timesteps"], tb["coef"], cfg_scale=gen.SCALE)
    err = rel_l2(got, want)
    print(f"R-shape DDIM-100 mode {mode}: rel-L2 {err:.2e}")
    assert err < tol


# ---------------------------------------------------------------------------------------------------------------- kNN at BASELINE sizes
def _bits(a):
    return np.ascontiguousarray(a).view(np.int64)


@pytest.fixture(scope="module")
def imagenet_db(cuda):
    """cfg2's database: 1,281,167 x 512, N(0,1) rows stored fp16, seed 1."""
    from rdm_b200.knn import B200Searcher
    rng = np.random.default_rng(1)
    db = rng.standard_normal((1_281_167, 512), dtype=np.float32).astype(np.float16)
    return db, B200Searcher(db, device=cuda)


def _check(db, searcher, q_rows, k, cuda):
    from oracle import knn as oknn
    qh = oknn.normalize_queries(db[q_rows].astype(np.float32))
    idx, dist, score = searcher.search_device(torch.from_numpy(qh).to(cuda), k, return_scores=True)
    ref_idx, ref_dist, ref_score = oknn.search(db, qh, k, return_scores=True)
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert np.array_equal(_bits(score.cpu().numpy()), _bits(ref_score))
    assert np.array_equal(dist.cpu().numpy(), ref_dist)
    assert np.array_equal(ref_idx[:, 0], q_rows)                      # a DB-row query retrieves itself first (ddpm.py:897)


def test_knn_imagenet_size_16_queries_bit_exact(cuda, imagenet_db):
    db, s = imagenet_db
    _check(db, s, np.random.default_rng(2).integers(0, db.shape[0], size=16), 4, cuda)


@pytest.mark.parametrize("Q", [64, 65, 256, 1024])
def test_knn_imagenet_size_many_queries_k20_bit_exact(cuda, imagenet_db, Q):
    db, s = imagenet_db
    _check(db, s, np.random.default_rng(100 + Q).integers(0, db.shape[0], size=Q), 20, cuda)


def test_knn_fp32_db_bit_exact_300k(cuda):
    from rdm_b200.knn import B200Searcher
    rng = np.random.default_rng(7)
    db = rng.standard_normal((300_000, 512), dtype=np.float32)
    s = B200Searcher(db, device=cuda)
    _check(db, s, rng.integers(0, db.shape[0], size=37), 8, cuda)
