"""CPU: pin the oracle against every known answer the reference repository contains."""
import numpy as np
import torch

from oracle import ddim, knn, unet


def test_unet_param_count_matches_reference_notebook():
    # scripts/demo_rdm.ipynb:128 prints 400.92 M params; :129 'Keeping EMAs of 690' = 688 tensors + 2 LitEma buffers
    m = unet.UNetModel(**unet.IMAGENET_UNET)
    ps = list(m.parameters())
    assert sum(p.numel() for p in ps) == 400_920_579 and len(ps) == 688
    # scripts/demo_rdm.ipynb:112-127 head counts at d_head 32
    assert m.head_counts == [12, 12, 18, 18, 30, 30, 30, 30, 30, 30, 18, 18, 18, 12, 12, 12]


def test_unet_state_dict_key_layout():
    # checkpoint keys the loader must accept
    keys = set(unet.UNetModel(**unet.IMAGENET_UNET).state_dict())
    for k in ["time_embed.0.weight", "input_blocks.0.0.weight", "input_blocks.1.0.in_layers.2.weight",
              "input_blocks.1.0.emb_layers.1.bias", "input_blocks.1.0.out_layers.3.weight", "input_blocks.3.0.op.weight",
              "input_blocks.4.0.skip_connection.weight", "input_blocks.4.1.proj_in.weight",
              "input_blocks.4.1.transformer_blocks.0.attn2.to_k.weight", "input_blocks.4.1.transformer_blocks.0.ff.net.0.proj.weight",
              "input_blocks.4.1.transformer_blocks.0.ff.net.2.bias", "input_blocks.4.1.transformer_blocks.0.norm3.weight",
              "middle_block.1.proj_out.bias", "output_blocks.2.2.conv.weight", "out.2.weight"]:
        assert k in keys, k


def test_ddim_schedule_known_answers():
    ts = ddim.make_ddim_timesteps(100)
    assert ts[0] == 1 and ts[-1] == 991 and len(ts) == 100 and (np.diff(ts) == 10).all()
    ts = ddim.make_ddim_timesteps(250)
    assert ts[0] == 1 and ts[-1] == 997
    ac = ddim.alphas_cumprod_f32()
    assert ac.dtype == np.float32 and ac.shape == (1000,) and 0.99 < ac[0] < 1 and ac[-1] < 0.01
    s = ddim.Schedule(100)
    assert (s.sigmas == 0).all() and s.alphas_prev[0] == ac[0] and s.alphas_prev[1] == ac[1]


def test_ddim_update_is_identity_preserving():
    # with eps = 0 the x0 prediction is x / sqrt(a_t) and x_prev = sqrt(a_prev) * x0
    s = ddim.Schedule(10)
    x = torch.randn(2, 4, 8, 8)
    xp, x0 = ddim.ddim_update(x, torch.zeros_like(x), *s.coeffs(5))
    a_t, a_prev = s.coeffs(5)[:2]
    assert torch.allclose(x0, x / a_t.sqrt()) and torch.allclose(xp, a_prev.sqrt() * x0)


def test_knn_c_oracle_matches_numpy_and_breaks_ties_by_lowest_index():
    rng = np.random.default_rng(0)
    for dt in (np.float32, np.float16):
        db = rng.standard_normal((20000, 512)).astype(dt)
        db[1000] = db[5]
        db[7777] = db[5]
        q = knn.normalize_queries(db[[5, 17, 300]].astype(np.float32))
        i1, d1 = knn.search(db, q, 8)
        i2, d2 = knn.search_numpy(db, q, 8)
        assert (i1 == i2).all() and np.abs(d1 - d2).max() < 1e-6
        assert list(i1[0, :3]) == [5, 1000, 7777]          # exact duplicates: lowest index first
        assert i1[1, 0] == 17 and abs(d1[1, 0] - 1) < 1e-6  # a DB row retrieves itself (ddpm.py:897)


def test_knn_sharding_is_result_invariant():
    rng = np.random.default_rng(1)
    db = rng.standard_normal((6000, 512)).astype(np.float16)
    q = knn.normalize_queries(rng.standard_normal((5, 512)))
    full = knn.search(db, q, 4, return_scores=True)
    parts = []
    for s in range(0, 6000, 1500):
        i, _, sc = knn.search(db[s:s + 1500], q, 4, idx_base=s, return_scores=True)
        parts.append((i, sc))
    mi, ms = knn.merge_shards(parts, 4)
    assert (mi == full[0]).all() and (ms == full[2]).all()


def test_knn_oracle_agrees_with_an_independent_exact_searcher():
    """ScaNN (the reference's searcher, scann==1.2.4) is not installable here, so the exact semantics it approximates are checked against an
    independent third-party implementation instead: scikit-learn's brute-force cosine NearestNeighbors on the raw rows.  On tie-free data
    the neighbour lists must be identical and the reported dot products equal 1 - cosine distance."""
    import pytest
    sklearn_neighbors = pytest.importorskip("sklearn.neighbors")
    rng = np.random.default_rng(12)
    db = (rng.standard_normal((20_000, 512)) * rng.uniform(0.3, 9.0, (20_000, 1))).astype(np.float16)      # raw rows with very different norms
    q = rng.standard_normal((37, 512)).astype(np.float32)
    q[:5] = db[[3, 777, 4096, 19_999, 12_345]].astype(np.float32)                                         # DB rows as queries (ddpm.py:897)
    qh = knn.normalize_queries(q)
    idx, dist = knn.search(db, qh, 20)
    nn = sklearn_neighbors.NearestNeighbors(n_neighbors=20, algorithm="brute", metric="cosine").fit(db.astype(np.float64))
    d_sk, i_sk = nn.kneighbors(q.astype(np.float64))
    assert np.array_equal(idx, i_sk)
    assert np.allclose(dist, 1.0 - d_sk, atol=2e-6)
    assert list(idx[:5, 0]) == [3, 777, 4096, 19_999, 12_345] and np.allclose(dist[:5, 0], 1.0, atol=1e-6)


def test_query_normalisation_restatement_matches_numpy_bitwise():
    """oracle/knn.py pairwise_sum_f32 / normalize_queries_restated (the order csrc/knn.cu knn_normalize_kernel follows) against NumPy's own
    `q / np.linalg.norm(q, axis=1)[:, np.newaxis]` (ddpm.py:907), every supported width plus ragged ones, bit patterns compared."""
    from oracle import knn as oknn
    rng = np.random.default_rng(12)
    for d in (5, 8, 100, 128, 136, 256, 512, 768, 1024, 1000):
        q = (rng.standard_normal((6, d)) * rng.uniform(0.01, 30.0, size=(6, 1))).astype(np.float32)
        want = q / np.linalg.norm(q, axis=1)[:, np.newaxis]
        got = oknn.normalize_queries_restated(q)
        assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), want.view(np.uint32)), d
        assert np.array_equal(oknn.normalize_queries(q).view(np.uint32), want.view(np.uint32))
