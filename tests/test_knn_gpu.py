"""GPU parity: librdm_b200 kNN (through the C ABI) vs the exact oracle.  Indices bit-exact, fp64 scores bit-exact."""
import numpy as np
import pytest
import torch

from oracle import knn as oknn

pytestmark = pytest.mark.gpu


def _db(n, d, dtype, seed, dup=True):
    rng = np.random.default_rng(seed)
    db = rng.standard_normal((n, d)).astype(np.float32)
    db *= rng.uniform(0.5, 12.0, size=(n, 1)).astype(np.float32)        # raw CLIP rows are NOT unit norm (F7)
    db = db.astype(dtype)
    if dup and n > 5000:
        db[1234] = db[7]; db[4321] = db[7]; db[n - 1] = db[7]            # exact duplicates -> score ties
    return db


@pytest.mark.parametrize("dtype", [np.float16, np.float32])
@pytest.mark.parametrize("n,nq,k", [(50_000, 1, 4), (100_003, 16, 4), (30_001, 5, 8), (65_537, 37, 20), (999, 3, 24), (40, 2, 4)])
def test_search_matches_oracle(cuda, dtype, n, nq, k):
    from rdm_b200.knn import B200Searcher
    db = _db(n, 512, dtype, seed=n)
    rng = np.random.default_rng(n + 1)
    rows = rng.integers(0, n, size=nq)
    rows[0] = 7 if n > 5000 else rows[0]
    q = db[rows].astype(np.float32)
    q[nq // 2:] = rng.standard_normal((nq - nq // 2, 512)).astype(np.float32)   # half DB rows (ddpm.py:897), half free queries
    qh = oknn.normalize_queries(q)
    s = B200Searcher(db, device=cuda)
    inv_ref = oknn.inv_norms(db)
    assert np.array_equal(s.inv_norms().cpu().numpy().view(np.uint32), inv_ref.view(np.uint32)), "inverse norms must be bit-exact"
    idx, dist, sc = s.search_device(torch.from_numpy(qh).to(cuda), k, return_scores=True)
    ri, rd, rs = oknn.search(db, qh, k, inv=inv_ref, return_scores=True)
    kk = min(k, n)
    assert np.array_equal(idx.cpu().numpy()[:, :kk], ri[:, :kk]), "kNN indices must be bit-exact"
    assert np.array_equal(sc.cpu().numpy()[:, :kk].view(np.uint64), rs[:, :kk].view(np.uint64)), "fp64 scores must be bit-exact"
    assert np.array_equal(dist.cpu().numpy()[:, :kk], rd[:, :kk])
    if n > 5000 and nq >= 2:
        assert list(idx[0, :4].cpu().numpy()) == [7, 1234, 4321, n - 1]     # query 0 is DB row 7; its exact duplicates tie -> index order


def test_scann_shaped_api_and_gather(cuda):
    from rdm_b200.knn import B200Searcher
    db = _db(20_000, 512, np.float16, seed=3)
    s = B200Searcher(db, device=cuda)
    q = oknn.normalize_queries(db[[11, 12, 13]].astype(np.float32))
    nns, distances = s.search_batched(q, final_num_neighbors=4)
    assert nns.shape == (3, 4) and distances.dtype == np.float32 and list(nns[:, 0]) == [11, 12, 13]
    got = s.gather_device(torch.from_numpy(nns).to(cuda)).cpu().numpy()
    assert got.dtype == np.float32 and np.array_equal(got, db[nns].astype(np.float32))      # raw rows, un-normalised (F7)


def test_shards_merge_to_the_unsharded_result(cuda):
    from rdm_b200.knn import B200Searcher, merge_device
    db = _db(80_000, 512, np.float16, seed=5)
    qh = oknn.normalize_queries(np.random.default_rng(6).standard_normal((9, 512)))
    qd = torch.from_numpy(qh).to(cuda)
    full = B200Searcher(db, device=cuda).search_device(qd, 8, return_scores=True)
    parts_i, parts_s = [], []
    for base in range(0, 80_000, 20_000):
        sh = B200Searcher(db[base:base + 20_000], device=cuda, idx_base=base)
        i, _, sc = sh.search_device(qd, 8, return_scores=True)
        parts_i.append(i); parts_s.append(sc)
    mi, md, ms = merge_device(torch.stack(parts_i), torch.stack(parts_s), 8)
    assert torch.equal(mi, full[0]) and torch.equal(ms, full[2]) and torch.equal(md, full[1])


def test_unrepresentative_sample_takes_the_fallback_path(cuda):
    """The main scan's threshold comes from a strided sample of row groups.  Hide 5000 near-duplicates of the query in
    rows the sample never visits: the candidate buffer (2048) overflows and the device-side fallback must still be exact."""
    from rdm_b200.knn import B200Searcher
    n = 300_000
    rng = np.random.default_rng(9)
    db = rng.standard_normal((n, 512)).astype(np.float32)
    qv = rng.standard_normal(512).astype(np.float32)
    stride = max(1, min(16, ((n + 7) // 8) // 8192))
    assert stride >= 2
    hidden = np.array([r for r in rng.permutation(n) if (r // 8) % stride != 0][:5000])
    db[hidden] = qv[None] + 0.05 * rng.standard_normal((5000, 512)).astype(np.float32)
    db = db.astype(np.float16)
    qh = oknn.normalize_queries(np.stack([qv, rng.standard_normal(512).astype(np.float32)]))
    idx, _, sc = B200Searcher(db, device=cuda).search_device(torch.from_numpy(qh).to(cuda), 8, return_scores=True)
    ri, _, rs = oknn.search(db, qh, 8, return_scores=True)
    assert np.array_equal(idx.cpu().numpy(), ri) and np.array_equal(sc.cpu().numpy().view(np.uint64), rs.view(np.uint64))
    assert set(ri[0]) <= set(hidden)


def test_many_exact_duplicates_tie_break(cuda):
    """4000 identical rows (e.g. placeholder images): ties must resolve to the lowest indices without overflowing anything."""
    from rdm_b200.knn import B200Searcher
    rng = np.random.default_rng(10)
    db = rng.standard_normal((120_000, 512)).astype(np.float16)
    dup = np.sort(rng.permutation(120_000)[:4000])
    db[dup] = db[dup[0]]
    qh = oknn.normalize_queries(db[dup[:1]].astype(np.float32))
    idx, _ = B200Searcher(db, device=cuda).search_device(torch.from_numpy(qh).to(cuda), 20)
    assert list(idx[0].cpu().numpy()) == list(dup[:20])


def test_other_row_widths(cuda):
    from rdm_b200.knn import B200Searcher
    for d in (256, 768, 1024):
        db = _db(9000, d, np.float16, seed=d)
        qh = oknn.normalize_queries(np.random.default_rng(d).standard_normal((4, d)))
        idx, _ = B200Searcher(db, device=cuda).search_device(torch.from_numpy(qh).to(cuda), 4)
        assert np.array_equal(idx.cpu().numpy(), oknn.search(db, qh, 4)[0])


def test_bad_arguments_fail_loudly(cuda):
    from rdm_b200.knn import B200Searcher
    s = B200Searcher(_db(100, 512, np.float32, seed=1, dup=False), device=cuda)
    with pytest.raises(RuntimeError):
        s.search_device(torch.zeros(1, 512, device=cuda), 25)          # k > RDM_KNN_MAX_K
    with pytest.raises(RuntimeError):
        B200Searcher(np.zeros((10, 100), np.float32), device=cuda)     # unsupported row width


@pytest.mark.parametrize("d", [256, 512, 768, 1024])
def test_query_normalisation_inside_the_library_is_bit_identical_to_numpy(cuda, d):
    """rdm_knn_normalize / rdm_knn_search_raw: the reference's `q / np.linalg.norm(q, axis=1)[:, np.newaxis]` (ddpm.py:907) on the device."""
    from rdm_b200.knn import B200Searcher, normalize_device
    rng = np.random.default_rng(d)
    q = (rng.standard_normal((19, d)) * rng.uniform(0.01, 30.0, size=(19, 1))).astype(np.float32)
    want = q / np.linalg.norm(q, axis=1)[:, np.newaxis]
    got = normalize_device(torch.from_numpy(q).to(cuda)).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(oknn.normalize_queries_restated(q).view(np.uint32), want.view(np.uint32))
    db = _db(20_000, d, np.float16, seed=d + 1)
    s = B200Searcher(db, device=cuda)
    a = s.search_raw_device(torch.from_numpy(q).to(cuda), 6, return_scores=True)
    b = s.search_device(torch.from_numpy(want).to(cuda), 6, return_scores=True)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    ri, _, rs = oknn.search(db, want, 6, return_scores=True)
    assert np.array_equal(a[0].cpu().numpy(), ri) and np.array_equal(a[2].cpu().numpy().view(np.uint64), rs.view(np.uint64))


@pytest.mark.parametrize("nq,k,dtype", [(37, 20, np.float16), (64, 8, np.float16), (100, 20, np.float16), (130, 4, np.float16),
                                        (5, 8, np.float32), (37, 20, np.float32), (100, 4, np.float32)])
def test_wide_hi_only_passes_stay_exact_on_near_duplicate_clusters(cuda, nq, k, dtype):
    """Passes of more than 16 queries over a database large enough for the fused scan keep only the fp16 hi rows of the queries (up to 128
    queries per pass): the scan score is then off by up to 2^-11, far more than the spacing inside a cluster of near-duplicate rows.  The
    wider score slack of such a pass (thresholds and the select cut, knn_tc.cu: HI_SLACK) must keep every true neighbour among the rows
    that are re-ranked exactly: indices and fp64 scores bit-identical to the oracle, on clustered and on plain queries.  fp32 databases take
    the same fused scan on kind::tf32 MMAs (rows AND queries cut to 10 mantissa bits, slack 4.2e-3, passes of 64)."""
    from oracle import knn as oknn
    from rdm_b200.knn import B200Searcher
    n = 200_000
    rng = np.random.default_rng(1000 + nq)
    db = (rng.standard_normal((n, 512)) * rng.uniform(0.5, 8, (n, 1))).astype(dtype)
    db[5000:5600] = db[5000] + (rng.standard_normal((600, 512)) * 2e-3).astype(dtype)                # 600 rows within ~1e-6 of each other in cosine
    db[90_000:90_050] = db[90_000]                                                                      # exact duplicates: ties broken by index
    q = rng.standard_normal((nq, 512)).astype(np.float32)
    q[0], q[nq // 2], q[nq - 1] = db[5000].astype(np.float32), db[5300].astype(np.float32) * 3.0, db[90_010].astype(np.float32)
    qh = oknn.normalize_queries(q)
    idx, dist, sc = B200Searcher(db, device=cuda).search_device(torch.from_numpy(qh).to(cuda), k, return_scores=True)
    wi, wd, ws = oknn.search(db, qh, k, return_scores=True)
    assert set(wi[0]) <= set(range(5000, 5600)) and list(wi[nq - 1][:k]) == list(range(90_000, 90_000 + k))
    assert np.array_equal(idx.cpu().numpy(), wi)
    assert np.array_equal(sc.cpu().numpy().view(np.int64), ws.view(np.int64))


@pytest.mark.parametrize("n,nq", [(30_000, 100), (30_000, 129), (90_000, 70)])
def test_more_than_64_queries_on_databases_with_and_without_the_fused_scan(cuda, n, nq):
    """Passes of up to 128 queries exist only on the fused scan; a database too small for it (fewer than 4 tiles per SM) takes the
    three-kernel path, which splits a pass of more than 64 queries in two.  Both entry points (normalised and raw queries) against the
    oracle, bit-exact."""
    from oracle import knn as oknn
    from rdm_b200.knn import B200Searcher
    rng = np.random.default_rng(n + nq)
    db = (rng.standard_normal((n, 512)) * rng.uniform(0.5, 8, (n, 1))).astype(np.float16)
    q = rng.standard_normal((nq, 512)).astype(np.float32) * 5.0
    q[3] = db[777].astype(np.float32)
    qh = oknn.normalize_queries(q)
    s = B200Searcher(db, device=cuda)
    wi, wd, ws = oknn.search(db, qh, 5, return_scores=True)
    for idx, dist, sc in (s.search_device(torch.from_numpy(qh).to(cuda), 5, return_scores=True), s.search_raw_device(torch.from_numpy(q).to(cuda), 5, return_scores=True)):
        assert np.array_equal(idx.cpu().numpy(), wi) and wi[3, 0] == 777
        assert np.array_equal(sc.cpu().numpy().view(np.int64), ws.view(np.int64))
