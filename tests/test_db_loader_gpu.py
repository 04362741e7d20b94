"""GPU: the direct-to-HBM database loader (rdm_b200/db_loader.py load_rows_to_device, DatasetBuilder(direct_to_hbm=True)):
rows go from the multi-part .npz database to the searcher's device buffer through pinned staging chunks; a row-sharded load + search
equals the unsharded search; the builder's result dictionary equals the one built from the host arrays."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _write_db(tmp_path, sizes=(5000, 1, 7003, 2999), d=512, dtype=np.float16):
    rng = np.random.default_rng(0)
    parts, start = [], 0
    for i, n in enumerate(sizes):
        emb = (rng.standard_normal((n, d)) * rng.uniform(0.5, 8, (n, 1))).astype(dtype)
        np.savez(tmp_path / f"part_{i:03d}.npz", embedding=emb, img_id=np.arange(start, start + n), patch_coords=np.zeros((n, 4), np.int32))
        parts.append(emb); start += n
    return np.concatenate(parts)


@pytest.mark.parametrize("dtype", [np.float16, np.float32])
def test_rows_reach_hbm_unchanged_for_any_range_and_chunk_size(cuda, tmp_path, dtype):
    from rdm_b200 import db_loader
    full = _write_db(tmp_path, dtype=dtype)
    n = len(full)
    for lo, hi, chunk in ((0, n, 1 << 20), (4999, 5003, 4096), (5001, n, 3 * 512 * 2), (123, 9000, 1 << 16)):
        t, stats = db_loader.load_rows_to_device(str(tmp_path), lo, hi, cuda, chunk_bytes=chunk)
        assert t.dtype == (torch.float16 if dtype == np.float16 else torch.float32) and t.shape == (hi - lo, 512)
        assert np.array_equal(t.cpu().numpy(), full[lo:hi]) and stats["rows"] == hi - lo and stats["n_total"] == n and stats["gb_per_s"] > 0


def test_sharded_load_and_search_equals_the_unsharded_search(cuda, tmp_path):
    from oracle import knn as oknn
    from rdm_b200 import db_loader
    from rdm_b200.knn import B200Searcher, merge_device, shard_range
    full = _write_db(tmp_path)
    n = len(full)
    q = torch.from_numpy(np.concatenate([full[[7, 9000]].astype(np.float32), np.random.default_rng(5).standard_normal((5, 512)).astype(np.float32)])).to(cuda)
    whole, _ = db_loader.load_rows_to_device(str(tmp_path), 0, n, cuda)
    want = B200Searcher(whole, device=cuda).search_raw_device(q, 6, return_scores=True)
    for world in (2, 3):
        parts = []
        for r in range(world):
            lo, hi = shard_range(n, r, world)
            rows, _ = db_loader.load_rows_to_device(str(tmp_path), lo, hi, cuda, chunk_bytes=1 << 18)
            parts.append(B200Searcher(rows, device=cuda, idx_base=lo).search_raw_device(q, 6, return_scores=True))
        idx, dist, sc = merge_device(torch.stack([p[0] for p in parts]), torch.stack([p[2] for p in parts]), 6)
        assert torch.equal(idx, want[0]) and torch.equal(sc, want[2]) and torch.equal(dist, want[1])
    ref_i, _ = oknn.search(full, oknn.normalize_queries(q.cpu().numpy()), 6)
    assert np.array_equal(want[0].cpu().numpy(), ref_i)


def test_dataset_builder_direct_to_hbm_matches_the_host_path(cuda, tmp_path):
    import rdm  # noqa: F401
    from rdm.data.retrieval_dataset.dsetbuilder import DatasetBuilder
    full = _write_db(tmp_path)
    kw = dict(retriever_config=None, saved_embeddings=str(tmp_path), load_patch_dataset=False, gpu=True, k=5)
    host = DatasetBuilder(**kw)
    host.train_searcher()
    direct = DatasetBuilder(direct_to_hbm=True, **kw)
    assert direct.searcher is not None and direct.load_stats["rows"] == len(full) and len(direct.data_pool["embedding"]) == len(full)
    assert direct.num_rows == len(full) and direct.data_pool["embedding"].shape == full.shape
    queries = np.concatenate([full[[17, 12000]].astype(np.float32), np.random.default_rng(2).standard_normal((3, 512)).astype(np.float32)])
    a, b = host.search_k_nearest(queries, k=5, query_embedded=True), direct.search_k_nearest(queries, k=5, query_embedded=True)
    assert np.array_equal(a["nns"], b["nns"]) and np.array_equal(a["distances"], b["distances"]) and np.array_equal(a["img_ids"], b["img_ids"])
    assert np.array_equal(np.asarray(a["embeddings"], dtype=np.float32), b["embeddings"])
