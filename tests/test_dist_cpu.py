"""CPU, world_size 2, gloo: the row-sharded database exchange -- shard ranges, idx_base, one all_gather
of (idx, fp64 score), merge by (score desc, idx asc).  The local scan is played by the oracle so no GPU is needed; the
CUDA kernels behind the same interfaces are covered by tests/test_knn_gpu.py::test_shards_merge_to_the_unsharded_result."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

from conftest import PKG, ROOT


class OracleLocalSearcher:
    """CPU stand-in with the B200Searcher device-level interface."""

    def __init__(self, db, idx_base):
        from oracle import knn as oknn
        self.db, self.base, self.oknn, self.device, self.d = db, idx_base, oknn, torch.device("cpu"), db.shape[1]
        self.inv = oknn.inv_norms(db)

    def search_device(self, q_hat, k, return_scores=False):
        i, d, s = self.oknn.search(self.db, q_hat.numpy(), k, inv=self.inv, idx_base=self.base, return_scores=True)
        out = (torch.from_numpy(i), torch.from_numpy(d), torch.from_numpy(s))
        return out if return_scores else out[:2]

    def gather_device(self, idx):
        rows = idx.numpy() - self.base
        ok = (rows >= 0) & (rows < self.db.shape[0])
        out = np.zeros(idx.shape + (self.d,), np.float32)
        out[ok] = self.db[rows[ok]].astype(np.float32)
        return torch.from_numpy(out)


def cpu_merge(idx_parts, score_parts, k):
    from oracle import knn as oknn
    parts = [(idx_parts[p].numpy(), score_parts[p].numpy()) for p in range(idx_parts.shape[0])]
    i, s = oknn.merge_shards(parts, k)
    return torch.from_numpy(i), torch.from_numpy(s.astype(np.float32)), torch.from_numpy(s)


def _worker(rank, world, port, q):
    import sys
    sys.path[:0] = [ROOT, PKG]
    import torch.distributed as dist
    from oracle import knn as oknn
    from rdm_b200.knn import ShardedSearcher, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)                              # same DB / queries on every rank
    db = rng.standard_normal((5001, 512)).astype(np.float16)
    db[4000] = db[3]                                            # a duplicate that lives in the OTHER shard: tie -> lowest global index
    qh = oknn.normalize_queries(np.concatenate([db[[3, 4990]].astype(np.float32), rng.standard_normal((3, 512)).astype(np.float32)]))
    lo, hi = shard_range(len(db), rank, world)
    full_i, full_d = oknn.search(db, qh, 6)
    ok = list(full_i[0, :2]) == [3, 4000]
    # (a) the same queries on every rank, declared: one packed all_gather, all_reduce of the owners' rows
    s = ShardedSearcher(OracleLocalSearcher(db[lo:hi], lo), merge_fn=cpu_merge, same_queries=True)
    idx, dist_ = s.search_device(torch.from_numpy(qh), 6)
    ctx = s.gather_device(idx)
    ok = ok and np.array_equal(idx.numpy(), full_i) and np.array_equal(dist_.numpy(), full_d) and np.array_equal(ctx.numpy(), db[full_i].astype(np.float32))
    # (b) the default contract: every rank brings its OWN queries (same count) -- all_gather of the queries, all_to_all of the packed
    #     lists, reduce_scatter of the owners' rows; each rank must get exactly the unsharded answer for ITS queries
    s2 = ShardedSearcher(OracleLocalSearcher(db[lo:hi], lo), merge_fn=cpu_merge)
    mine = oknn.normalize_queries(np.concatenate([db[[3 + 4000 * rank, 17 + rank]].astype(np.float32),
                                                  np.random.default_rng(100 + rank).standard_normal((3, 512)).astype(np.float32)]))
    idx2, dist2 = s2.search_device(torch.from_numpy(mine), 6)
    ctx2 = s2.gather_device(idx2)
    want_i, want_d = oknn.search(db, mine, 6)
    ok = ok and np.array_equal(idx2.numpy(), want_i) and np.array_equal(dist2.numpy(), want_d) and np.array_equal(ctx2.numpy(), db[want_i].astype(np.float32))
    ok = ok and ctx2.shape == (5, 6, 512)
    # (c) unequal query counts are reported, not mis-merged
    try:
        s2.search_device(torch.from_numpy(mine[:4 + rank]), 6)
        ok = False
    except RuntimeError as e:
        ok = ok and "same number of queries" in str(e)
    q.put((rank, bool(ok), (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_search_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] == (0, 2500) and res[1][2] == (2500, 5001)


def test_shard_ranges_cover_the_database():
    from rdm_b200.knn import shard_range
    for n in (1, 7, 1_281_167, 20_927_907):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, i, w) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))


def _builder_worker(rank, world, port, q, dbdir):
    """DatasetBuilder(shard=True) under a real process group: each rank reads only its rows of the multi-part database, the sharded
    searcher exchanges (idx, score) lists, search_k_nearest returns the same dictionary as an unsharded search on every rank."""
    import sys
    sys.path[:0] = [ROOT, PKG]
    import torch.distributed as dist
    from oracle import knn as oknn
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rdm  # noqa: F401
    import rdm.data.retrieval_dataset.dsetbuilder as dsb
    import rdm_b200.knn as bknn
    dsb.B200Searcher = lambda emb, device=None, idx_base=0: OracleLocalSearcher(np.ascontiguousarray(emb), idx_base)      # stand-in for the device scan
    bknn.merge_device = cpu_merge                                                                                        # stand-in for rdm_knn_merge
    full = np.concatenate([np.load(os.path.join(dbdir, f))["embedding"] for f in sorted(os.listdir(dbdir))])
    b = dsb.DatasetBuilder(retriever_config=None, saved_embeddings=dbdir, load_patch_dataset=False, gpu=False, shard=True, k=5)
    lo, hi = bknn.shard_range(len(full), rank, world)
    ok = b.data_pool["embedding"].shape[0] == hi - lo and b._row_base == lo
    b.train_searcher()
    rng = np.random.default_rng(3)
    queries = np.concatenate([full[[7, 2600]].astype(np.float32), rng.standard_normal((2, 512)).astype(np.float32)])
    out = b.search_k_nearest(queries, k=5, query_embedded=True)
    want_i, want_d = oknn.search(full, oknn.normalize_queries(queries), 5)
    ok = ok and np.array_equal(out["nns"], want_i) and np.array_equal(out["distances"], want_d)
    ok = ok and np.array_equal(out["embeddings"], full[want_i].astype(np.float32)) and np.array_equal(out["img_ids"], want_i)
    # pseudo-queries are drawn from the WHOLE database on every rank (ddpm.py:867), not from the rows this rank happens to hold
    ok = ok and b.num_rows == len(full) and len(b.data_pool["embedding"]) < len(full)
    from rdm.models.diffusion.ddpm import MinimalRETRODiffusion
    from rdm.models.autoregression.transformer import LatentImageRETRO
    import types
    for cls in (MinimalRETRODiffusion, LatentImageRETRO):
        host = types.SimpleNamespace(use_memory=False, retriever=b)
        np.random.seed(5)
        qids = cls.get_qids(host, 100, 4000)
        np.random.seed(5)
        ok = ok and np.array_equal(qids, np.random.choice(len(full), size=4000)) and int(qids.max()) >= hi - lo
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_dataset_builder_world2_gloo(tmp_path):
    rng = np.random.default_rng(1)
    start = 0
    for i, n in enumerate((1200, 900, 1501)):                  # parts that do not align with the shard boundary
        np.savez(tmp_path / f"part_{i}.npz", embedding=rng.standard_normal((n, 512)).astype(np.float16), img_id=np.arange(start, start + n),
                 patch_coords=np.zeros((n, 4), np.int32))
        start += n
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_builder_worker, args=(r, 2, port, q, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
