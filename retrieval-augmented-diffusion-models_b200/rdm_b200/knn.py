"""Host side of the retrieval sink: a drop-in for the ScaNN searcher object the reference builds in
``rdm/data/retrieval_dataset/dsetbuilder.py:534-619`` and queries through
``searcher.search_batched(q_hat, final_num_neighbors=k) -> (indices, distances)``
(``dsetbuilder.py:490``, ``ddpm.py:298,906-908``, ``transformer.py:327-329``, ``base.py:81-83``).

All arithmetic happens in librdm_b200 (csrc/knn.cu); this file only moves pointers.
"""
import ctypes

import numpy as np
import torch

from . import _lib

MAX_K = 24
_DT = {torch.float32: 0, torch.float16: 1, np.dtype("float32"): 0, np.dtype("float16"): 1}


class B200Searcher:
    """Exact cosine top-k over a (shard of a) CLIP database resident in HBM.

    ``embedding``: RAW (un-normalised) rows, fp16 or fp32, ``[n, d]`` -- a numpy array (copied to the
    device once) or a CUDA tensor (used in place, kept alive by this object).  ``idx_base`` is the global
    index of row 0 when the database is row-sharded across ranks.
    """

    def __init__(self, embedding, device=None, idx_base=0):
        L = _lib.lib()
        if device is None:
            device = embedding.device if isinstance(embedding, torch.Tensor) and embedding.is_cuda else "cuda"
        self.device = _lib.resolve_device(device)
        if isinstance(embedding, np.ndarray):
            if embedding.dtype not in (np.float16, np.float32):
                embedding = embedding.astype(np.float32)
            embedding = torch.from_numpy(np.ascontiguousarray(embedding))
        if embedding.dtype not in (torch.float16, torch.float32):
            embedding = embedding.float()
        assert embedding.ndim == 2
        self._db = embedding.contiguous().to(self.device, non_blocking=False)     # one H2D copy; stays resident
        self.n, self.d = self._db.shape
        self.idx_base = int(idx_base)
        self._h = ctypes.c_void_p()
        _lib.check(L.rdm_knn_create(ctypes.byref(self._h), _lib.ptr(self._db), self.n, self.d, _DT[self._db.dtype], 1,
                                    self.idx_base, int(self.device.index or 0)), "rdm_knn_create")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().rdm_knn_destroy(h)
            except Exception:
                pass

    # ---- device-level API (no host round trip) -------------------------------------------------------
    def search_device(self, q_hat, k, return_scores=False):
        """q_hat: CUDA float32 [nq, d], already L2-normalised -> (idx int64 [nq,k], dist float32 [nq,k][, score float64])."""
        assert q_hat.device == self.device and q_hat.dtype == torch.float32 and q_hat.shape[1] == self.d, "queries must be float32 [nq, d] on the searcher's device"
        q_hat = q_hat.contiguous()
        nq = q_hat.shape[0]
        idx = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        dist = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        sc = torch.empty((nq, k), dtype=torch.float64, device=self.device) if return_scores else None
        with _lib.device_ctx(self.device):
            _lib.check(_lib.lib().rdm_knn_search(self._h, _lib.ptr(q_hat), nq, k, _lib.ptr(idx), _lib.ptr(dist), _lib.ptr(sc),
                                                  _lib.stream_ptr(self.device)), "rdm_knn_search")
        return (idx, dist, sc) if return_scores else (idx, dist)

    def search_raw_device(self, q_raw, k, return_scores=False):
        """q_raw: CUDA float32 [nq, d] RAW query embeddings; the reference's NumPy normalisation (ddpm.py:907) runs inside the library,
        bit-identical to ``q / np.linalg.norm(q, axis=1)[:, None]`` -> same outputs as ``search_device``."""
        assert q_raw.device == self.device and q_raw.dtype == torch.float32 and q_raw.shape[1] == self.d, "queries must be float32 [nq, d] on the searcher's device"
        q_raw = q_raw.contiguous()
        nq = q_raw.shape[0]
        idx = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        dist = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        sc = torch.empty((nq, k), dtype=torch.float64, device=self.device) if return_scores else None
        with _lib.device_ctx(self.device):
            _lib.check(_lib.lib().rdm_knn_search_raw(self._h, _lib.ptr(q_raw), nq, k, _lib.ptr(idx), _lib.ptr(dist), _lib.ptr(sc),
                                                      _lib.stream_ptr(self.device)), "rdm_knn_search_raw")
        return (idx, dist, sc) if return_scores else (idx, dist)

    def gather_device(self, idx):
        """``data_pool['embedding'][nns]`` as float32 on the device (ddpm.py:921): idx int64 [...] -> [..., d]."""
        flat = idx.reshape(-1).to(self.device, torch.int64).contiguous()
        out = torch.empty((flat.numel(), self.d), dtype=torch.float32, device=self.device)
        with _lib.device_ctx(self.device):
            _lib.check(_lib.lib().rdm_knn_gather(self._h, _lib.ptr(flat), flat.numel(), _lib.ptr(out), _lib.stream_ptr(self.device)),
                       "rdm_knn_gather")
        return out.reshape(*idx.shape, self.d)

    def inv_norms(self):
        out = torch.empty(self.n, dtype=torch.float32, device=self.device)
        with _lib.device_ctx(self.device):
            _lib.check(_lib.lib().rdm_knn_get_inv_norms(self._h, _lib.ptr(out), _lib.stream_ptr(self.device)), "rdm_knn_get_inv_norms")
        return out

    # ---- the ScaNN-shaped API the reference calls ----------------------------------------------------
    def search_batched(self, queries, final_num_neighbors=None, **_ignored):
        """numpy [nq, d] (already normalised by the caller) -> (uint32-like indices [nq,k], float32 distances [nq,k])."""
        k = int(final_num_neighbors)
        q = torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32)).to(self.device)
        idx, dist = self.search_device(q, k)
        return idx.cpu().numpy(), dist.cpu().numpy()

    def search(self, query, final_num_neighbors=None, **kw):
        i, d = self.search_batched(np.asarray(query)[None], final_num_neighbors, **kw)
        return i[0], d[0]


def normalize_device(q):
    """``q / np.linalg.norm(q, axis=1)[:, np.newaxis]`` (ddpm.py:297,907) for CUDA float32 rows, bit-identical to NumPy (csrc/knn.cu)."""
    dev = _lib.resolve_device(q.device)                  # raises unless CUDA: there is no CPU path
    assert q.dtype == torch.float32 and q.ndim == 2
    q = q.contiguous()
    out = torch.empty_like(q)
    with _lib.device_ctx(dev):
        _lib.check(_lib.lib().rdm_knn_normalize(_lib.ptr(q), q.shape[0], q.shape[1], _lib.ptr(out), int(dev.index or 0),
                                                 _lib.stream_ptr(dev)), "rdm_knn_normalize")
    return out


def search_raw(searcher, q_raw, k):
    """Retrieval front end of the host mirrors: RAW query embeddings (CUDA float32 [nq, d]) -> (idx, dist).  This package's searchers
    normalise inside the library (``rdm_knn_search_raw``: no eager tensor arithmetic between ``get_qids`` and the scan).  A foreign
    searcher object assigned to ``retriever.searcher`` (the reference allows any object with the ScaNN call shape) gets queries
    normalised by the reference's own NumPy statement (``ddpm.py:907``)."""
    if hasattr(searcher, "search_raw_device"):
        return searcher.search_raw_device(q_raw.float().contiguous(), k)
    q = q_raw.detach().float().cpu().numpy()
    qh = torch.from_numpy(q / np.linalg.norm(q, axis=1)[:, np.newaxis]).to(q_raw.device)
    return searcher.search_device(qh.contiguous(), k)[:2]


def merge_device(idx_parts, score_parts, k):
    """[parts, nq, k] int64 / float64 (CUDA) -> global (idx [nq,k], dist float32, score float64)."""
    parts, nq, kk = idx_parts.shape
    assert kk == k
    dev = idx_parts.device
    idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
    dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
    sc = torch.empty((nq, k), dtype=torch.float64, device=dev)
    with _lib.device_ctx(dev):
        _lib.check(_lib.lib().rdm_knn_merge(_lib.ptr(idx_parts.contiguous()), _lib.ptr(score_parts.contiguous()), parts, nq, k,
                                             _lib.ptr(idx), _lib.ptr(dist), _lib.ptr(sc), int(dev.index or 0), _lib.stream_ptr(dev)), "rdm_knn_merge")
    return idx, dist, sc


def shard_range(n, rank, world):
    """Rows [lo, hi) owned by `rank` of `world` for an n-row database (contiguous, balanced; DatasetBuilder.train_searcher)."""
    return (n * rank) // world, (n * (rank + 1)) // world


class ShardedSearcher:
    """Row-sharded database: rank r owns rows [base_r, base_r + n_r).  Works with any initialised
    torch.distributed backend (NCCL on the box, gloo in CI).

    Contract: every rank calls with the SAME NUMBER of queries; the query ROWS may differ per rank (images sharded by batch index, each
    rank drawing its own pseudo-queries) -- the exchange below is correct either way:
      1. all_gather of the query rows: every rank holds the union [G*nq, d]                              (G*nq*d*4 B)
      2. exact local search of the union over the rank's own rows (global indices, fp64 scores)
      3. ONE all_to_all of the packed (idx | score bits) int64 lists: rank s receives, from every shard, the lists of ITS queries
      4. on-device merge by (score desc, index asc): bit-identical to the unsharded result for any shard count
    `gather_device` fetches neighbour rows from their owners: all_gather of the index lists, every rank gathers the rows it owns (zeros
    elsewhere) in the database dtype, one reduce_scatter hands each rank the rows of its own queries (exact: one value plus zeros).
    `same_queries=True` (callers that guarantee identical queries on every rank, e.g. one script replicated per rank over a sharded
    DatasetBuilder): step 1 is skipped and step 3 is one all_gather, every rank ends with the full result."""

    def __init__(self, local, group=None, merge_fn=None, same_queries=False, validate=True):
        import torch.distributed as dist
        self.local, self.group, self.dist, self.same_queries, self.validate = local, group, dist, same_queries, validate
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.merge = merge_fn or merge_device          # injectable so the exchange logic is testable on CPU/gloo

    @property
    def device(self):
        return self.local.device

    def _check_equal_counts(self, nq):
        """One 8-byte all_gather per call (`validate=False` switches it off for callers that own the batching): unequal counts would
        mis-shape -- or hang -- the exchange that follows."""
        if not self.validate:
            return
        mine = torch.tensor([nq], dtype=torch.int64, device=self.local.device)
        out = torch.empty(self.world, dtype=torch.int64, device=self.local.device)
        self.dist.all_gather_into_tensor(out, mine, group=self.group)
        counts = out.tolist()
        if any(c != nq for c in counts):
            raise RuntimeError(f"ShardedSearcher: every rank must pass the same number of queries, got {counts} (pad the last batch)")

    @staticmethod
    def _pack(idx, sc):
        return torch.cat([idx, sc.view(torch.int64)], dim=1).contiguous()              # [Q, 2k] int64: indices | fp64 score bit patterns

    def search_device(self, q_hat, k):
        G = self.world
        if G == 1:
            idx, _, sc = self.local.search_device(q_hat, k, return_scores=True)
            return self.merge(idx[None], sc[None], k)[:2]
        nq = q_hat.shape[0]
        self._check_equal_counts(nq)
        q_hat = q_hat.contiguous()
        if self.same_queries:
            idx, _, sc = self.local.search_device(q_hat, k, return_scores=True)
            packed = self._pack(idx, sc)
            allp = torch.empty((G,) + packed.shape, dtype=torch.int64, device=packed.device)
            self.dist.all_gather_into_tensor(allp.view(-1), packed.view(-1), group=self.group)
        else:
            q_all = torch.empty((G * nq, q_hat.shape[1]), dtype=q_hat.dtype, device=q_hat.device)
            self.dist.all_gather_into_tensor(q_all.view(-1), q_hat.view(-1), group=self.group)
            idx, _, sc = self.local.search_device(q_all, k, return_scores=True)          # rows ordered by owner rank of the query
            packed = self._pack(idx, sc)
            allp = torch.empty_like(packed)
            self.dist.all_to_all_single(allp.view(-1), packed.view(-1), group=self.group)
            allp = allp.view(G, nq, 2 * k)                                                # [shard, my query, idx | score]
        return self.merge(allp[:, :, :k].contiguous(), allp[:, :, k:].contiguous().view(torch.float64), k)[:2]

    def search_raw_device(self, q_raw, k):
        """Raw queries: normalised once (bit-identical to NumPy, csrc/knn.cu) and searched on every shard."""
        return self.search_device(normalize_device(q_raw), k)

    def gather_device(self, idx):
        G = self.world
        if G == 1:
            return self.local.gather_device(idx)
        if self.same_queries:
            out = self.local.gather_device(idx)          # rows outside the local shard come back as zeros
            self.dist.all_reduce(out, group=self.group)
            return out
        flat = idx.reshape(-1).to(self.local.device, torch.int64).contiguous()
        self._check_equal_counts(flat.numel())
        idx_all = torch.empty(G * flat.numel(), dtype=torch.int64, device=flat.device)
        self.dist.all_gather_into_tensor(idx_all, flat, group=self.group)
        rows = self.local.gather_device(idx_all)                                          # [G*n, d] fp32, zeros for rows other ranks own
        db = getattr(self.local, "_db", None)
        if db is not None and db.dtype == torch.float16:
            rows = rows.half()                                                            # exact (the rows ARE fp16 values): halves the bytes on the wire
        mine = torch.empty((flat.numel(), rows.shape[1]), dtype=rows.dtype, device=rows.device)
        self.dist.reduce_scatter_tensor(mine.view(-1), rows.contiguous().view(-1), group=self.group)      # owner's value + zeros: exact
        return mine.float().reshape(*idx.shape, rows.shape[1])

    def search_batched(self, queries, final_num_neighbors=None, **_):
        q = torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32)).to(self.local.device)
        idx, dist = self.search_device(q, int(final_num_neighbors))
        return idx.cpu().numpy(), dist.cpu().numpy()
