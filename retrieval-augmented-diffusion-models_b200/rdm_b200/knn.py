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
    """Row-sharded database: rank r owns rows [base_r, base_r + n_r).  Every rank passes the SAME queries;
    one all_gather of the per-shard exact (index, fp64 score) lists, then an on-device merge by
    (score desc, index asc) -- identical results on every rank, independent of the number of shards
    (SURVEY.md section 8e).  Works with any initialised torch.distributed backend (NCCL on the box, gloo in CI)."""

    def __init__(self, local, group=None, merge_fn=None):
        import torch.distributed as dist
        self.local, self.group, self.dist = local, group, dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.merge = merge_fn or merge_device          # injectable so the exchange logic is testable on CPU/gloo

    def search_device(self, q_hat, k):
        idx, _, sc = self.local.search_device(q_hat, k, return_scores=True)
        if self.world == 1:
            return self.merge(idx[None], sc[None], k)[:2]
        idx_all = [torch.empty_like(idx) for _ in range(self.world)]
        sc_all = [torch.empty_like(sc) for _ in range(self.world)]
        self.dist.all_gather(idx_all, idx, group=self.group)
        self.dist.all_gather(sc_all, sc, group=self.group)
        return self.merge(torch.stack(idx_all), torch.stack(sc_all), k)[:2]

    def gather_device(self, idx):
        out = self.local.gather_device(idx)          # rows outside the local shard come back as zeros
        if self.world > 1:
            self.dist.all_reduce(out, group=self.group)
        return out

    def search_batched(self, queries, final_num_neighbors=None, **_):
        q = torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32)).to(self.local.device)
        idx, dist = self.search_device(q, int(final_num_neighbors))
        return idx.cpu().numpy(), dist.cpu().numpy()
