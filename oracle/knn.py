"""Oracle kNN: exact cosine top-k.  TEST INFRASTRUCTURE (see oracle/__init__.py).

``search`` drives the C restatement ``oracle/knn_ref.c`` (the definition: sequential fp64
dot products, ties -> lowest index).  ``search_numpy`` is an independent vectorised numpy
statement of the same definition (``dsetbuilder.py:574,487-490`` semantics) used to
cross-check the C code; it may differ from it in the last fp64 bits (pairwise summation),
never in the ranking of well-separated scores.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libknn_ref.so")
_lib = None


def build(force=False):
    """gcc the C oracle into oracle/_build/ (git-ignored; travels to the GPU box with gpurun)."""
    src = os.path.join(_HERE, "knn_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.knn_ref_inv_norms.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib.knn_ref_search.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def _dt(db):
    assert db.dtype in (np.float32, np.float16) and db.ndim == 2 and db.flags.c_contiguous
    return 0 if db.dtype == np.float32 else 1


def normalize_queries(q):
    """The caller-side query normalisation of the reference (numpy fp32): ``ddpm.py:907``, ``dsetbuilder.py:487``.  Queries are taken as
    float32 (fp16 database rows are widened first, as the product's gather does, ``ddpm.py:921``): the reference's fp16-arithmetic norm of
    an fp16 row differs from this by a positive per-query factor (1 + O(1e-3)), which cannot change a ranking."""
    q = np.asarray(q, dtype=np.float32)
    return q / np.linalg.norm(q, axis=1)[:, np.newaxis]


def pairwise_sum_f32(a):
    """NumPy's float32 `add.reduce` over a contiguous row, spelled out (numpy/core/src/umath/loops_utils.h.src `FLOAT_pairwise_sum`,
    PW_BLOCKSIZE = 128): fewer than 8 elements -> sequential; up to 128 -> eight strided accumulators r[j] += a[8 i + j] combined as
    ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the n % 8 tail sequentially; longer -> split at n2 = n/2 rounded down to a multiple of 8 and
    add the two halves.  This is the summation order the CUDA normalisation kernel (csrc/knn.cu knn_normalize_kernel) reproduces; the test
    suite pins this restatement against numpy itself."""
    a = np.asarray(a, dtype=np.float32)
    n = a.shape[0]
    f = np.float32
    if n < 8:
        res = f(0.0)
        for v in a:
            res = f(res + v)
        return res
    if n <= 128:
        r = [f(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = f(r[j] + a[i + j])
            i += 8
        res = f(f(f(r[0] + r[1]) + f(r[2] + r[3])) + f(f(r[4] + r[5]) + f(r[6] + r[7])))
        while i < n:
            res = f(res + a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return f(pairwise_sum_f32(a[:n2]) + pairwise_sum_f32(a[n2:]))


def normalize_queries_restated(q):
    """`normalize_queries` without calling np.linalg.norm: separately rounded fp32 squares, `pairwise_sum_f32`, IEEE sqrt and division."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    out = np.empty_like(q)
    for i in range(q.shape[0]):
        sq = (q[i] * q[i]).astype(np.float32)
        nrm = np.sqrt(np.float32(np.float32(0.0) + pairwise_sum_f32(sq)), dtype=np.float32)
        out[i] = q[i] / nrm
    return out


def inv_norms(db):
    lib, inv = _load(), np.empty(db.shape[0], dtype=np.float32)
    lib.knn_ref_inv_norms(db.ctypes.data, db.shape[0], db.shape[1], _dt(db), inv.ctypes.data)
    return inv


def search(db, q_hat, k, inv=None, idx_base=0, return_scores=False):
    """-> (idx int64 [Q,k], dist float32 [Q,k]) best first."""
    lib = _load()
    q_hat = np.ascontiguousarray(q_hat, dtype=np.float32)
    inv = inv_norms(db) if inv is None else inv
    nq = q_hat.shape[0]
    idx, dist = np.empty((nq, k), np.int64), np.empty((nq, k), np.float32)
    sc = np.empty((nq, k), np.float64)
    lib.knn_ref_search(db.ctypes.data, inv.ctypes.data, db.shape[0], db.shape[1], _dt(db), q_hat.ctypes.data, nq, k,
                       idx_base, idx.ctypes.data, dist.ctypes.data, sc.ctypes.data)
    return (idx, dist, sc) if return_scores else (idx, dist)


def search_numpy(db, q_hat, k, block=262144):
    """Vectorised numpy restatement (fp64 GEMM + stable lexicographic sort)."""
    n, best = db.shape[0], None
    q64 = np.asarray(q_hat, dtype=np.float64)
    for s in range(0, n, block):
        d = db[s:s + block].astype(np.float64)
        inv = (1.0 / np.sqrt((d * d).sum(1))).astype(np.float32).astype(np.float64)
        sc = (q64 @ d.T) * inv[None]
        ids = np.broadcast_to(np.arange(s, s + d.shape[0])[None], sc.shape)
        if best is not None:
            sc, ids = np.concatenate([best[0], sc], 1), np.concatenate([best[1], ids], 1)
        order = np.lexsort((ids, -sc), axis=1)[:, :k]
        best = (np.take_along_axis(sc, order, 1), np.take_along_axis(ids, order, 1))
    return best[1].astype(np.int64), best[0].astype(np.float32)


def merge_shards(parts, k):
    """Merge per-shard (idx, score64) lists by (score desc, idx asc) -- the multi-GPU exchange's definition."""
    idx = np.concatenate([p[0] for p in parts], 1)
    sc = np.concatenate([p[1] for p in parts], 1)
    order = np.lexsort((idx, -sc), axis=1)[:, :k]
    return np.take_along_axis(idx, order, 1), np.take_along_axis(sc, order, 1)
