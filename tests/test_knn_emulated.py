"""CPU: the SIMT path of the exact cosine searcher -- csrc/knn.cu as written (row scan with per-warp bulk-copy rings and mbarriers,
sample / threshold / main / select kernels, locked fallback lists, shard merge, raw-row gather) -- under the host emulation of CUDA
(tests/emu/), through the C ABI and the product's Python wrapper, bit-exact against the oracle (indices AND fp64 scores).  The
tensor-core scan of knn_tc.cu needs hardware: RDM_KNN_NO_TC=1 routes every batch through the SIMT kernels, as tests/test_knn_gpu.py
does for its comparison of the two."""
import contextlib
import ctypes
import os
import shutil
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import knn as oknn

sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ to build the emulated library")


@pytest.fixture()
def emulated(monkeypatch):
    os.environ["RDM_KNN_NO_TC"] = "1"                      # read once by the library, before the first fp16 / d = 512 search
    import build_emu
    from rdm_b200 import _lib
    L = _lib.bind(ctypes.CDLL(build_emu.build()), [n for n in _lib.SIGNATURES if n.startswith("rdm_knn_")] + ["rdm_last_error", "rdm_launch_count"])
    monkeypatch.setattr(_lib, "_lib", L)
    monkeypatch.setattr(_lib, "resolve_device", lambda d: torch.device("cpu"))
    monkeypatch.setattr(_lib, "device_ctx", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(_lib, "stream_ptr", lambda d=None: None)
    return L


def check(db, q, k, idx_base=0):
    from rdm_b200.knn import B200Searcher
    s = B200Searcher(db, device="cpu", idx_base=idx_base)
    qh = oknn.normalize_queries(q)
    idx, dist, sc = s.search_device(torch.from_numpy(qh), k, return_scores=True)
    wi, wd, ws = oknn.search(db, qh, k, idx_base=idx_base, return_scores=True)
    assert np.array_equal(idx.numpy(), wi), "neighbour indices"
    assert np.array_equal(sc.numpy().view(np.int64), ws.view(np.int64)), "fp64 scores, bit patterns"
    assert np.array_equal(dist.numpy(), wd)
    return s, wi


@pytest.mark.parametrize("dtype", [np.float16, np.float32])
@pytest.mark.parametrize("n,nq,k", [(3001, 1, 4), (2500, 5, 8), (999, 16, 24), (40, 2, 4)])
def test_search_is_bit_exact(emulated, dtype, n, nq, k):
    rng = np.random.default_rng(n + nq)
    db = (rng.standard_normal((n, 512)) * rng.uniform(0.5, 8, (n, 1))).astype(dtype)
    q = rng.standard_normal((nq, 512)).astype(np.float32)
    q[0] = db[n // 2].astype(np.float32)                  # a database row as query: retrieves itself first
    s, wi = check(db, q, k)
    assert wi[0, 0] == n // 2
    rows = s.gather_device(torch.from_numpy(wi))
    assert np.array_equal(rows.numpy(), db[wi].astype(np.float32))


def test_duplicates_break_ties_by_lowest_index_and_shards_merge(emulated):
    from rdm_b200.knn import B200Searcher, merge_device
    rng = np.random.default_rng(7)
    db = rng.standard_normal((1200, 256)).astype(np.float16)
    db[700] = db[3]; db[1100] = db[3]                    # exact duplicates: equal scores -> ascending index
    q = np.concatenate([db[[3]].astype(np.float32), rng.standard_normal((2, 256)).astype(np.float32)])
    _, wi = check(db, q, 6)
    assert list(wi[0, :3]) == [3, 700, 1100]
    qh = torch.from_numpy(oknn.normalize_queries(q))
    parts = [B200Searcher(db[lo:hi], device="cpu", idx_base=lo).search_device(qh, 6, return_scores=True) for lo, hi in ((0, 500), (500, 1200))]
    idx, dist, sc = merge_device(torch.stack([p[0] for p in parts]), torch.stack([p[2] for p in parts]), 6)
    assert np.array_equal(idx.numpy(), wi)
    rows = sum(B200Searcher(db[lo:hi], device="cpu", idx_base=lo).gather_device(idx) for lo, hi in ((0, 500), (500, 1200)))      # out-of-shard rows are zeros
    assert np.array_equal(rows.numpy(), db[wi].astype(np.float32))


@pytest.mark.parametrize("dtype,n,nq,k", [(np.float16, 5000, 5, 8), (np.float32, 3000, 2, 24)])
def test_near_duplicate_cluster_is_exact(emulated, dtype, n, nq, k):
    """400 rows that differ from the query row by less than the fp32 rounding of the scan's score: the exact fp64 top-k (and the
    self-match at rank 0) must still come out -- the scans relax every fp32 decision by the score slack and the select kernel re-ranks
    every survivor within the slack exactly (knn.cu: SCORE_SLACK).  (Without the slack this returned other members of the cluster.)"""
    rng = np.random.default_rng(n)
    db = (rng.standard_normal((n, 512)) * rng.uniform(0.5, 8, (n, 1))).astype(dtype)
    db[1000:1400] = db[1000] + (rng.standard_normal((400, 512)) * 1e-3).astype(dtype)
    q = rng.standard_normal((nq, 512)).astype(np.float32)
    q[0] = db[1000].astype(np.float32)
    _, wi = check(db, q, k)
    assert set(wi[0]) <= set(range(1000, 1400))         # (which member ranks first is decided by the last bits of the exact scores: the oracle's call)


@pytest.mark.parametrize("ndup", [1500, 2600])
def test_candidate_overflow_takes_the_fallback_pass(emulated, ndup):
    """More exact duplicates of the query row than the select kernel re-ranks in one pass (1500 > SEL_MAX = 1024 survivors within the
    slack) or than the main scan keeps per query (2600 > CAND_CAP = 2048): the overflow flag routes the batch through the locked
    per-CTA lists and the FROM_LISTS select, which must still return the lowest-index duplicates, bit-exact; the second query of the
    same batch (no cluster) is answered by the same fallback pass."""
    rng = np.random.default_rng(ndup)
    n = 4000
    db = (rng.standard_normal((n, 512)) * rng.uniform(0.5, 8, (n, 1))).astype(np.float16)
    dup = rng.choice(n, ndup, replace=False)
    db[dup] = db[dup[0]]
    q = np.concatenate([db[[dup[0]]].astype(np.float32), rng.standard_normal((1, 512)).astype(np.float32)])
    _, wi = check(db, q, 8)
    assert list(wi[0]) == sorted(dup)[:8]


@pytest.mark.parametrize("d", [256, 512, 768, 1024])
def test_query_normalisation_is_bit_identical_to_numpy(emulated, d):
    """rdm_knn_normalize (csrc/knn.cu knn_normalize_kernel) == `q / np.linalg.norm(q, axis=1)[:, np.newaxis]` (ddpm.py:907), bit patterns;
    rdm_knn_search_raw == normalise + rdm_knn_search."""
    from rdm_b200.knn import B200Searcher, normalize_device
    rng = np.random.default_rng(d)
    q = (rng.standard_normal((7, d)) * rng.uniform(0.01, 30.0, size=(7, 1))).astype(np.float32)
    want = q / np.linalg.norm(q, axis=1)[:, np.newaxis]
    got = normalize_device(torch.from_numpy(q)).numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    db = rng.standard_normal((700, d)).astype(np.float16)
    s = B200Searcher(db, device="cpu")
    a = s.search_raw_device(torch.from_numpy(q), 5, return_scores=True)
    b = s.search_device(torch.from_numpy(want), 5, return_scores=True)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
