"""`DatasetBuilder.build_data_pool` (reference dsetbuilder.py:317-437) -- bulk embedding of patch batches and the
chunked `.npz` pool writer -- on CPU with a stand-in retriever; the files must load back through the product's own `load_embeddings`."""
import glob
import os

import numpy as np
import pytest
import torch

import rdm  # noqa: F401
from rdm.data.retrieval_dataset.dsetbuilder import DatasetBuilder


class _Retriever(torch.nn.Module):
    """embedding = fixed random projection of the mean-pooled patch (deterministic, device-free)"""
    def __init__(self, dim=32):
        super().__init__()
        self.proj = torch.nn.Parameter(torch.randn(3, dim, generator=torch.Generator().manual_seed(0)), requires_grad=False)

    def forward(self, x):                       # [b, 3, h, w]
        return x.mean(dim=(2, 3)) @ self.proj


def _loader(n_batches, bs, seed=0, with_class=False):
    g = torch.Generator().manual_seed(seed)
    for i in range(n_batches):
        b = {"patch": torch.rand(bs, 8, 8, 3, generator=g) * 2 - 1, "img_id": torch.arange(i * bs, (i + 1) * bs),
             "patch_coords": torch.randint(0, 100, (bs, 4), generator=g)}
        if with_class:
            b["class_id"] = torch.full((bs,), i)
        yield b


def _builder(tmp_path, **kw):
    b = DatasetBuilder(retriever_config=None, load_patch_dataset=False, gpu=False, batch_size=4, **kw)
    b._retriever = _Retriever()
    b.pool_dir = str(tmp_path)
    return b


def test_chunked_pool_files_and_reload(tmp_path):
    b = _builder(tmp_path, max_pool_size=20, chunk_size=8)
    files = b.build_data_pool(_loader(10, 4, with_class=True))
    # like the reference, max_pool_size is checked when a chunk completes: 20 rows are reached inside the third chunk of 8
    assert [os.path.basename(f) for f in files] == ["8x32-part_1.npz", "8x32-part_2.npz", "8x32-part_3.npz"]
    parts = [np.load(f) for f in files]
    emb = np.concatenate([p["embedding"] for p in parts])
    assert emb.dtype == np.float16 and emb.shape == (24, 32)
    want = torch.cat([_Retriever()(x["patch"].permute(0, 3, 1, 2)) for x in _loader(6, 4)]).numpy().astype(np.float16)
    assert np.array_equal(emb, want)
    assert np.array_equal(np.concatenate([p["img_id"] for p in parts]), np.arange(24))
    assert np.array_equal(np.concatenate([p["class_id"] for p in parts]), np.repeat(np.arange(6), 4))
    # the directory loads back through load_embeddings (part order = name order)
    r = DatasetBuilder(retriever_config=None, load_patch_dataset=False, gpu=False, saved_embeddings=b.saved_embeddings)
    assert r.data_pool["embedding"].shape == (24, 32)
    assert sorted(map(tuple, r.data_pool["embedding"].tolist())) == sorted(map(tuple, emb.tolist()))


def test_single_file_and_restart(tmp_path):
    b = _builder(tmp_path, max_pool_size=12)
    files = b.build_data_pool(_loader(10, 4))
    assert len(files) == 1 and os.path.basename(files[0]) == "12x32.npz"
    first = np.load(files[0])["embedding"]
    # a longer pool is continued from the saved one: the first 12 rows are skipped, not recomputed
    c = _builder(tmp_path / "more", max_pool_size=20, saved_embeddings=files[0])
    c.max_pool_size = 20
    calls = []
    orig = c.embed
    c.embed = lambda batch, is_caption=False: (calls.append(batch.shape[0]), orig(batch))[1]
    new = c.build_data_pool(_loader(10, 4))
    assert calls == [4, 4] and len(new) == 1
    assert np.array_equal(np.load(new[0])["img_id"], np.arange(12, 20))
    with pytest.raises(NotImplementedError):
        c.build_data_pool()
