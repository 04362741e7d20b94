"""CPU: row-range loading of a multi-part .npz database (rdm_b200/db_loader.py; dsetbuilder.py:181-236 layout) -- every range must equal
the slice of the reference-style full concatenation, shard ranges must tile the database, and only overlapping parts may be opened."""
import numpy as np
import pytest

from rdm_b200 import db_loader
from rdm_b200.knn import shard_range


def _make_db(tmp_path, sizes=(5, 1, 7, 3), dtype=np.float16):
    rng = np.random.default_rng(0)
    full = {"embedding": [], "img_id": [], "patch_coords": []}
    start = 0
    for i, n in enumerate(sizes):
        part = {"embedding": rng.standard_normal((n, 8)).astype(dtype), "img_id": np.arange(start, start + n), "patch_coords": rng.integers(0, 9, (n, 4)).astype(np.int32)}
        np.savez(tmp_path / f"part_{i:03d}.npz", **part)
        for k in full:
            full[k].append(part[k])
        start += n
    return {k: np.concatenate(v) for k, v in full.items()}


def test_headers_give_the_row_counts_without_reading_data(tmp_path):
    _make_db(tmp_path)
    parts = db_loader.list_parts(str(tmp_path))
    assert [p.split("/")[-1] for p in parts] == [f"part_{i:03d}.npz" for i in range(4)]
    assert db_loader.part_row_counts(parts) == [5, 1, 7, 3]


@pytest.mark.parametrize("world", [1, 2, 3, 5, 16])
def test_shard_ranges_tile_the_database(tmp_path, world):
    full = _make_db(tmp_path)
    n = full["embedding"].shape[0]
    got = []
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        if lo == hi:
            continue
        rows = db_loader.load_rows(str(tmp_path), lo, hi)
        assert rows["n_total"] == n and rows["embedding"].dtype == np.float16
        for k in ("embedding", "img_id", "patch_coords"):
            assert np.array_equal(rows[k], full[k][lo:hi]), (k, lo, hi)
        got.append(rows["img_id"])
    assert np.array_equal(np.concatenate(got), np.arange(n))


def test_only_overlapping_parts_are_opened(tmp_path, monkeypatch):
    full = _make_db(tmp_path)
    opened = []
    real = np.load
    monkeypatch.setattr(np, "load", lambda p, *a, **k: (opened.append(str(p).split("/")[-1]), real(p, *a, **k))[1])
    rows = db_loader.load_rows(str(tmp_path), 6, 9)                      # rows 6..8 live in part 2 only (parts hold 0-4, 5, 6-12, 13-15)
    assert opened == ["part_002.npz"] and np.array_equal(rows["embedding"], full["embedding"][6:9])
    with pytest.raises(ValueError):
        db_loader.load_rows(str(tmp_path), 16, 20)


@pytest.mark.parametrize("rank", [0, 1, 2])
def test_dataset_builder_reads_only_its_shard(tmp_path, monkeypatch, rank):
    """DatasetBuilder(shard=True) under an initialised process group: rank r holds rows shard_range(n, r, world) of the embeddings
    (what `train_searcher` uploads with idx_base = lo), the id / coordinate arrays stay complete (results index them globally)."""
    import torch
    import rdm  # noqa: F401  (installs the shims)
    from rdm.data.retrieval_dataset.dsetbuilder import DatasetBuilder
    full = _make_db(tmp_path)
    n, world = full["embedding"].shape[0], 3
    monkeypatch.setattr(torch.distributed, "is_initialized", lambda: True)
    monkeypatch.setattr(torch.distributed, "get_rank", lambda *a: rank)
    monkeypatch.setattr(torch.distributed, "get_world_size", lambda *a: world)
    b = DatasetBuilder(retriever_config=None, saved_embeddings=str(tmp_path), load_patch_dataset=False, gpu=False, shard=True)
    lo, hi = shard_range(n, rank, world)
    assert (b._row_base, b._n_total) == (lo, n) and b.max_pool_size == n
    assert np.array_equal(b.data_pool["embedding"], full["embedding"][lo:hi]) and b.data_pool["embedding"].dtype == np.float16
    assert np.array_equal(b.data_pool["img_id"], full["img_id"]) and np.array_equal(b.data_pool["patch_coords"], full["patch_coords"])
    # without shard=True every rank keeps the reference behaviour: the whole database
    b = DatasetBuilder(retriever_config=None, saved_embeddings=str(tmp_path), load_patch_dataset=False, gpu=False, shard=False)
    assert b._row_base is None and np.array_equal(b.data_pool["embedding"], full["embedding"])


@pytest.mark.parametrize("compressed", [False, True])
def test_members_are_memory_mapped_when_stored(tmp_path, compressed):
    """open_member: np.savez parts (stored) come back as read-only memmaps over the payload inside the zip -- no copy until rows are sliced --
    and compressed parts as ordinary arrays; both equal the saved data."""
    rng = np.random.default_rng(3)
    emb = rng.standard_normal((37, 8)).astype(np.float16)
    ids = np.arange(37)
    (np.savez_compressed if compressed else np.savez)(tmp_path / "p.npz", embedding=emb, img_id=ids)
    got = db_loader.open_member(str(tmp_path / "p.npz"), "embedding")
    assert isinstance(got, np.memmap) != compressed
    assert got.dtype == np.float16 and np.array_equal(np.asarray(got), emb)
    assert np.array_equal(np.asarray(db_loader.open_member(str(tmp_path / "p.npz"), "img_id")), ids)


def test_device_rows_view_indexes_like_the_host_array():
    class FakeSearcher:
        device = "cpu"

        def __init__(self, rows):
            self.rows = rows

        def gather_device(self, idx):
            import torch
            return torch.from_numpy(self.rows[idx.numpy()].astype(np.float32))
    rows = np.random.default_rng(0).standard_normal((20, 8)).astype(np.float16)
    v = db_loader.DeviceRows(FakeSearcher(rows), 20, 8, "float16")
    assert len(v) == 20 and v.shape == (20, 8) and v.dtype == np.float16
    nns = np.array([[3, 4], [19, 0]])
    assert np.array_equal(v[nns], rows[nns].astype(np.float32)) and v[nns].shape == (2, 2, 8)
