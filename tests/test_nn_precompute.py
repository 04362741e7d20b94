"""CPU: host logic of the offline neighbour precompute (rdm_b200/nn_precompute.py) with a stand-in builder -- file naming, the
per-example pickle layout `QueryDataset.load_nns` reads (rdm/data/base.py:925-939), merging of several patch grids into one file,
the corrupt-file policy of scripts/search_neighbors.py:355-379, and the nn_memory histogram."""
import pickle

import numpy as np
import torch

from rdm_b200.nn_precompute import build_nn_memory, save_pkl, search_nns


class FakeBuilder:
    """search_k_nearest with the mirror's output keys; neighbours of query i are rows i, i+1, ... (deterministic)."""
    k, searcher = 3, object()

    def __init__(self):
        self.pool = np.arange(50 * 4, dtype=np.float32).reshape(50, 4)
        self.calls = 0

    def search_k_nearest(self, queries, visualize=False, is_caption=False):
        n = len(queries)
        nns = (np.arange(n)[:, None] + np.arange(self.k)[None] + 7 * self.calls) % 50
        self.calls += 1
        return {"embeddings": self.pool[nns], "nns": nns, "img_ids": nns * 10, "patch_coords": np.zeros((n, self.k, 4), np.int32), "queries": queries}


class Loader(list):
    batch_size = 2


def _batches():
    return Loader({"patches": torch.zeros(2, 4, 8, 8, 3)} for _ in range(3))


def test_pickles_have_the_reference_layout_and_merge_patch_grids(tmp_path):
    (tmp_path / "embeddings").mkdir()
    b = FakeBuilder()
    paths = search_nns(b, _batches(), device="cpu", save=True, npatches_perside=2, base_savedir=str(tmp_path), start_id=10)
    assert sorted(paths) == list(range(10, 16)) and paths[13] == "embeddings/3_nns-img000000013.p"
    with open(tmp_path / paths[13], "rb") as f:
        e = pickle.load(f)
    assert list(e) == [2] and set(e[2]) == {"embeddings", "img_ids", "patch_coords", "nn_ids"}
    assert e[2]["nn_ids"].shape == (4, 3) and e[2]["embeddings"].shape == (4, 3, 4) and e[2]["patch_coords"].shape == (4, 3, 4)
    assert np.array_equal(e[2]["embeddings"], b.pool[e[2]["nn_ids"]]) and np.array_equal(e[2]["img_ids"], e[2]["nn_ids"] * 10)
    # a second pass with the 1 x 1 grid lands in the SAME files next to the 2 x 2 entry
    one = Loader({"patches": torch.zeros(2, 1, 16, 16, 3)} for _ in range(3))
    search_nns(b, one, device="cpu", save=True, npatches_perside=1, base_savedir=str(tmp_path), start_id=10, nn_paths=paths)
    with open(tmp_path / paths[13], "rb") as f:
        e = pickle.load(f)
    assert sorted(e) == [1, 2] and e[1]["nn_ids"].shape == (1, 3)
    # max_its stops early; captions count one query per example
    few = search_nns(b, Loader({"caption": ["a", "b"]} for _ in range(5)), mode="text", save=False, max_its=2)
    assert sum(few.values()) == 2 * 2 * 3


def test_corrupt_file_policy(tmp_path):
    f = tmp_path / "x.p"
    f.write_bytes(b"not a pickle")
    bad = save_pkl(str(f), {2: {"nn_ids": np.zeros(1)}}, 2, set(), i=1, j=1, start_id=100, dset_batch_size=4)
    assert bad == {105} and f.read_bytes() == b"not a pickle"                 # other grids: recorded, file untouched
    ok = save_pkl(str(f), {1: {"nn_ids": np.ones(1)}}, 1, set(), i=1, j=1, start_id=100, dset_batch_size=4)
    assert ok == set() and list(pickle.loads(f.read_bytes())) == [1]          # the 1 x 1 grid overwrites


def test_histogram_and_nn_memory():
    hist = search_nns(FakeBuilder(), _batches(), device="cpu", save=False)
    assert sum(hist.values()) == 3 * 2 * 4 * 3
    mem = build_nn_memory(hist)
    counts = [hist[int(i)] for i in mem["nn_memory"]]
    assert counts == sorted(counts, reverse=True) and mem["id_count"] == hist


def test_files_and_return_values_equal_the_reference_writer(tmp_path):
    """tests/golden/ref_search_nns.p: everything the REFERENCE's `search_nns` / `save_pkl` (scripts/search_neighbors.py:355-450) returned
    and wrote for the scenario of tests/golden/retro_stub.py (2 x 2 grid pass, a truncated file, 1 x 1 pass into the same files, caption
    counting with max_its).  The product's implementation must produce the same names, the same pickle contents and the same returns."""
    import os
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import retro_stub
    with open(os.path.join(ROOT, "tests", "golden", "ref_search_nns.p"), "rb") as f:
        want = pickle.load(f)
    got = retro_stub.precompute_scenario(search_nns, str(tmp_path))
    assert got["paths"] == want["paths"] and got["paths_after_second_pass"] == want["paths_after_second_pass"] and got["counts"] == want["counts"]
    assert sorted(got["files"]) == sorted(want["files"])
    for name, entry in want["files"].items():
        assert sorted(got["files"][name]) == sorted(entry), name                      # patch grids present (the truncated file holds grid 1 only)
        for grid, fields in entry.items():
            assert sorted(got["files"][name][grid]) == sorted(fields)
            for key, arr in fields.items():
                g = got["files"][name][grid][key]
                assert g.dtype == arr.dtype and np.array_equal(g, arr), (name, grid, key)
