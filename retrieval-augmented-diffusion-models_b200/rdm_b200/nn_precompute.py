"""Offline neighbour precompute over the device-resident searcher.

What the reference does in `scripts/search_neighbors.py:355-450` (functions `search_nns`, `save_pkl`), re-built around the GPU path:
every batch of query patches (or captions) is CLIP-embedded on the device, L2-normalised, searched exactly, and every example gets ONE
pickle `embeddings/{k}_nns-img{id:09d}.p` holding `{patch_grid: {'embeddings', 'img_ids', 'patch_coords', 'nn_ids'}}` -- the layout
`QueryDataset.load_nns` reads back (`rdm/data/base.py:925-939`).  In counting mode the neighbour-id histogram that becomes `nn_memory`
(`{'nn_memory', 'id_count'}`, `ddpm.py:168-176`) is returned instead.  Dataset construction and image I/O stay with the caller, who passes
any iterable of `{'patches': [b, n, h, w, c] in [-1, 1]}` / `{'caption': [str, ...]}` batches carrying a `batch_size` attribute.
"""
import os
import pickle
from collections import Counter

import numpy as np
import torch

ENTRY_KEYS = (("embeddings", "embeddings"), ("img_ids", "img_ids"), ("patch_coords", "patch_coords"), ("nn_ids", "nns"))   # file key <- search_k_nearest key


def _dump(path, obj):
    with open(path, "wb") as f:
        pickle.dump(obj, f, protocol=pickle.HIGHEST_PROTOCOL)


def save_pkl(filepath, save_it, npatches_perside, corrupts, i, j, start_id, dset_batch_size):
    """Adds `save_it[npatches_perside]` to the example's pickle (one file can hold several patch grids).  An unreadable existing file is
    replaced when this is the 1 x 1 grid, otherwise the example id is recorded in `corrupts` -- same policy as the reference."""
    example = start_id + i * dset_batch_size + j
    if not os.path.isfile(filepath):
        _dump(filepath, save_it)
        return corrupts
    try:
        with open(filepath, "rb") as f:
            merged = pickle.load(f)
        merged[npatches_perside] = save_it[npatches_perside]
        _dump(filepath, merged)
    except Exception as e:                                          # truncated / foreign file
        print(f"{type(e).__name__} while updating {filepath}: {e}")
        if npatches_perside == 1:
            _dump(filepath, save_it)
        else:
            corrupts.add(example)
    return corrupts


def search_nns(dataset_builder, qloader, device="cuda", mode="img", save=False, npatches_perside=None, base_savedir=None, nn_paths=None,
               corrupts=None, start_id=0, max_its=None):
    """Same arguments and return values as the reference function: `{example id: relative pickle path}` when saving, else `{row id: count}`."""
    if dataset_builder.searcher is None:
        raise RuntimeError("train_searcher() must be called before the neighbour precompute")
    per_batch = qloader.batch_size
    if save:
        if base_savedir is None or npatches_perside is None or not os.path.isdir(os.path.join(base_savedir, "embeddings")):
            raise ValueError("saving needs npatches_perside and an existing <base_savedir>/embeddings directory")
        nn_paths = {} if nn_paths is None else nn_paths
        corrupts = set() if corrupts is None else corrupts
    histogram = Counter()
    for it, batch in enumerate(qloader):
        if max_its is not None and it >= max_its:
            break
        if mode == "img":
            patches = batch["patches"].to(device)
            b, n = patches.shape[:2]
            found = dataset_builder.search_k_nearest(patches.flatten(0, 1), visualize=False, is_caption=False)       # (b n) h w c
        else:
            b, n = len(batch["caption"]), 1
            found = dataset_builder.search_k_nearest(batch["caption"], visualize=False, is_caption=True)
        if not save:
            ids, counts = np.unique(found["nns"], return_counts=True)
            histogram.update({int(i): int(c) for i, c in zip(ids, counts)})
            continue
        per_example = {fk: np.asarray(found[sk]).reshape(b, n, *np.asarray(found[sk]).shape[1:]) for fk, sk in ENTRY_KEYS}
        for j in range(b):
            example = start_id + it * per_batch + j
            rel = f"embeddings/{dataset_builder.k}_nns-img{example:09d}.p"
            entry = {npatches_perside: {fk: per_example[fk][j] for fk, _ in ENTRY_KEYS}}
            corrupts = save_pkl(os.path.join(base_savedir, rel), entry, npatches_perside, corrupts, it, j, start_id, per_batch)
            nn_paths[example] = rel
    return nn_paths if save else dict(histogram)


def build_nn_memory(return_ids):
    """`{'nn_memory': ids sorted by how often they were retrieved, 'id_count': {id: count}}` -- what `ddpm.py:168-176` loads."""
    order = sorted(return_ids.items(), key=lambda kv: (-kv[1], kv[0]))
    return {"nn_memory": np.asarray([k for k, _ in order], dtype=np.int64), "id_count": dict(return_ids)}
