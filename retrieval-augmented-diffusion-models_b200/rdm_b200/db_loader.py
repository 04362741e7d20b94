"""Row-range loader for the retrieval database on disk (host side).

The database is a directory of `.npz` parts with the keys `embedding`, `img_id`, `patch_coords` (`dsetbuilder.py:181-236`); the reference
concatenates ALL parts on every process (184-300 s, 21-43 GB of host memory).  For a row-sharded searcher a rank only needs the rows
`[lo, hi)` it owns: `part_row_counts` reads just the `.npy` headers inside the zip members (no data), `load_rows` then opens only the parts
that overlap the range and slices them.  fp16 stays fp16.  (Pure host code; the upload itself is `B200Searcher(rows, idx_base=lo)`.)"""
import os
import zipfile
from glob import glob

import numpy as np


def _npz_member_shape(path, key):
    with zipfile.ZipFile(path) as z, z.open(key + ".npy") as f:
        major, minor = np.lib.format.read_magic(f)
        read = np.lib.format.read_array_header_1_0 if (major, minor) == (1, 0) else np.lib.format.read_array_header_2_0
        shape, _, dtype = read(f)
    return shape, dtype


def list_parts(saved_embeddings):
    """Sorted `.npz` parts of a database path (a single file or a directory), the order `load_embeddings` concatenates them in."""
    if os.path.isfile(saved_embeddings):
        return [saved_embeddings]
    parts = sorted(glob(os.path.join(saved_embeddings, "*.npz")))
    if not parts:
        raise FileNotFoundError(f"no .npz parts under {saved_embeddings}")
    return parts


def part_row_counts(parts, key="embedding"):
    """Rows per part, from the array headers only."""
    return [int(_npz_member_shape(p, key)[0][0]) for p in parts]


def load_rows(saved_embeddings, lo, hi, keys=("embedding", "img_id", "patch_coords")):
    """`{key: rows [lo, hi) of the concatenated database}` plus `'n_total'`; touches only the parts that overlap the range."""
    parts = list_parts(saved_embeddings)
    counts = part_row_counts(parts)
    n_total = sum(counts)
    lo, hi = max(0, int(lo)), min(int(hi), n_total)
    if lo >= hi:
        raise ValueError(f"empty row range [{lo}, {hi}) of a {n_total}-row database")
    out = {k: [] for k in keys}
    base = 0
    for path, cnt in zip(parts, counts):
        a, b = max(lo, base), min(hi, base + cnt)
        if a < b:
            with np.load(path) as z:
                for k in keys:
                    if k in z.files:
                        out[k].append(z[k][a - base:b - base])
        base += cnt
    res = {k: (np.concatenate(v, axis=0) if len(v) > 1 else v[0]) for k, v in out.items() if v}
    res["n_total"] = n_total
    return res
