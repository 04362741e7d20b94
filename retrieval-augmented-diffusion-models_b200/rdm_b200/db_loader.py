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


# ---- direct-to-HBM path ---------------------------------------------------------------------------------------
def open_member(path, key):
    """The array `key` of one `.npz` part WITHOUT reading it: a read-only `np.memmap` over the `.npy` payload inside the zip when the member is
    stored uncompressed (what `np.savez` writes -- the reference's `save_datapool`, dsetbuilder.py:439-459), else the decompressed array."""
    with zipfile.ZipFile(path) as z:
        info = z.getinfo(key + ".npy")
        if info.compress_type != zipfile.ZIP_STORED:
            with np.load(path) as npz:
                return npz[key]
        with open(path, "rb") as f:
            f.seek(info.header_offset)
            local = f.read(30)                                         # local file header: name / extra lengths at bytes 26..30
            name_len, extra_len = int.from_bytes(local[26:28], "little"), int.from_bytes(local[28:30], "little")
            payload = info.header_offset + 30 + name_len + extra_len
            f.seek(payload)
            major, minor = np.lib.format.read_magic(f)
            read = np.lib.format.read_array_header_1_0 if (major, minor) == (1, 0) else np.lib.format.read_array_header_2_0
            shape, fortran, dtype = read(f)
            data_off = f.tell()
        if fortran or dtype.hasobject:
            with np.load(path) as npz:
                return npz[key]
        return np.memmap(path, dtype=dtype, mode="r", offset=data_off, shape=shape)


def load_rows_to_device(saved_embeddings, lo, hi, device, key="embedding", chunk_bytes=64 << 20, out=None):
    """Rows [lo, hi) of the concatenated database straight into ONE device tensor (dtype of the file: fp16 stays fp16), part by part and
    chunk by chunk through two pinned staging buffers: the host copy of chunk i+1 (page cache / disk -> pinned) overlaps the DMA of chunk i.
    No host-side concatenation, never more than `2 * chunk_bytes` of extra host memory.  -> (tensor [hi-lo, d] on `device`, stats dict)."""
    import time

    import torch
    parts = list_parts(saved_embeddings)
    counts = part_row_counts(parts, key)
    n_total = sum(counts)
    lo, hi = max(0, int(lo)), min(int(hi), n_total)
    if lo >= hi:
        raise ValueError(f"empty row range [{lo}, {hi}) of a {n_total}-row database")
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("load_rows_to_device needs a CUDA device (the host path is load_rows)")
    shape0, dtype = _npz_member_shape(parts[0], key)
    d = int(np.prod(shape0[1:]))
    tdtype = {np.dtype("float16"): torch.float16, np.dtype("float32"): torch.float32}[np.dtype(dtype)]
    if out is None:
        out = torch.empty((hi - lo, d), dtype=tdtype, device=device)
    assert out.shape == (hi - lo, d) and out.dtype == tdtype and out.is_contiguous()
    rows_per_chunk = max(1, chunk_bytes // (d * np.dtype(dtype).itemsize))
    stage = [torch.empty((rows_per_chunk, d), dtype=tdtype).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    copy_stream = torch.cuda.Stream(device=device)
    t0, nbytes, slot, base = time.time(), 0, 0, 0
    with torch.cuda.stream(copy_stream):
        for path, cnt in zip(parts, counts):
            a, b = max(lo, base), min(hi, base + cnt)
            if a < b:
                src = open_member(path, key)
                for r0 in range(a, b, rows_per_chunk):
                    r1 = min(b, r0 + rows_per_chunk)
                    done[slot].synchronize()                            # the DMA that last used this staging buffer has finished
                    buf = stage[slot][:r1 - r0]
                    buf.numpy()[...] = np.asarray(src[r0 - base:r1 - base]).reshape(r1 - r0, d)      # disk / page cache -> pinned
                    out[r0 - lo:r1 - lo].copy_(buf, non_blocking=True)                               # pinned -> HBM (async DMA)
                    done[slot].record(copy_stream)
                    nbytes += buf.numel() * buf.element_size()
                    slot ^= 1
                del src
            base += cnt
    copy_stream.synchronize()
    torch.cuda.current_stream(device).wait_stream(copy_stream)
    sec = time.time() - t0
    return out, {"rows": hi - lo, "n_total": n_total, "bytes": nbytes, "seconds": sec, "gb_per_s": nbytes / sec / 1e9, "parts_touched": sum(
        1 for i, c in enumerate(counts) if max(lo, sum(counts[:i])) < min(hi, sum(counts[:i]) + c))}


class DeviceRows:
    """Stand-in for `data_pool['embedding']` when the rows live only in HBM: `len`, `.shape`, `.dtype` and fancy indexing (a device gather of
    RAW rows returned as a float32 numpy array -- every consumer of the reference converts to float right away, ddpm.py:921)."""

    def __init__(self, searcher, n_total, d, dtype):
        self.searcher, self.shape, self.dtype, self.ndim = searcher, (int(n_total), int(d)), np.dtype(dtype), 2

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        import torch
        idx = np.asarray(idx)
        flat = torch.as_tensor(idx.reshape(-1).astype(np.int64), device=self.searcher.device)
        return self.searcher.gather_device(flat).cpu().numpy().reshape(*idx.shape, self.shape[1])
