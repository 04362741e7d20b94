"""Deterministic weights addressed by parameter NAME, shared by the golden generator (which loads them into the reference's
modules) and the tests (which load them into the oracle and the CUDA executors): the fixtures then only need to store inputs and
outputs.  numpy's legacy RandomState stream is frozen by numpy's compatibility policy, and seeding per name makes the values
independent of parameter order -- so a state-dict key that differs between the reference and a restatement shows up as a failed
comparison, not as a silently re-ordered tensor."""
import zlib

import numpy as np


def tensor_for(name, shape, seed):
    rs = np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0xFFFFFFFF)
    shape = tuple(int(s) for s in shape)
    v = rs.standard_normal(shape).astype(np.float32)
    if name.endswith("positional_encoding"):
        return v / np.float32(shape[0] ** 0.5)
    if len(shape) >= 2:                                   # conv / linear / embedding: O(1) activations
        return v / np.float32(np.prod(shape[1:]) ** 0.5)
    if name.endswith("weight"):                           # norm scales
        return np.float32(1.0) + np.float32(0.1) * v
    return np.float32(0.05) * v                           # biases / norm shifts


def state_dict_for(named_shapes, seed):
    """named_shapes: iterable of (name, shape) -> {name: torch.float32 tensor}"""
    import torch
    return {k: torch.from_numpy(tensor_for(k, s, seed)) for k, s in named_shapes}


def fill_(module, seed):
    """Loads name-addressed weights into every parameter of a torch module."""
    sd = state_dict_for(((k, v.shape) for k, v in module.state_dict().items()), seed)
    module.load_state_dict(sd)
    return module


def make_db(n=600):
    """The synthetic retrieval database of ref_pipeline_tiny.npz: RAW (un-normalised) fp16 CLIP-like rows with very different norms,
    plus a neighbour-frequency memory in the reference's pickle layout {'nn_memory', 'id_count'} (ddpm.py:166-176)."""
    scale = 0.5 + 60.0 * np.abs(tensor_for("db.scale", (n,), 41))[:, None]
    db = (tensor_for("db.embedding", (n, 512), 41) * np.float32(22.0) * scale).astype(np.float16)
    mem = np.random.RandomState(42).permutation(n)[:200].astype(np.int64)
    return db, mem, {int(i): int(1 + (7 * i) % 13) for i in mem}


def round_dense_weights_to_fp16(sd):
    """The RARM executor's fp16 mode keeps every dense-layer weight matrix in fp16 (embedding table, positional encoding, biases and norms
    stay fp32, accumulation is fp32): the same state dict with exactly those tensors rounded, for a tight comparison of that mode."""
    import torch
    dense = (".to_q.weight", ".to_k.weight", ".to_v.weight", ".to_out.0.weight", ".ff.net.0.proj.weight", ".ff.net.2.weight")
    return {k: (v.to(torch.float16).to(torch.float32) if (k.endswith(dense) or k == "proj_out.weight") else v) for k, v in sd.items()}
