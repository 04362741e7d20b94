"""CPU: csrc/rarm.cu -- kernels AND host code, as written -- compiled against a host emulation of CUDA (tests/emu/: fibers for the
threads of a block, yield points at __syncthreads / warp shuffles, recorded-and-replayed graph launches) and driven through the same
C ABI and Python wrapper as on the GPU.  This is how the RARM decoder was checked in a round whose GPU budget was already spent: it
exercises indexing, barrier placement, the transposing shuffle reduction, weight re-packing, the key/value cache, the radix-select
sampler and the graph-replayed loop against the reference-pinned oracle.  It does not replace tests/test_zz_rarm_gpu.py (alignment,
resource limits, real concurrency and numerics of the device math library are only seen on hardware)."""
import ast
import ctypes
import os
import shutil
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import rarm as orarm

GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLD)
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import ref_weights  # noqa: E402

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ to build the emulated library")


@pytest.fixture(scope="module")
def emu_cls():
    import build_emu
    from rdm_b200 import _lib
    from rdm_b200.rarm import B200Rarm
    L = _lib.bind(ctypes.CDLL(build_emu.build()), [n for n in _lib.SIGNATURES if n.startswith("rdm_rarm_")] + ["rdm_last_error", "rdm_launch_count"])

    class EmuRarm(B200Rarm):
        def _library(self):
            return L

        def _resolve_device(self, device):
            return torch.device("cpu"), 0

        def _run(self, fn, *args):
            self._check(getattr(L, fn)(self._h, *args, None), fn)

    return EmuRarm


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm())


def small(emu_cls, mode):
    d = np.load(os.path.join(GOLD, "ref_rarm_small.npz"))
    cfg = ast.literal_eval(str(d["cfg_json"]))
    net = emu_cls("cpu", **cfg)
    sd = ref_weights.state_dict_for(net.shapes.items(), int(d["weight_seed"]))
    net.load_state_dict(sd)
    assert net.missing() == 0
    net.set_mode(mode)
    return d, cfg, sd, net


def consistent_with_uniform(probs, token, u, tol):
    cdf = probs.double().cumsum(-1)
    cdf = cdf / cdf[-1]
    lo = float(cdf[token - 1]) if token > 0 else 0.0
    return float(probs[token]) > 0 and lo - tol <= u <= float(cdf[token]) + tol


@pytest.mark.parametrize("mode,tol", [(0, 2e-6), (4, 3e-3)])
def test_cached_logits_match_reference_code(emu_cls, mode, tol):
    d, cfg, sd, net = small(emu_cls, mode)
    tok, ctx = torch.from_numpy(d["tokens"])[:, :6], torch.from_numpy(d["context"])
    got = net.forward(tok, ctx)
    assert rel(got, d["logits"][:, :6]) < tol                                  # the reference's own logits (causal: a prefix of them)
    assert rel(got[:, :5], d["logits_prefix5"]) < tol
    net.set_context(torch.cat([ctx, torch.zeros_like(ctx)]))                   # guidance doubling: 3 token rows against 6 context rows
    both = torch.stack([net.forward_token(tok[:, t], t) for t in range(3)], 1)
    assert rel(both[:3], d["logits"][:, :3]) < tol and rel(both[3:], d["logits_uncond"][:, :3]) < tol
    if mode == 4:                                                              # fp16 weights really are in use: far from fp32, tight against the oracle on rounded weights
        assert rel(got, d["logits"][:, :6]) > 1e-6
        assert rel(got, orarm.forward(ref_weights.round_dense_weights_to_fp16(sd), tok, ctx, cfg["n_heads"])) < 2e-6


def test_more_rows_than_one_chunk_and_rewind(emu_cls):
    """M = 11 rows (two 8-row chunks in the GEMV grid), then re-feeding position 1 with other tokens overwrites the cache row."""
    d, cfg, sd, net = small(emu_cls, 0)
    g = torch.Generator().manual_seed(9)
    tok = torch.randint(0, cfg["in_channels"], (11, 3), generator=g)
    ctx = torch.randn(11, 2, cfg["context_dim"], generator=g)
    assert rel(net.forward(tok, ctx), orarm.forward(sd, tok, ctx, cfg["n_heads"])) < 2e-6
    tok2 = tok.clone(); tok2[:, 1] = (tok[:, 1] + 7) % cfg["in_channels"]
    net.forward_token(tok2[:, 1], 1)
    l2 = net.forward_token(tok2[:, 2], 2)
    assert rel(l2, orarm.forward(sd, tok2, ctx, cfg["n_heads"])[:, 2]) < 2e-6


@pytest.mark.parametrize("V,top_k,guided", [(16384, 256, True), (48, 5, True), (1000, None, False), (5000, 1, True), (4097, 40, False)])
def test_guided_topk_draw_kernel_matches_oracle(emu_cls, V, top_k, guided):
    _, _, _, net = small(emu_cls, 0)
    g = torch.Generator().manual_seed(V + (top_k or 0))
    B = 3
    lc, lu = torch.randn(B, V, generator=g) * 3, torch.randn(B, V, generator=g) * 3
    lc[0, 7] = lc[0, 9]
    lu[0, 7] = lu[0, 9]
    lc[1] = -lc[1].abs()                                                       # all-negative logits: the other branch of the radix key
    scale, temp = (2.5, 0.8) if guided else (1.0, 1.3)
    u = torch.rand(B, generator=g)
    logits = torch.cat([lc, lu]) if guided else lc
    tok, probs = net.sample_step(logits, guidance_scale=scale, temperature=temp, top_k=top_k, uniforms=u, want_probs=True)
    want = orarm.step_probs(lc, lu if guided else None, scale, temp, top_k)
    assert torch.equal(probs > 0, want > 0)
    assert float((probs - want).abs().max()) < 1e-6
    assert torch.equal(tok, orarm.draw(want, u)) or all(consistent_with_uniform(want[b], int(tok[b]), float(u[b]), 1e-6) for b in range(B))
    greedy, _ = net.sample_step(logits, guidance_scale=scale, temperature=temp, top_k=top_k, uniforms=None)
    assert torch.equal(greedy, want.argmax(-1))


@pytest.mark.parametrize("scale", [1.0, 2.0])
def test_sampling_loop_graph_replay_token_by_token(emu_cls, scale):
    d, cfg, sd, net = small(emu_cls, 0)
    ctx = torch.from_numpy(d["context"])[:2]
    B, steps, top_k, temp = 2, 5, 6, 0.9
    c = torch.full((B, 1), cfg["in_channels"] - 1)
    g = torch.Generator().manual_seed(5)
    u = torch.rand(steps, B, generator=g)
    r = torch.cat([ctx, torch.zeros_like(ctx)]) if scale > 1.0 else ctx
    net.set_context(r)
    toks = net.sample(c, steps, temperature=temp, top_k=top_k, guidance_scale=scale, uniforms=u)       # first step eager, the rest replayed
    want, probs = orarm.sample(sd, cfg["n_heads"], c, torch.zeros((B, 0), dtype=torch.long), ctx, steps, temp, top_k, scale, u)
    for t in range(steps):
        for b in range(B):
            assert consistent_with_uniform(probs[t, b], int(toks[b, t + 1]), float(u[t, b]), 1e-5), (t, b)
        if not torch.equal(toks[:, t + 1], want[:, t]):
            break                                                             # a boundary draw: later tokens legitimately differ
    else:
        assert torch.equal(toks[:, 1:], want)
    net.set_context(r)                                                         # second call: every step from the captured graph
    assert torch.equal(net.sample(c, steps, temperature=temp, top_k=top_k, guidance_scale=scale, uniforms=u), toks)
    net.set_graph(False)
    net.set_context(r)
    assert torch.equal(net.sample(c, steps, temperature=temp, top_k=top_k, guidance_scale=scale, uniforms=u), toks)
    net.set_graph(True)
    net.set_context(r)                                                         # a given start prefix is kept and continued
    assert torch.equal(net.sample(toks[:, :3], steps - 2, temperature=temp, top_k=top_k, guidance_scale=scale, uniforms=u[2:]), toks)
    net.set_context(r)
    gr = net.sample(c, 3, temperature=1.0, top_k=None, guidance_scale=scale, uniforms=None)
    assert torch.equal(gr[:, 1:], orarm.sample(sd, cfg["n_heads"], c, torch.zeros((B, 0), dtype=torch.long), ctx, 3, guidance_scale=scale)[0])


def test_errors_are_reported(emu_cls):
    d, cfg, sd, net = small(emu_cls, 0)
    with pytest.raises(RuntimeError, match="context"):
        net.sample(torch.zeros((2, 1), dtype=torch.long), 2)                   # no context set
    net.set_context(torch.zeros(2, 3, cfg["context_dim"]))
    with pytest.raises(RuntimeError, match="sequence_length"):
        net.sample(torch.zeros((2, 1), dtype=torch.long), cfg["sequence_length"] + 1)
    with pytest.raises(RuntimeError):
        emu_cls("cpu", **dict(cfg, d_head=32))
    fresh = emu_cls("cpu", **cfg)
    with pytest.raises(RuntimeError, match="not loaded"):
        fresh.set_context(torch.zeros(2, 3, cfg["context_dim"]))


def test_gpu_test_bodies_run_under_emulation(emu_cls, monkeypatch):
    """The GPU parity tests of tests/test_zz_rarm_gpu.py, executed as they are against the emulated library (device "cpu"): keeps the
    GPU test file itself honest -- its helpers, shapes, tolerances and oracle calls -- in a container that cannot run it on hardware."""
    import rdm_b200.rarm as rr
    import test_zz_rarm_gpu as gpu_tests
    monkeypatch.setattr(rr, "B200Rarm", emu_cls)
    gpu_tests.test_cached_logits_match_reference_code("cpu", 4, 5e-3)
    gpu_tests.test_guided_topk_draw_kernel_matches_oracle("cpu", 16384, 256, True)
    gpu_tests.test_sampling_loop_token_by_token("cpu", 4, 2.0)


def test_product_sampling_flow_matches_reference_code(emu_cls):
    """The product end to end on the CPU -- `LatentImageRETRO.sample_from_rdata` / `sample` of this repository driving csrc/rarm.cu through
    the C ABI (under emulation) -- against tests/golden/ref_rarm_sampling.npz, the REFERENCE's own LatentImageRETRO run over its own
    RetrievalPatchTransformer (tests/golden/make_golden_ref.py): same NumPy / torch seeds, same database, same weights ->
    same query ids, same sampled token ids, same decoded images; greedy continuation of a given prefix likewise."""
    import types
    import make_golden_ref as gen
    import rdm  # noqa: F401
    from oracle import knn as oknn
    from rdm.models.autoregression.transformer import LatentImageRETRO
    from rdm_b200.rarm import MODE_FP32
    g = np.load(os.path.join(GOLD, "ref_rarm_sampling.npz"))
    model = LatentImageRETRO(**gen.rarm_model_cfg()).eval()
    assert sorted(model.state_dict().keys()) == sorted(str(k) for k in g["sd_keys"])          # the reference model's checkpoint keys
    tsd = ref_weights.state_dict_for(((k, v.shape) for k, v in model.transformer.state_dict().items()), 21)
    model.transformer.load_state_dict(tsd)
    eng = emu_cls("cpu", **model.transformer._cfg)
    eng.load_state_dict(tsd)
    eng.set_mode(MODE_FP32)
    model.transformer.engine = lambda device: eng
    db = ref_weights.make_db(int(g["n_db"]))[0][:, :128].copy()

    class Searcher:                                                            # stand-in for the device searcher (exact kNN by the oracle)
        def search_device(self, q_hat, k):
            i, d = oknn.search(db, q_hat.numpy(), k)
            return torch.from_numpy(i), torch.from_numpy(d)

        def gather_device(self, idx):
            return torch.from_numpy(db[idx.numpy()].astype(np.float32))
    model.retriever = types.SimpleNamespace(searcher=Searcher(), data_pool={"embedding": db})
    for tag, kw in gen.RARM_SAMPLING_CASES.items():
        np.random.seed(kw["seed"])
        torch.manual_seed(kw["seed"])
        logs = model.sample_from_rdata(2, qids=None, k_nn=4, memsize=100, top_k=kw["top_k"], temperature=kw["temperature"], code_side_len=3,
                                       z_dimensionality=8, guidance_scale=kw["guidance_scale"])
        assert sorted(logs.keys()) == ["qids", "samples_with_sampled_nns"]     # the reference's keys
        assert np.array_equal(np.asarray(logs["qids"]), g[f"{tag}:qids"])
        assert torch.equal(logs["samples_with_sampled_nns"], torch.from_numpy(g[f"{tag}:images"])), tag
    _, c = model.encode_to_c(torch.zeros((2, 0)))
    got = model.sample(torch.from_numpy(g["greedy:start"]), torch.from_numpy(g["greedy:r"]), c, steps=6, sample=False, top_k=None, guidance_scale=3.0)
    assert torch.equal(got, torch.from_numpy(g["greedy:tokens"]))


@pytest.mark.parametrize("schedule", ["reverse", "random:5"])
def test_results_do_not_depend_on_the_thread_scheduling_order(emu_cls, monkeypatch, schedule):
    """The emulator resumes the runnable threads of a block in index order by default; a missing barrier (a read-after-write or
    write-after-read hazard between threads) would make results depend on that order.  Same checks with the order reversed / shuffled."""
    monkeypatch.setenv("EMU_SCHEDULE", schedule)
    d, cfg, sd, net = small(emu_cls, 0)
    tok, ctx = torch.from_numpy(d["tokens"])[:, :4], torch.from_numpy(d["context"])
    assert rel(net.forward(tok, ctx), d["logits"][:, :4]) < 2e-6
    g = torch.Generator().manual_seed(1)
    lc, lu, u = torch.randn(2, 16384, generator=g) * 3, torch.randn(2, 16384, generator=g) * 3, torch.rand(2, generator=g)
    tokn, probs = net.sample_step(torch.cat([lc, lu]), guidance_scale=2.0, temperature=0.9, top_k=256, uniforms=u, want_probs=True)
    want = orarm.step_probs(lc, lu, 2.0, 0.9, 256)
    assert torch.equal(probs > 0, want > 0) and float((probs - want).abs().max()) < 1e-6 and torch.equal(tokn, orarm.draw(want, u))
    c = torch.full((2, 1), cfg["in_channels"] - 1)
    uu = torch.rand(4, 2, generator=g)
    net.set_context(torch.cat([ctx[:2], torch.zeros_like(ctx[:2])]))
    toks = net.sample(c, 4, temperature=0.9, top_k=6, guidance_scale=2.0, uniforms=uu)
    want_t, _ = orarm.sample(sd, cfg["n_heads"], c, torch.zeros((2, 0), dtype=torch.long), ctx[:2], 4, 0.9, 6, 2.0, uu)
    assert torch.equal(toks[:, 1:], want_t)
