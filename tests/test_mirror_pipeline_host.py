"""CPU: the product's `MinimalRETRODiffusion.sample_from_rdata` / `sample_with_query` orchestration (conditioning assembly, option
handling, EMA scope, sampler hand-off) against tests/golden/ref_pipeline_tiny.npz -- the REFERENCE's own methods run end to end on CPU
(tests/golden/make_golden_ref.py).  The device executors are replaced by oracle-backed stand-ins with the same call surface
(tests/test_reference_scripts.py), so what is compared is the host code of `rdm/models/diffusion/ddpm.py` + `ddim.py` of this repo."""
import copy
import os
import pickle
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import unet as ounet
from test_mirror_host import TINY_CFG
from test_reference_scripts import cpu_executors  # noqa: F401  (fixture)

GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLD)
import ref_weights  # noqa: E402


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm())


@pytest.fixture()
def model_and_golden(tmp_path, cpu_executors):  # noqa: F811
    import rdm  # noqa: F401
    from ldm.util import instantiate_from_config
    from omegaconf import OmegaConf
    from rdm_b200.unet import unet_param_shapes
    p = np.load(os.path.join(GOLD, "ref_pipeline_tiny.npz"))
    db, mem, id_count = ref_weights.make_db(int(p["n_db"]))
    np.savez(tmp_path / "db.npz", embedding=db, img_id=np.arange(len(db)), patch_coords=np.zeros((len(db), 4), np.int32))
    with open(tmp_path / "nn_memory.p", "wb") as f:
        pickle.dump({"nn_memory": mem, "id_count": id_count}, f)
    cfg = copy.deepcopy(TINY_CFG)
    cfg["params"]["nn_memory"] = str(tmp_path / "nn_memory.p")
    cfg["params"]["retrieval_cfg"]["params"]["saved_embeddings"] = str(tmp_path / "db.npz")
    model = instantiate_from_config(OmegaConf.create(cfg))
    shapes = unet_param_shapes(**ounet.TINY_UNET)
    live, ema = (ref_weights.state_dict_for(shapes.items(), int(p[s])) for s in ("live_seed", "ema_seed"))
    ck = {"model.diffusion_model." + k: v for k, v in live.items()}
    ck.update({"model_ema." + ("diffusion_model." + k).replace(".", ""): v for k, v in ema.items()})
    model.load_state_dict(ck, strict=False)
    return model.eval(), p


COMMON = dict(unconditional_guidance_scale=2.0, ddim_steps=4, ddim=True, unconditional_retro_guidance_label=0.)


def test_sample_from_rdata_matches_reference_code(model_and_golden):
    model, p = model_and_golden
    xT = torch.from_numpy(p["rdata:x_T"])
    np.random.seed(44)
    logs = model.sample_from_rdata(3, qids=None, k_nn=4, use_weights=False, memsize=50, x_T=xT.clone(), **COMMON)
    assert list(logs.keys()) == ["samples_with_sampled_nns"]
    assert list(logs["nns"][:, 0].numpy()) == list(p["rdata:qids"])
    assert rel(logs["samples_with_sampled_nns"], p["rdata:samples"]) < 1e-5


@pytest.mark.parametrize("tag,extra", [("query", dict(omit_query=False)), ("query_omit", dict(omit_query=True)), ("query_normalize", dict(normalize=True)),
                                       ("query_reps", dict(n_reps=2)), ("query_single", dict(bs=2, single=True))])
def test_sample_with_query_options_match_reference_code(model_and_golden, tag, extra):
    model, p = model_and_golden
    extra = dict(extra)
    q = torch.from_numpy(p["query:q"])
    q = q[:1] if extra.pop("single", False) else q
    xT = torch.from_numpy(p["rdata:x_T"])[:2]
    logs = model.sample_with_query(query=q, query_embedded=True, k_nn=4, visualize_nns=False, x_T=xT.clone(), **COMMON, **extra)
    assert list(logs.keys()) == ["query_samples"]
    assert rel(logs["query_samples"], p[f"{tag}:samples"]) < 1e-5, tag


def test_get_nn_and_encoding_matches_reference_code(model_and_golden):
    """ddpm.py:263-316: image -> n x n patches -> retriever -> q / |q| -> exact kNN -> RAW neighbour rows [b, n*n, k, d] (the per-step
    re-retrieval of BASELINE cfg4); channel-first and channel-last inputs."""
    import retro_stub
    model, p = model_and_golden
    model.retriever._retriever = retro_stub.PatchEmbedStub()
    imgs = torch.from_numpy(p["nnenc:images"])
    for tag, x, n in (("nnenc_2x2", imgs, 2), ("nnenc_1x1_channels_last", imgs.permute(0, 2, 3, 1).contiguous(), 1)):
        r = model.get_nn_and_encoding(x, k_nn=3, n_patches_per_side=n)
        want = p[f"{tag}:nn_embeddings"]
        assert tuple(r[model.nn_key].shape) == want.shape == (2, n * n, 3, 512)
        assert np.array_equal(r[model.nn_key].numpy(), want), tag


def test_dataset_builder_search_matches_reference_code(tmp_path, cpu_executors):  # noqa: F811
    """tests/golden/ref_dsetbuilder.npz: the REFERENCE's DatasetBuilder.load_embeddings / train_searcher / embed / search_k_nearest over
    an exact stand-in for scann's brute-force scorer (the reference normalises the fp16 rows itself, dsetbuilder.py:574).  The product's
    DatasetBuilder must return the same neighbours and the same result arrays -- for embedded queries and for channel-last image patches."""
    import retro_stub
    import rdm  # noqa: F401
    from rdm.data.retrieval_dataset.dsetbuilder import DatasetBuilder
    g = np.load(os.path.join(GOLD, "ref_dsetbuilder.npz"))
    n = int(g["n_db"])
    db, _, _ = ref_weights.make_db(n)
    np.savez(tmp_path / "db.npz", embedding=db, img_id=np.arange(n) * 3, patch_coords=np.stack([np.arange(n)] * 4, 1).astype(np.int32))
    b = DatasetBuilder(retriever_config=None, saved_embeddings=str(tmp_path / "db.npz"), load_patch_dataset=False, gpu=False, k=5, max_pool_size=1000)
    b._retriever = retro_stub.PatchEmbedStub()
    b.train_searcher()
    r = b.search_k_nearest(g["embedded:queries"], k=5, query_embedded=True)
    assert np.array_equal(r["nns"], g["embedded:nns"]) and r["nns"][0, 0] == 17
    for key in ("embeddings", "img_ids", "patch_coords", "q_embeddings"):
        assert np.array_equal(np.asarray(r[key]), g[f"embedded:{key}"]), key
    r = b.search_k_nearest(torch.from_numpy(g["images:queries"]), k=4, is_caption=False)
    assert np.array_equal(r["nns"], g["images:nns"])
    for key in ("embeddings", "img_ids", "patch_coords"):
        assert np.array_equal(np.asarray(r[key]), g[f"images:{key}"]), key
    assert np.allclose(np.asarray(r["q_embeddings"]), g["images:q_embeddings"], rtol=1e-6, atol=1e-6)
