"""CPU: the reference's OWN sampling scripts, unchanged (`/root/reference/scripts/rdm_sample.py`, `scripts/rarm_sample.py`), run against the
drop-in packages of this repository: `parse_args` -> `load_model` (the SHIPPED `models/**/config.yaml` with sizes reduced, a synthetic
Lightning checkpoint in the reference's key layout, `load_state_dict(strict=False)`, `.eval()`) -> `sample_unconditional` /
`sample_conditional` / `sample` -> PNG files.  Because this container has no GPU the device executors are either oracle-backed stand-ins with the same call
surface (searcher, U-Net engine, RARM decoder, first-stage decoder) or -- second parametrisation of both tests -- the
product's OWN kNN / U-Net / DDIM / RARM source executed under the host emulation of CUDA (tests/emu/); everything else -- configuration handling,
class resolution through the YAML `target:` strings, checkpoint loading, conditioning assembly, the sampler, the first-stage
containers, the returned dictionaries the scripts iterate -- is the product's host code.  Skipped where /root/reference is absent
(the GPU box)."""
import importlib.util
import os
import pickle
import sys

import numpy as np
import pytest
import torch
import yaml

from conftest import ROOT
from oracle import knn as oknn, rarm as orarm, unet as ounet, vqdecoder as ovq

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLD)
import ref_weights  # noqa: E402

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "scripts")), reason="needs the reference checkout (build container only)")


def load_script(name):
    import rdm  # noqa: F401  (the drop-in package; installs the stand-ins for omegaconf / pytorch_lightning / clip / ldm / taming when missing)
    spec = importlib.util.spec_from_file_location("ref_script_" + name, os.path.join(REF, "scripts", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def plain(cfg):
    from omegaconf import OmegaConf
    return OmegaConf.to_container(cfg)


# ---- oracle-backed stand-ins for the device executors --------------------------------------------------------------------------
class CpuSearcher:
    """rdm_b200.knn.B200Searcher's call surface over oracle/knn (exact cosine top-k, raw-row gather)."""

    def __init__(self, embedding, device=None, idx_base=0):
        self.db, self.device, self.local = np.ascontiguousarray(embedding), torch.device("cpu"), self

    def search_device(self, q_hat, k, return_scores=False):
        idx, dist = oknn.search(self.db, q_hat.numpy(), k)
        return torch.from_numpy(idx), torch.from_numpy(dist)

    def gather_device(self, idx):
        return torch.from_numpy(self.db[idx.numpy()].astype(np.float32))

    def search_batched(self, q, final_num_neighbors=None, **_):
        return oknn.search(self.db, np.asarray(q, np.float32), int(final_num_neighbors))


class CpuUNetEngine:
    """rdm_b200.unet.B200UNet's call surface over oracle/unet + the DDIM update from the same coefficient tables."""

    def __init__(self, cfg, sd):
        self.net = ounet.UNetModel(**cfg).eval()
        self.net.load_state_dict({k: v for k, v in sd.items()})

    def set_context(self, ctx):
        self.ctx = ctx

    @torch.no_grad()
    def ddim_sample(self, x, timesteps, coef, cfg_scale=1.0, first_step=0, num_steps=None, noise=None, want_pred_x0=False):
        n = len(timesteps) - first_step if num_steps is None else num_steps
        b, p0 = x.shape[0], None
        for i in range(first_step, first_step + n):
            t = torch.full((self.ctx.shape[0],), int(timesteps[i]), dtype=torch.long)
            out = self.net(torch.cat([x] * (self.ctx.shape[0] // b)), t, self.ctx)
            e = out[b:] + cfg_scale * (out[:b] - out[b:]) if cfg_scale > 1.0 else out
            c0, c1, c2, c3 = (coef[i, j] for j in range(4))
            p0 = (x - c0 * e) / c1
            x = c2 * p0 + c3 * e
        return (x, p0) if want_pred_x0 else x


class CpuRarmEngine:
    def __init__(self, sd, heads):
        self.sd, self.heads = sd, heads

    def set_context(self, r):
        self.r = r

    def sample(self, prefix, steps, temperature=1.0, top_k=None, guidance_scale=1.0, uniforms=None):
        B = prefix.shape[0]
        toks, _ = orarm.sample(self.sd, self.heads, prefix[:, :1], prefix[:, 1:], self.r[:B], steps, temperature, top_k, guidance_scale, uniforms)
        return torch.cat([prefix[:, :1], toks], 1)


@pytest.fixture()
def emulated_executors(monkeypatch):
    """The product's own kNN / U-Net / DDIM / RARM source (csrc/knn.cu, unet.cu, kernels.cu, gemm_simt.cu, rarm.cu) under the host emulation of CUDA
    (tests/emu/) instead of oracle stand-ins; only the tensor-core-only first-stage decoder keeps an oracle stand-in."""
    import contextlib
    import ctypes
    import shutil
    if shutil.which("g++") is None:
        pytest.skip("needs g++ to build the emulated library")
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    import rdm  # noqa: F401
    from ldm.models.autoencoder import VQModelInterface
    from rdm_b200 import _lib
    os.environ["RDM_KNN_NO_TC"] = "1"
    names = [n for n in _lib.SIGNATURES if n.startswith(("rdm_unet_", "rdm_ddim_", "rdm_knn_", "rdm_rarm_"))] + ["rdm_last_error", "rdm_launch_count"]
    monkeypatch.setattr(_lib, "_lib", _lib.bind(ctypes.CDLL(build_emu.build()), names))
    monkeypatch.setattr(_lib, "resolve_device", lambda d: torch.device("cpu"))
    monkeypatch.setattr(_lib, "device_ctx", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(_lib, "stream_ptr", lambda d=None: None)
    monkeypatch.setenv("RDM_B200_MODE", "fp32")                                # strict mode: the tensor-core modes need hardware
    monkeypatch.setenv("RDM_B200_RARM_MODE", "fp32")

    def first_stage_decode(self, h, force_not_quantize=False):
        ref = ovq.VQModelInterface(self.embed_dim, self.quantize.embedding.num_embeddings, self._ddconfig).eval()
        ref.load_state_dict(self.state_dict())
        with torch.no_grad():
            return ref.decode(h, force_not_quantize)
    monkeypatch.setattr(VQModelInterface, "decode", first_stage_decode)


@pytest.fixture()
def cpu_executors(monkeypatch):
    import rdm  # noqa: F401
    import rdm.data.retrieval_dataset.dsetbuilder as dsb
    from rdm.modules.attention import RetrievalPatchTransformer
    from rdm.modules.diffusionmodules.openaimodel import UNetModel
    monkeypatch.setattr(dsb, "B200Searcher", CpuSearcher)

    def unet_set_context(self, context, device=None):
        context = context[0] if isinstance(context, (list, tuple)) else context
        sd = self._weight_override[1] if self._weight_override is not None else self.state_dict()
        eng = CpuUNetEngine(dict(self._cfg, image_size=self.image_size), sd)
        eng.set_context(context)
        return eng
    monkeypatch.setattr(UNetModel, "set_context", unet_set_context)
    monkeypatch.setattr(RetrievalPatchTransformer, "engine", lambda self, device: CpuRarmEngine(self.state_dict(), self._cfg["n_heads"]))
    from ldm.models.autoencoder import VQModelInterface

    def first_stage_decode(self, h, force_not_quantize=False):                 # rdm_b200.vqdecoder.B200VQDecoder's role, by the oracle decoder
        ref = ovq.VQModelInterface(self.embed_dim, self.quantize.embedding.num_embeddings, self._ddconfig).eval()
        ref.load_state_dict(self.state_dict())
        with torch.no_grad():
            return ref.decode(h, force_not_quantize)
    monkeypatch.setattr(VQModelInterface, "decode", first_stage_decode)


def make_db(tmp_path, n=400):
    db, mem, id_count = ref_weights.make_db(n)
    os.makedirs(tmp_path / "database", exist_ok=True)
    np.savez(tmp_path / "database" / "part0.npz", embedding=db[:150], img_id=np.arange(150), patch_coords=np.zeros((150, 4), np.int32))
    np.savez(tmp_path / "database" / "part1.npz", embedding=db[150:], img_id=np.arange(150, n), patch_coords=np.zeros((n - 150, 4), np.int32))
    with open(tmp_path / "nn_memory.p", "wb") as f:
        pickle.dump({"nn_memory": mem, "id_count": id_count}, f)
    return db


@pytest.mark.parametrize("executors", ["cpu_executors", "emulated_executors"])
def test_rdm_sample_script_runs_unchanged(tmp_path, monkeypatch, request, executors):
    request.getfixturevalue(executors)                                         # oracle stand-ins, or the product's own source under emulation
    from omegaconf import OmegaConf
    script = load_script("rdm_sample")
    db = make_db(tmp_path)
    # the SHIPPED config with reduced sizes: every key of models/rdm/imagenet/config.yaml reaches the product's constructors
    cfg = OmegaConf.load(os.path.join(REF, "models", "rdm", "imagenet", "config.yaml"))
    p = cfg.model.params
    p.image_size, p.nn_memory = 16, str(tmp_path / "nn_memory.p")
    p.unet_config.params.update(dict(image_size=16, model_channels=64, attention_resolutions=[2, 4], num_res_blocks=1, channel_mult=[1, 2, 3]))
    p.first_stage_config.params.update(dict(n_embed=64))
    p.first_stage_config.params.ddconfig.update(dict(resolution=32, ch=32, ch_mult=[1, 2], num_res_blocks=1))
    p.retrieval_cfg.params.saved_embeddings = str(tmp_path / "database")
    model_dir = tmp_path / "model"
    model_dir.mkdir()
    with open(model_dir / "config.yaml", "w") as f:
        yaml.safe_dump(plain(cfg), f)
    # a Lightning checkpoint in the reference's layout: live + EMA U-Net, first stage, schedule buffers; no guidance vector
    ucfg = {k: v for k, v in plain(p.unet_config.params).items()}
    unet = ounet.randomize_(ounet.UNetModel(**ucfg), 1)
    ema = ounet.randomize_(ounet.UNetModel(**ucfg), 2)
    fs = ovq.randomize_(ovq.VQModelInterface(**{k: v for k, v in plain(p.first_stage_config.params).items()}), 3)
    sd = {"model.diffusion_model." + k: v for k, v in unet.state_dict().items()}
    sd.update({"model_ema." + ("diffusion_model." + k).replace(".", ""): v for k, v in ema.state_dict().items()})
    sd.update({"model_ema.decay": torch.tensor(0.9999), "model_ema.num_updates": torch.tensor(10, dtype=torch.int)})
    sd.update({"first_stage_model." + k: v for k, v in fs.state_dict().items()})
    torch.save({"state_dict": sd, "global_step": 1}, model_dir / "model.ckpt")

    out = tmp_path / "out"
    bs = 2 if executors == "cpu_executors" else 1                              # (the emulated kernels are slow: one image is enough there)
    steps = 4 if executors == "cpu_executors" else 2
    argv = ["rdm_sample.py", "-s", str(out), "--model_path", str(model_dir), "-bs", str(bs), "--gpu", "-1", "--n_runs", "1", "--steps", str(steps), "--k_nn", "4",
            "--guidance_scale", "2.0", "--top_m", "0.5"]
    monkeypatch.setattr(sys, "argv", argv)
    opt = script.parse_args()
    opt.savepath.mkdir(parents=True, exist_ok=True)
    model = script.load_model(opt)                                             # unchanged reference code from here on
    assert type(model).__module__ == "rdm.models.diffusion.ddpm" and "retrieval-augmented-diffusion-models_b200" in sys.modules[type(model).__module__].__file__
    np.random.seed(0)
    torch.manual_seed(0)
    script.sample_unconditional(model, opt)
    pngs = sorted(os.listdir(out))
    assert len(pngs) == bs and all("samples_with_sampled_nns" in f for f in pngs)         # exactly the files the reference writes
    from PIL import Image
    assert Image.open(out / pngs[0]).size == (32, 32)
    # what was saved is the EMA-weight, retrieval-conditioned sample: recompute sample 0 with the oracle pipeline
    np.random.seed(0)
    torch.manual_seed(0)
    logs = model.sample_from_rdata(bs, qids=None, k_nn=4, use_weights=False, memsize=0.5, unconditional_guidance_scale=2.0, ddim_steps=steps, ddim=True,
                                   unconditional_retro_guidance_label=0.)
    assert list(logs.keys()) == ["samples_with_sampled_nns"]                   # the reference's keys only (extras: logs.extras)
    nns = logs.extras["nns"].numpy()
    qh = oknn.normalize_queries(db[nns[:, 0]].astype(np.float32))
    assert np.array_equal(oknn.search(db, qh, 4)[0], nns)
    # the text-conditioned path of the script: tokenize -> retriever CLIP -> sample_with_query; the CLIP tower itself is a device executor,
    # so a fixed embedding stands in for `clip.encode_text`
    class FakeClip:
        def encode_text(self, tokens):
            assert tokens.shape == (bs, 77) and tokens.dtype in (torch.int64, torch.int32)
            return torch.from_numpy(ref_weights.tensor_for("caption", (bs, 512), 5) * 22.0)
    model.retriever._retriever = type("R", (), {"model": FakeClip(), "to": lambda self, d: self})()
    argv2 = argv + ["-c", "a corgi wearing a hat", "--omit_query"]
    monkeypatch.setattr(sys, "argv", argv2)
    opt2 = script.parse_args()
    out2 = tmp_path / "out2"
    out2.mkdir()
    opt2.savepath = out2
    script.sample_conditional(model, opt2)
    pngs = sorted(os.listdir(out2))
    assert len(pngs) == bs and all("query_samples" in f for f in pngs)


@pytest.mark.parametrize("executors", ["cpu_executors", "emulated_executors"])
def test_rarm_sample_script_runs_unchanged(tmp_path, monkeypatch, request, executors):
    request.getfixturevalue(executors)
    from omegaconf import OmegaConf
    script = load_script("rarm_sample")
    make_db(tmp_path)
    cfg = OmegaConf.load(os.path.join(REF, "models", "rarm", "imagenet", "dogs", "config.yaml"))
    p = cfg.model.params
    p.nn_memory = str(tmp_path / "nn_memory.p")
    p.mask_token, p.sos_token = 64, 65
    p.transformer_config.params.update(dict(in_channels=66, n_heads=2, depth=1, out_channels=64))     # sequence_length 256 stays: 16 x 16 codes
    p.first_stage_config.params.update(dict(n_embed=64))                      # embed_dim / z_channels stay 256: the script relies on z_dimensionality=256
    p.first_stage_config.params.ddconfig.update(dict(resolution=32, ch=32, ch_mult=[1, 2], num_res_blocks=1, attn_resolutions=[16]))
    p.retrieval_cfg.params.saved_embeddings = str(tmp_path / "database")
    model_dir = tmp_path / "model"
    model_dir.mkdir()
    with open(model_dir / "config.yaml", "w") as f:
        yaml.safe_dump(plain(cfg), f)
    tshapes = orarm.param_shapes(**plain(p.transformer_config.params))
    fs = ovq.randomize_(ovq.VQModelInterface(embed_dim=256, n_embed=64, ddconfig=plain(p.first_stage_config.params.ddconfig)), 3)
    sd = {"transformer." + k: v for k, v in ref_weights.state_dict_for(tshapes.items(), 9).items()}
    sd.update({"first_stage_model." + k: v for k, v in fs.state_dict().items()})
    sd.update({"sos_token": torch.LongTensor([65]), "mask_token": torch.LongTensor([64])})
    torch.save({"state_dict": sd, "global_step": 1}, model_dir / "model.ckpt")
    out = tmp_path / "out"
    bs = 2 if executors == "cpu_executors" else 1
    argv = ["rarm_sample.py", "-s", str(out), "--model_path", str(model_dir), "-bs", str(bs), "--gpu", "-1", "--n_runs", "1", "--k_nn", "4", "--top_k", "16",
            "--guidance_scale", "2.0", "--top_m", "0.5"]
    monkeypatch.setattr(sys, "argv", argv)
    opt = script.parse_args()
    opt.savepath.mkdir(parents=True, exist_ok=True)
    model = script.load_model(opt)
    assert type(model).__module__ == "rdm.models.autoregression.transformer"
    np.random.seed(0)
    torch.manual_seed(0)
    script.sample(model, opt)                                                  # 256 tokens per image, then the VQGAN-layout first stage
    pngs = sorted(os.listdir(out))
    assert len(pngs) == bs and all("samples_with_sampled_nns" in f for f in pngs)
    from PIL import Image
    assert Image.open(out / pngs[0]).size == (32, 32)
