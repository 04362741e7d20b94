#!/usr/bin/env python
"""Headline benchmark: images/sec of the retrieval-augmented DDIM sampler (BASELINE.json cfg2).

One "step" = one pass of the hot path over one batch per GPU:
    kNN (16 queries, k=4, exact cosine over the fp16 CLIP database in HBM) -> gather raw neighbour rows ->
    cross-attention K/V projection of [cond | uncond=0] -> DDIM-100 with classifier-free guidance 2.0 over the
    ImageNet-RDM U-Net (mc 192, mult 1-2-3-5, 16 SpatialTransformers) on the 32x32x4 latent -> 16 latents.
Synthetic data / random-init weights of the named architecture (no network for checkpoints or databases).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode fp16|fp16x2|bf16x3|bf16|fp32] [--chains C]

* `value`: the step with its inputs (query ids, x_T) already resident in HBM, through the library wrappers (rdm_b200.*).
* `e2e`: the SAME step through the reference-facing drop-in API -- `rdm.models.diffusion.ddpm.MinimalRETRODiffusion` instantiated from a
  YAML-style config, `.sample_from_rdata(N, qids=<host numpy>, k_nn=4, unconditional_guidance_scale=2., ddim_steps=100, ddim=True,
  unconditional_retro_guidance_label=0., x_T=<pinned host tensor>)` (scripts/rdm_sample.py:241-252: EMA scope, DDIMSampler with its per-step
  RNG draws) -- with host->device copies of the inputs and a device->host read of the latents inside the timed region.
* N = 1: the 1,281,167 x 512 fp16 database of cfg2, resident on the GPU.  A second line of numbers (`r_shape`) times the reference's SHIPPED
  shape (64x64x3 latent, models/rdm/imagenet/config.yaml:14-59) the same way.
* N > 1 (torchrun): the cfg3 data path is part of the timed step: the 20,927,907 x 512 fp16 OpenImages-size database is ROW-SHARDED N ways;
  every rank brings its own 16 queries; all_gather(queries) -> local exact scan -> all_to_all(packed idx|score lists) -> on-device merge ->
  neighbour rows fetched from their owners (all_gather idx, local gather, reduce_scatter) -> DDIM-100.  Weak scaling (16 images per GPU) is the
  headline; `strong` repeats the step with the global batch fixed at 64.  Before timing, the sharded neighbours are asserted bit-identical to an
  unsharded search of the whole database on rank 0.  Timing = max over ranks of CUDA-event time.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "retrieval-augmented-diffusion-models_b200")]

import numpy as np
import torch

METRIC = "images/sec (256^2, DDIM-100, k=4)"
UNET = dict(image_size=32, in_channels=4, out_channels=4, model_channels=192, attention_resolutions=[8, 4, 2], num_res_blocks=2,
            channel_mult=[1, 2, 3, 5], num_head_channels=32, transformer_depth=1, context_dim=512)
UNET_R = dict(UNET, image_size=64, in_channels=3, out_channels=3)          # models/rdm/imagenet/config.yaml:36-59 (the shipped shape)
N_DB, D, K_NN, BATCH, S_DDIM, CFG_SCALE = 1_281_167, 512, 4, 16, 100, 2.0
N_DB_SHARDED = 20_927_907                                                  # OpenImages database (scripts/download_databases.sh:6-15)
STRONG_GLOBAL_BATCH = 64                                                   # cfg3's batch
FLOP_PER_FWD_SAMPLE = 50.56e9          # SURVEY.md section 6 (32x32x4 latent, k=4)
FLOP_PER_FWD_SAMPLE_R = 208.56e9       # 64x64x3 latent
DEFAULT_MODE, DEFAULT_CHAINS = "fp16", 1
MODES = {"fp32": 0, "bf16x3": 1, "bf16": 2, "fp16x2": 3, "fp16": 4}
DB_CHUNK = 1 << 20                     # the synthetic databases are defined chunk by chunk, so a shard is the same rows for any world size


def peaks():
    p = dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, src="fallback")
    try:
        j = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        p.update(hbm_gbs=j["hbm_gbs"], bf16_tflops=j["bf16_tflops"], bf16_tflops_sustained=j.get("bf16_tflops_sustained", j["bf16_tflops"]), src="measured")
    except Exception:
        pass
    return p


def traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum over the tcgen05 GEMM launches of one forward, from the committed ncu pass of this round."""
    for name in ("gemm_tc_dram_r2zb.json", "gemm_tc_dram_r2z.json"):          # newest capture of this round's kernels first
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            break
    try:
        j = json.load(open(path))
        return j["bytes_per_forward"], f"profiles/{name} ({j['how']})"
    except Exception:
        return None, "no ncu capture of this round's kernels committed"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {getattr(nv, n): n[len("nvmlClocksEventReason"):] for n in dir(nv) if n.startswith("nvmlClocksEventReason") and isinstance(getattr(nv, n), int)}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if bit and (r & bit) and nm not in ("GpuIdle", "None", "All"):
                        self.reasons.add(nm)
                time.sleep(0.1)
        except Exception as e:           # NVML missing: report that instead of inventing numbers
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        s = sorted(self.samples)
        return dict(sm_mhz=s[len(s) // 2] if s else None, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(s))


def make_weights(seed=3, cfg=None):
    """Random-init weights of the named architecture, generated directly as a state dict (no nn.Module needed)."""
    from rdm_b200.unet import unet_param_shapes
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shp in unet_param_shapes(**(cfg or UNET)).items():
        if len(shp) >= 2:
            fan_in = int(np.prod(shp[1:]))
            sd[name] = torch.randn(shp, generator=g) / fan_in ** 0.5
        elif name.endswith("weight"):
            sd[name] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            sd[name] = 0.05 * torch.randn(shp, generator=g)
    return sd


def synthetic_db_rows(lo, hi, dev):
    """Rows [lo, hi) of the synthetic fp16 database: chunk c (DB_CHUNK rows) = randn(seed 7000 + c) -- independent of how it is sharded."""
    out = torch.empty((hi - lo, D), dtype=torch.float16, device=dev)
    c = lo // DB_CHUNK
    while c * DB_CHUNK < hi:
        g = torch.Generator(device=dev).manual_seed(7000 + c)
        rows = torch.randn((DB_CHUNK, D), generator=g, device=dev, dtype=torch.float32).to(torch.float16)
        a, b = max(lo, c * DB_CHUNK), min(hi, (c + 1) * DB_CHUNK)
        out[a - lo:b - lo] = rows[a - c * DB_CHUNK:b - c * DB_CHUNK]
        c += 1
    return out


# ------------------------------------------------------------------------------------------------ reference arm / CPU baseline
def cpu_reference_images_per_sec(ddim_steps, reps=1, threads=None):
    """The reference's own code cannot be installed here (ldm/scann/clip/... absent), so the CPU arm is
    the oracle port: torch-CPU fp32 U-Net + DDIM with CFG for ONE image (B2 = 2) over `ddim_steps` steps, plus the C kNN
    oracle for that image's query over a 100K-row slice, extrapolated to DDIM-100 and the full DB."""
    from oracle import ddim as oddim, knn as oknn, unet as ounet
    if threads:
        torch.set_num_threads(threads)
    net = ounet.UNetModel(**UNET).eval()
    net.load_state_dict(make_weights())
    rng = np.random.default_rng(1)
    db = rng.standard_normal((100_000, D)).astype(np.float16)
    qh = oknn.normalize_queries(db[:1].astype(np.float32))
    inv = oknn.inv_norms(db)
    x = torch.randn(1, 4, 32, 32, generator=torch.Generator().manual_seed(0))
    sch = oddim.Schedule(S_DDIM)
    times = []
    for _ in range(reps):
        t0 = time.time()
        idx, _ = oknn.search(db, qh, K_NN, inv=inv)
        t_knn = (time.time() - t0) * (N_DB / db.shape[0])
        cond = torch.from_numpy(db[idx].astype(np.float32))
        t1 = time.time()
        with torch.no_grad():
            xx = x
            for i in range(ddim_steps):
                ts = torch.full((2,), int(np.flip(sch.timesteps)[i]))
                out = net(torch.cat([xx] * 2), ts, torch.cat([cond, torch.zeros_like(cond)]))
                e = out[1:] + CFG_SCALE * (out[:1] - out[1:])
                xx, _ = oddim.ddim_update(xx, e, *sch.coeffs(S_DDIM - i - 1))
        t_step = (time.time() - t1) / ddim_steps
        times.append(t_knn + t_step * S_DDIM)
    t_img = float(np.median(times))
    return 1.0 / t_img, dict(sec_per_image=t_img, sec_per_ddim_step=t_step, knn_sec_extrapolated=t_knn)


def run_reference(args, out=sys.stdout):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    os.environ["OMP_NUM_THREADS"] = str(cores)          # torchrun pins it to 1; the reference arm may use every host thread
    torch.set_num_threads(cores)
    vals = []
    for _ in range(args.warmup):
        cpu_reference_images_per_sec(1)
    t0 = time.time()
    for _ in range(args.steps):
        v, info = cpu_reference_images_per_sec(2)
        vals.append(v)
    ms = (time.time() - t0) * 1e3 / max(1, args.steps)
    v = float(np.median(vals))
    sample = "per step: 2 DDIM steps (CFG, B2=2) of 1 image on the torch-CPU oracle + C kNN oracle over 100K rows, EXTRAPOLATED to DDIM-100 / 1.28M rows"
    print(file=out, flush=True, *[json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "cfg2 RDM ImageNet-arch 32x32x4, DDIM-100, CFG 2.0, k=4 over 1,281,167x512 fp16 DB (CPU oracle port; reference not installable)",
                                 "extrapolated": True, "extrapolation": "value = 1 / (100 x measured seconds per guided DDIM step of one image + C kNN oracle seconds over 100K rows x 12.8); "
                                                                        "the reference cannot run this path on a CUDA-less host at all (DDIMSampler.register_buffer forces cuda, SURVEY F6)"},
                      "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})])


# ------------------------------------------------------------------------------------------------ the drop-in model (reference API)
def dropin_config(unet_cfg, mode):
    """YAML-style config of the shipped model (models/rdm/imagenet/config.yaml:1-106) with the benchmark's U-Net; no first stage: latents
    count as images, the retrieval database is attached on the device after construction."""
    return {"target": "rdm.models.diffusion.ddpm.MinimalRETRODiffusion",
            "params": {"k_nn": K_NN, "query_key": "clip_img_emb", "linear_start": 0.0015, "linear_end": 0.0195, "num_timesteps_cond": 1, "log_every_t": 200,
                       "timesteps": 1000, "first_stage_key": "image", "cond_stage_key": "nixda", "image_size": unet_cfg["image_size"], "channels": unet_cfg["in_channels"],
                       "cond_stage_trainable": False, "nn_key": "nn_embeddings", "conditioning_key": "retro_only", "monitor": "val/loss_simple_ema",
                       "scale_by_std": False, "ignore_keys": ["unconditional_guidance_vex"],
                       "unet_config": {"target": "rdm.modules.diffusionmodules.openaimodel.UNetModel",
                                       "params": dict(unet_cfg, use_spatial_transformer=True, use_checkpoint=True, use_scale_shift_norm=False, resblock_updown=False)},
                       "first_stage_config": None,
                       "retrieval_cfg": {"target": "rdm.data.retrieval_dataset.dsetbuilder.DatasetBuilder",
                                         "params": {"patch_size": 256, "batch_size": 100, "k": 20, "max_pool_size": 20000000, "gpu": True, "load_patch_dataset": False,
                                                    "retriever_config": None, "saved_embeddings": None}},
                       "retrieval_encoder_cfg": {"target": "torch.nn.Identity"}, "cond_stage_config": "__is_unconditional__"}}


def build_dropin(unet_cfg, sd, searcher, mode, chains, dev):
    import rdm  # noqa: F401  (installs the stand-ins for ldm / omegaconf / pytorch_lightning when they are absent)
    from ldm.util import instantiate_from_config
    from omegaconf import OmegaConf
    with contextlib.redirect_stdout(io.StringIO()):
        model = instantiate_from_config(OmegaConf.create(dropin_config(unet_cfg, mode)))
        ck = {"model.diffusion_model." + k: v for k, v in sd.items()}
        ck.update({"model_ema." + ("diffusion_model." + k).replace(".", ""): v for k, v in sd.items()})       # sampling uses the EMA copy (ddpm.py:977)
        model.load_state_dict(ck, strict=False)                                                                # scripts/rdm_sample.py:170
        model = model.eval().to(dev)
        model.model.diffusion_model.engine_mode = mode
        model.model.diffusion_model.engine_chains = chains
        model.retriever.searcher = searcher                                                                    # the database is already resident in HBM
    return model


def claim_stdout():
    """stdout carries exactly ONE JSON line: whatever else is printed there by libraries during the run (NCCL's version banner at
    communicator creation, progress lines of the reference-facing API) is routed to stderr; returns the real stdout for the final line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--mode", default=DEFAULT_MODE, choices=list(MODES),
                    help="U-Net contraction mode; DDIM-100 batch-16 final-latent rel-L2 vs the fp32 oracle is pinned per mode by "
                         "tests/test_zx_benchmarked_config_gpu.py (tolerance 1e-3)")
    ap.add_argument("--chains", type=int, default=DEFAULT_CHAINS, help="concurrent batch chains of the U-Net executor (rdm_unet_set_chains)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-r-shape", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    args = ap.parse_args()
    out = claim_stdout()
    if args.impl == "reference":
        return run_reference(args, out)

    import torch.distributed as dist
    from rdm_b200 import _lib, sampler
    from rdm_b200.knn import B200Searcher, ShardedSearcher, shard_range
    from rdm_b200.unet import B200UNet

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mode = MODES[args.mode]
    quiet = lambda: contextlib.redirect_stdout(io.StringIO())            # the reference API prints progress lines; stdout carries ONE JSON line

    # ---- resident state: database (N = 1: cfg2's 1.28 M rows; N > 1: this rank's rows of the 20.9 M-row database), weights, schedule tables
    n_db = N_DB if world == 1 else N_DB_SHARDED
    lo, hi = shard_range(n_db, rank, world)
    t_load = time.time()
    db = synthetic_db_rows(lo, hi, dev)
    local_searcher = B200Searcher(db, device=dev, idx_base=lo)
    searcher = ShardedSearcher(local_searcher, validate=False) if world > 1 else local_searcher      # equal query counts are guaranteed by construction here
    torch.cuda.synchronize()
    t_load = time.time() - t_load
    sd = make_weights()
    unet = B200UNet(dev, **UNET)
    unet.load_state_dict(sd)
    unet.set_mode(mode)
    unet.set_chains(args.chains)
    tables = sampler.make_ddim_tables(sampler.alphas_cumprod_linear(), S_DDIM, 0.0, device=dev)
    model = build_dropin(UNET, sd, searcher, args.mode, args.chains, dev)
    nbatches = args.steps + args.warmup
    rng = np.random.default_rng(2 + rank)
    pin = lambda t: t.pin_memory()
    qids_np = [rng.integers(0, n_db, size=BATCH) for _ in range(nbatches)]
    h_xT = [pin(torch.randn(BATCH, 4, 32, 32, generator=torch.Generator().manual_seed(100 * rank + i))) for i in range(nbatches)]
    h_out = pin(torch.empty(BATCH, 4, 32, 32))
    uncond = torch.zeros(BATCH, K_NN, D, device=dev)          # unconditional_retro_guidance_label = 0 (rdm_sample.py:251, ddpm.py:673-680)

    def retrieve(qids_dev, bs_uncond):
        q = searcher.gather_device(qids_dev)                                  # query = DB rows (ddpm.py:897); sharded: fetched from the owner rank
        nns, _ = searcher.search_raw_device(q, K_NN)                          # q / ||q|| (in the library) + search_batched   ddpm.py:906-908
        cond = searcher.gather_device(nns)                                    # ddpm.py:921 (raw rows, fp32)
        return nns, torch.cat([cond, bs_uncond])                              # cat([c, uc]) ddim.py:232

    def one_batch(qids_dev, xT_dev, net=unet, unc=uncond, tb=tables):
        nns, ctx = retrieve(qids_dev, unc)
        net.set_context(ctx)
        return net.ddim_sample(xT_dev, tb["timesteps"], tb["coef"], cfg_scale=CFG_SCALE)

    def e2e_batch(i, mdl=model, xs=h_xT, qs=qids_np, out=h_out):
        with quiet():
            logs = mdl.sample_from_rdata(xs[i].shape[0], qids=qs[i], k_nn=K_NN, use_weights=False, memsize=100, unconditional_guidance_scale=CFG_SCALE,
                                         ddim_steps=S_DDIM, ddim=True, unconditional_retro_guidance_label=0., x_T=xs[i].to(dev, non_blocking=True))
        out.copy_(logs["samples_with_sampled_nns"], non_blocking=True)        # device -> pinned host read of the step's result
        return logs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    d_q = [torch.from_numpy(q).to(dev) for q in qids_np]
    d_x = [t.to(dev) for t in h_xT]

    # ---- N > 1: the sharded neighbours must be bit-identical to an unsharded search of the whole database (asserted in-run, before timing)
    shard_check = None
    if world > 1:
        nns_mine, _ = retrieve(d_q[0], uncond)
        allq = [torch.empty_like(d_q[0]) for _ in range(world)]
        alln = [torch.empty_like(nns_mine) for _ in range(world)]
        dist.all_gather(allq, d_q[0]); dist.all_gather(alln, nns_mine)
        if rank == 0:
            full = B200Searcher(synthetic_db_rows(0, n_db, dev), device=dev)
            qa = torch.cat(allq)
            want, _ = full.search_raw_device(full.gather_device(qa), K_NN)
            same = bool(torch.equal(want, torch.cat(alln)))
            self_first = bool(torch.equal(want[:, 0], qa))
            del full
            torch.cuda.empty_cache()
            assert same and self_first, "row-sharded kNN differs from the unsharded search"
            shard_check = {"queries": int(qa.numel()), "k": K_NN, "bit_identical_to_unsharded": same, "rows": n_db}
        dist.barrier()

    for i in range(args.warmup):
        one_batch(d_q[i], d_x[i])
        e2e_batch(i)
    clk = ClockSampler(local)
    clk.start()
    l0 = _lib.launch_count()
    ms_dev = timed(lambda i: one_batch(d_q[args.warmup + i], d_x[args.warmup + i]), args.steps)
    launches = _lib.launch_count() - l0
    ms_e2e = timed(lambda i: e2e_batch(args.warmup + i), args.steps)

    # ---- strong scaling (N > 1): the global batch fixed at 64 images -> 64 / N per GPU
    strong = None
    if world > 1 and not args.no_strong and STRONG_GLOBAL_BATCH % world == 0:
        bs = STRONG_GLOBAL_BATCH // world
        unc_s = torch.zeros(bs, K_NN, D, device=dev)
        sq = [torch.from_numpy(rng.integers(0, n_db, size=bs)).to(dev) for _ in range(args.steps + 1)]
        sx = [torch.randn(bs, 4, 32, 32, device=dev) for _ in range(args.steps + 1)]
        one_batch(sq[0], sx[0], unc=unc_s)
        ms_s = timed(lambda i: one_batch(sq[1 + i], sx[1 + i], unc=unc_s), args.steps)
        strong = {"global_batch": STRONG_GLOBAL_BATCH, "batch_per_gpu": bs, "value": STRONG_GLOBAL_BATCH * args.steps / (ms_s * 1e-3), "unit": "images/s",
                  "ms_per_step": ms_s / args.steps, "scaling": "strong"}
    clk.stop_flag = True
    clk.join(timeout=2)

    # ---- the exchange alone (N > 1): sharded retrieval of one batch, no U-Net
    exchange = None
    if world > 1:
        for _ in range(3):
            retrieve(d_q[0], uncond)
        ms_x = timed(lambda i: retrieve(d_q[i % nbatches], uncond), 10) / 10
        exchange = {"ms_per_batch": ms_x, "what": "gather(queries from owners) + all_gather(queries) + local scan of N_db/N rows x (N*16 queries) + all_to_all(idx|score) + merge + "
                                                 "all_gather(idx) + local gather + reduce_scatter(rows)",
                    "scan_rows_per_gpu": hi - lo, "aggregate_scan_gbs": n_db * D * 2 / ms_x / 1e6, "queries_per_scan": BATCH * world}

    # ---- live roofline numbers (same process, right after the timed region)
    pk = peaks()
    _, ctx0 = retrieve(d_q[0], uncond)
    unet.set_context(ctx0)
    prof = [unet.profile_forward(d_x[0], torch.full((2 * BATCH,), int(t), device=dev)) for t in (991, 501, 11)]
    tc_flop = float(np.mean([p["tc_flop"] for p in prof])); tc_ms_events = float(np.mean([p["tc_ms"] for p in prof]))
    n_tc = prof[0]["n_tc"]
    # In-graph duration of the tcgen05 launches: graph-replayed forward with every kernel vs. the same graph without the GEMM launches
    # (rdm_unet_set_ablation), CUDA events on the launching stream.  Per-launch event brackets cannot resolve ~10 us kernels (an event
    # record costs microseconds on the device timeline), so they are reported separately as tc_ms_event_brackets.
    t_probe = torch.full((2 * BATCH,), 501, device=dev)

    def fwd_ms(net, x, mask, chains, reps=20):
        net.set_chains(chains)
        net.set_ablation(mask)
        for _ in range(3):
            net.forward(x, t_probe)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(reps):
            net.forward(x, t_probe)
        b.record(); torch.cuda.synchronize()
        net.set_ablation(0)
        net.set_chains(args.chains)
        return a.elapsed_time(b) / reps
    # the roofline pair is taken with ONE chain: with concurrent chains the GEMM-less graph overlaps differently and the difference would
    # no longer be the duration of the GEMM launches; forward_ms_graph is the forward as the step runs it
    fwd_run = fwd_ms(unet, d_x[0], 0, args.chains)
    # (each graph twice, alternating, 50 replays, the smaller reading of each: one 20-replay window differed by 2 % from its repeat in r2zb)
    pairs = [(fwd_ms(unet, d_x[0], 0, 1, 50), fwd_ms(unet, d_x[0], 48, 1, 50)) for _ in range(2)]
    fwd_full, fwd_nogemm = min(p[0] for p in pairs), min(p[1] for p in pairs)
    tc_ms = fwd_full - fwd_nogemm
    achieved = tc_flop / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    step_flop = BATCH * 2 * S_DDIM * FLOP_PER_FWD_SAMPLE
    # kNN search alone (the HBM-bound sink), local shard
    q0 = local_searcher.gather_device(torch.arange(lo, lo + BATCH, device=dev))
    for _ in range(3):
        local_searcher.search_raw_device(q0, K_NN)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        local_searcher.search_raw_device(q0, K_NN)
    e1.record(); torch.cuda.synchronize()
    knn_ms = e0.elapsed_time(e1) / 20
    knn_gbs = (hi - lo) * D * 2 / knn_ms / 1e6

    # ---- the reference's shipped shape (64x64x3), N = 1 only: device-resident value, e2e through the drop-in API, forward time
    r_shape = None
    if world == 1 and not args.no_r_shape:
        sd_r = make_weights(cfg=UNET_R)
        unet_r = B200UNet(dev, **UNET_R)
        unet_r.load_state_dict(sd_r); unet_r.set_mode(mode); unet_r.set_chains(args.chains)
        model_r = build_dropin(UNET_R, sd_r, searcher, args.mode, args.chains, dev)
        nr = max(1, args.steps // 2)
        xr_h = [pin(torch.randn(BATCH, 3, 64, 64, generator=torch.Generator().manual_seed(500 + i))) for i in range(nr + 1)]
        xr_d = [t.to(dev) for t in xr_h]
        out_r = pin(torch.empty(BATCH, 3, 64, 64))
        one_batch(d_q[0], xr_d[0], net=unet_r)
        e2e_batch(0, mdl=model_r, xs=xr_h, qs=qids_np, out=out_r)
        ms_r = timed(lambda i: one_batch(d_q[1 + i], xr_d[1 + i], net=unet_r), nr)
        ms_re = timed(lambda i: e2e_batch(1 + i, mdl=model_r, xs=xr_h, qs=qids_np, out=out_r), nr)
        f_r = fwd_ms(unet_r, xr_d[0], 0, args.chains, reps=10)
        r_shape = {"workload": "cfg2-R: the shipped shape, 64x64x3 latent (models/rdm/imagenet/config.yaml:14-59), DDIM-100, CFG 2.0, k=4, batch 16",
                   "value": BATCH * nr / (ms_r * 1e-3), "e2e": BATCH * nr / (ms_re * 1e-3), "unit": "images/s", "steps": nr, "ms_per_step": ms_r / nr,
                   "forward_ms_graph": f_r, "tflops_whole_step": BATCH * 2 * S_DDIM * FLOP_PER_FWD_SAMPLE_R / (ms_r / nr * 1e-3) / 1e12,
                   "frac_of_sustained_peak_whole_step": BATCH * 2 * S_DDIM * FLOP_PER_FWD_SAMPLE_R / (ms_r / nr * 1e-3) / 1e12 / pk["bf16_tflops_sustained"]}
        del unet_r, model_r

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    imgs = BATCH * world * args.steps
    value, e2e = imgs / (ms_dev * 1e-3), imgs / (ms_e2e * 1e-3)
    tr_bytes, tr_src = traffic()
    mma = {"bf16x3": 3, "fp16x2": 2, "fp16": 1, "bf16": 1, "fp32": 0}[args.mode]
    db_desc = (f"{N_DB:,}x512 fp16 DB resident per GPU" if world == 1 else
               f"{N_DB_SHARDED:,}x512 fp16 DB row-sharded {world} ways ({hi - lo:,} rows per GPU), sharded search + owner fetch inside the timed step")
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16x3": "bf16x3 (hi/lo-split bf16 operands on tcgen05, fp32 accumulate; fp32 norms/softmax)", "bf16": "bf16", "fp32": "f32",
                  "fp16x2": "fp16x2 (fp16 activations x hi/lo-split fp16 weights on tcgen05, fp32 accumulate; fp32 norms/softmax)",
                  "fp16": "fp16 (fp16 operands on tcgen05, fp32 accumulate; fp32 norms/softmax/residual stream)"}[args.mode],
        "data": "synthetic",
        "config": {"workload": f"cfg2: RDM ImageNet-arch U-Net (400.9M params) on 32x32x4 latent, DDIM-100, CFG 2.0, k=4 exact kNN over {db_desc}",
                   "batch_per_gpu": BATCH, "global_batch": BATCH * world, "ddim_steps": S_DDIM, "k_nn": K_NN,
                   "parallelism": f"dp{world} (images sharded by batch" + (", DB replicated)" if world == 1 else f", DB row-sharded {world} ways)"),
                   "unet_mode": args.mode, "chains": args.chains, "db_rows": n_db, "db_rows_per_gpu": hi - lo, "db_build_s": t_load,
                   "l2": "inputs larger than L2 (0.8-1.6 GB of weights + the database streamed per step vs 126 MB L2)", "cuda_graph": True},
        "clocks": clk.summary(),
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": int(qids_np[0].nbytes + h_xT[0].numel() * 4), "d2h_bytes_per_step": int(h_out.numel() * 4),
                "ms_per_step": ms_e2e / args.steps,
                "api": "rdm.models.diffusion.ddpm.MinimalRETRODiffusion (instantiate_from_config) .sample_from_rdata(N, qids=<host>, k_nn=4, unconditional_guidance_scale=2.0, "
                       "ddim_steps=100, ddim=True, unconditional_retro_guidance_label=0., x_T=<pinned host>) -> ema_scope -> DDIMSampler.sample; result copied to pinned host memory"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 implicit GEMM; all conv / linear layers of one U-Net forward)", "achieved": achieved,
                     "peak": pk["bf16_tflops_sustained"], "peak_src": pk["src"] + " bf16 sustained (kernel timed inside a long step)", "unit": "TFLOP/s",
                     "frac": achieved / pk["bf16_tflops_sustained"], "traffic": tr_bytes, "traffic_src": tr_src,
                     "algorithmic_flop_per_forward": tc_flop, "tc_launches_per_forward": n_tc, "tc_ms_per_forward": tc_ms,
                     "forward_ms_graph": fwd_run, "forward_ms_graph_one_chain": fwd_full, "forward_ms_graph_without_gemms": fwd_nogemm, "tc_ms_event_brackets": tc_ms_events,
                     "whole_step_tflops": step_flop / (ms_dev / args.steps * 1e-3) / 1e12,
                     "whole_step_frac": step_flop / (ms_dev / args.steps * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
                     "mma_per_product": mma,
                     "note": "achieved = algorithmic FLOPs of all tcgen05 launches of one U-Net forward over their in-graph duration (single-chain forward minus the "
                             "same graph without the GEMM launches, CUDA events; algorithmic = the reference's dense conv / linear count 2MNK -- the three Upsample convs "
                             "execute 4/9 of theirs through the parity fold, 94 GFLOP per forward less than counted); whole_step_* = algorithmic FLOPs of the timed step over its wall time, "
                             "everything included (glue kernels, kNN, chains overlap)"},
        "knn": {"ms": knn_ms, "qps": BATCH / knn_ms * 1e3, "gbs": knn_gbs, "frac_hbm": knn_gbs / pk["hbm_gbs"], "peak_gbs": pk["hbm_gbs"], "rows": hi - lo,
                "what": "rdm_knn_search_raw: 16 raw queries, k = 4, normalisation + scans + exact re-rank, this GPU's rows; peak = the measured COPY bandwidth "
                        "(read + write), so a read-only scan of a large shard can exceed 1.0"},
    }
    if r_shape is not None:
        line["r_shape"] = r_shape
    if world > 1:
        line["sharded_knn"] = {"check": shard_check, "exchange": exchange}
        if strong is not None:
            line["strong"] = strong
    if not args.no_cpu_baseline and world == 1:
        torch.set_num_threads(os.cpu_count())
        v, info = cpu_reference_images_per_sec(4)
        line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "4 DDIM steps (CFG, B2=2) of 1 image on the torch-CPU fp32 oracle + C kNN oracle over 100K rows, extrapolated to DDIM-100 / 1.28M rows",
                                **info}
    print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
