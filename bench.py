#!/usr/bin/env python
"""Headline benchmark: images/sec of the retrieval-augmented DDIM sampler (BASELINE.json cfg2).

One "step" = one pass of the hot path over one batch per GPU:
    kNN (16 queries, k=4, exact cosine over the 1,281,167 x 512 fp16 DB in HBM) -> gather raw neighbour rows ->
    cross-attention K/V projection of [cond | uncond=0] -> DDIM-100 with classifier-free guidance 2.0 over the
    ImageNet-RDM U-Net (mc 192, mult 1-2-3-5, 16 SpatialTransformers) on the 32x32x4 latent -> 16 latents.
Synthetic data / random-init weights of the named architecture (no network for checkpoints or databases).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode fp16x2|bf16x3|fp16|bf16|fp32]
Under torchrun (N > 1) every rank runs the same per-GPU batch (weak scaling, images sharded by batch index,
no data-path collective; one all_gather of the finished latents per step), timing = max over ranks.
"""
import argparse
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
N_DB, D, K_NN, BATCH, S_DDIM, CFG_SCALE = 1_281_167, 512, 4, 16, 100, 2.0
FLOP_PER_FWD_SAMPLE = 50.56e9          # SURVEY.md section 6 (32x32x4 latent, k=4)
# dram__bytes_read.sum + dram__bytes_write.sum over the 197 gemm_tc launches of one forward (fp16x2, B2 = 32), from the ncu pass committed as
# profiles/gemm_tc_dram_r1f.csv (1.79 GB read = two fp16 weight planes + activations, 1.28 GB written)
TRAFFIC_BYTES, TRAFFIC_SRC = 3.07e9, "profiles/gemm_tc_dram_r1f.csv (ncu, per forward = 197 launches, fp16x2)"


def peaks():
    p = dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, src="fallback")
    try:
        j = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        p.update(hbm_gbs=j["hbm_gbs"], bf16_tflops=j["bf16_tflops"], bf16_tflops_sustained=j.get("bf16_tflops_sustained", j["bf16_tflops"]), src="measured")
    except Exception:
        pass
    return p


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


def make_weights(seed=3):
    """Random-init weights of the named architecture, generated directly as a state dict (no nn.Module needed)."""
    from rdm_b200.unet import unet_param_shapes
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shp in unet_param_shapes(**UNET).items():
        if len(shp) >= 2:
            fan_in = int(np.prod(shp[1:]))
            sd[name] = torch.randn(shp, generator=g) / fan_in ** 0.5
        elif name.endswith("weight"):
            sd[name] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            sd[name] = 0.05 * torch.randn(shp, generator=g)
    return sd


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


def run_reference(args):
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
    sample = "per step: 2 DDIM steps (CFG, B2=2) of 1 image on the torch-CPU oracle + C kNN oracle over 100K rows, extrapolated to DDIM-100 / 1.28M rows"
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "cfg2 RDM ImageNet-arch 32x32x4, DDIM-100, CFG 2.0, k=4 over 1,281,167x512 fp16 DB (CPU oracle port; reference not installable)"},
                      "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--mode", default="fp16x2", choices=["bf16x3", "bf16", "fp32", "fp16x2", "fp16"],
                    help="U-Net contraction mode; measured DDIM-100 final-latent rel-L2 vs the fp32 oracle (tools/ddim_error.py): "
                         "fp32 1.4e-6, bf16x3 8.9e-6, fp16x2 2.4e-4 (default: inside the 1e-3 tolerance with 4x margin), fp16 6.5e-4, bf16 6.0e-3 (outside)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from rdm_b200 import _lib, sampler
    from rdm_b200.knn import B200Searcher
    from rdm_b200.unet import B200UNet

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mode = {"fp32": 0, "bf16x3": 1, "bf16": 2, "fp16x2": 3, "fp16": 4}[args.mode]

    # ---- resident state: DB (replicated per GPU: 1.31 GB fp16), weights, schedule tables
    g = torch.Generator(device=dev).manual_seed(1)
    db = torch.randn((N_DB, D), generator=g, device=dev, dtype=torch.float32).to(torch.float16)
    searcher = B200Searcher(db, device=dev)
    unet = B200UNet(dev, **UNET)
    unet.load_state_dict(make_weights())
    unet.set_mode(mode)
    tables = sampler.make_ddim_tables(sampler.alphas_cumprod_linear(), S_DDIM, 0.0, device=dev)
    rng = np.random.default_rng(2 + rank)
    pin = lambda t: t.pin_memory()
    h_qids = [pin(torch.from_numpy(rng.integers(0, N_DB, size=BATCH))) for _ in range(args.steps + args.warmup)]
    h_xT = [pin(torch.randn(BATCH, 4, 32, 32, generator=torch.Generator().manual_seed(100 * rank + i))) for i in range(args.steps + args.warmup)]
    h_out = pin(torch.empty(BATCH, 4, 32, 32))
    uncond = torch.zeros(BATCH, K_NN, D, device=dev)          # unconditional_retro_guidance_label = 0 (rdm_sample.py:251, ddpm.py:673-680)

    def one_batch(qids_dev, xT_dev):
        q = searcher.gather_device(qids_dev)                                  # query = DB rows (ddpm.py:897)
        qh = q / q.norm(dim=1, keepdim=True)                                  # ddpm.py:907
        nns, _ = searcher.search_device(qh, K_NN)                             # ddpm.py:906-908
        cond = searcher.gather_device(nns)                                    # ddpm.py:921 (raw rows, fp32)
        unet.set_context(torch.cat([cond, uncond]))                           # cat([c, uc]) ddim.py:232
        return unet.ddim_sample(xT_dev, tables["timesteps"], tables["coef"], cfg_scale=CFG_SCALE)

    def e2e_batch(i):
        qd = h_qids[i].to(dev, non_blocking=True)
        xd = h_xT[i].to(dev, non_blocking=True)
        out = one_batch(qd, xd)
        if world > 1:
            allo = [torch.empty_like(out) for _ in range(world)]
            dist.all_gather(allo, out)
        h_out.copy_(out, non_blocking=True)
        return out

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

    d_q = [t.to(dev) for t in h_qids]
    d_x = [t.to(dev) for t in h_xT]
    for i in range(args.warmup):
        one_batch(d_q[i], d_x[i])
        e2e_batch(i)
    clk = ClockSampler(local)
    clk.start()
    l0 = _lib.launch_count()
    ms_dev = timed(lambda i: one_batch(d_q[args.warmup + i], d_x[args.warmup + i]), args.steps)
    launches = _lib.launch_count() - l0
    ms_e2e = timed(lambda i: e2e_batch(args.warmup + i), args.steps)
    clk.stop_flag = True
    clk.join(timeout=2)

    # ---- live roofline numbers (CUDA events around every GEMM launch, same process, right after the timed region)
    pk = peaks()
    unet.set_context(torch.cat([searcher.gather_device(searcher.search_device(torch.nn.functional.normalize(searcher.gather_device(d_q[0]), dim=1), K_NN)[0]), uncond]))
    prof = [unet.profile_forward(d_x[0], torch.full((2 * BATCH,), int(t), device=dev)) for t in (991, 501, 11)]
    tc_flop = float(np.mean([p["tc_flop"] for p in prof])); tc_ms_events = float(np.mean([p["tc_ms"] for p in prof]))
    n_tc = prof[0]["n_tc"]
    # In-graph duration of the tcgen05 launches: graph-replayed forward with every kernel vs. the same graph without the GEMM launches
    # (rdm_unet_set_ablation), CUDA events on the launching stream.  Per-launch event brackets cannot resolve ~10 us kernels (an event
    # record costs microseconds on the device timeline), so they are reported separately as tc_ms_event_brackets.
    t_probe = torch.full((2 * BATCH,), 501, device=dev)

    def fwd_ms(mask, reps=20):
        unet.set_ablation(mask)
        for _ in range(3):
            unet.forward(d_x[0], t_probe)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(reps):
            unet.forward(d_x[0], t_probe)
        b.record(); torch.cuda.synchronize()
        unet.set_ablation(0)
        return a.elapsed_time(b) / reps
    fwd_full, fwd_nogemm = fwd_ms(0), fwd_ms(48)
    tc_ms = fwd_full - fwd_nogemm
    achieved = tc_flop / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    # kNN scan alone (the HBM-bound sink)
    qh = torch.nn.functional.normalize(searcher.gather_device(d_q[0]), dim=1)
    for _ in range(3):
        searcher.search_device(qh, K_NN)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        searcher.search_device(qh, K_NN)
    e1.record(); torch.cuda.synchronize()
    knn_ms = e0.elapsed_time(e1) / 20
    knn_gbs = N_DB * D * 2 / knn_ms / 1e6

    if rank != 0:
        return
    imgs = BATCH * world * args.steps
    value, e2e = imgs / (ms_dev * 1e-3), imgs / (ms_e2e * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16x3": "bf16x3 (hi/lo-split bf16 operands on tcgen05, fp32 accumulate; fp32 norms/softmax)", "bf16": "bf16", "fp32": "f32",
                  "fp16x2": "fp16x2 (fp16 activations x hi/lo-split fp16 weights on tcgen05, fp32 accumulate; fp32 norms/softmax)",
                  "fp16": "fp16 (fp16 operands on tcgen05, fp32 accumulate; fp32 norms/softmax)"}[args.mode],
        "data": "synthetic",
        "config": {"workload": "cfg2: RDM ImageNet-arch U-Net (400.9M params) on 32x32x4 latent, DDIM-100, CFG 2.0, k=4 exact kNN over 1,281,167x512 fp16 DB",
                   "batch_per_gpu": BATCH, "global_batch": BATCH * world, "ddim_steps": S_DDIM, "k_nn": K_NN, "parallelism": f"dp{world} (images sharded by batch, DB replicated)",
                   "unet_mode": args.mode, "l2": "inputs larger than L2 (0.8-3.2 GB of weights + 1.3 GB DB streamed per step vs 126 MB L2)", "cuda_graph": True},
        "clocks": clk.summary(),
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": int(h_qids[0].numel() * 8 + h_xT[0].numel() * 4), "d2h_bytes_per_step": int(h_out.numel() * 4),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 implicit GEMM; all conv / linear layers of one U-Net forward)", "achieved": achieved,
                     "peak": pk["bf16_tflops_sustained"], "peak_src": pk["src"] + " bf16 sustained (kernel timed inside a long step)", "unit": "TFLOP/s",
                     "frac": achieved / pk["bf16_tflops_sustained"], "traffic": TRAFFIC_BYTES,
                     "algorithmic_flop_per_forward": tc_flop, "tc_launches_per_forward": n_tc, "tc_ms_per_forward": tc_ms,
                     "forward_ms_graph": fwd_full, "forward_ms_graph_without_gemms": fwd_nogemm, "tc_ms_event_brackets": tc_ms_events,
                     "traffic_src": TRAFFIC_SRC,
                     "mma_per_product": {"bf16x3": 3, "fp16x2": 2, "fp16": 1, "bf16": 1, "fp32": 0}[args.mode],
                     "note": "algorithmic FLOPs of all tcgen05 launches of one U-Net forward over their in-graph duration (forward minus GEMM-less forward, "
                             "CUDA events); split-operand modes issue 2-3 MMAs per product, so frac <= 1/2 (fp16x2) or 1/3 (bf16x3)"},
        "knn": {"ms": knn_ms, "qps": BATCH / knn_ms * 1e3, "gbs": knn_gbs, "frac_hbm": knn_gbs / pk["hbm_gbs"], "peak_gbs": pk["hbm_gbs"]},
    }
    if not args.no_cpu_baseline and world == 1:
        torch.set_num_threads(os.cpu_count())
        v, info = cpu_reference_images_per_sec(4)
        line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "4 DDIM steps (CFG, B2=2) of 1 image on the torch-CPU fp32 oracle + C kNN oracle over 100K rows, extrapolated to DDIM-100 / 1.28M rows",
                                **info}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
