#!/usr/bin/env python
"""Stage-by-stage measurements of BASELINE.json configs 3 and 4 -- the configurations next to the headline cfg2 that
bench.py reports.  Synthetic data, random-init weights of the named architectures.  One JSON line (rank 0); results under profiles/.

  cfg 3  text2img:  synthetic prompts -> CLIP ViT-B/32 text tower -> exact kNN (k=4) over a 20,927,907 x 512 fp16 database row-sharded over
         the ranks (all_gather + merge) -> context = [query | 3 neighbours] -> DDIM-250 with guidance 2.0, batch 64 sharded by image.
  cfg 4  per-step re-retrieval: batch 32 sharded by image, k=8, DDIM-100; every step: U-Net step -> VQ-f8 decode of the x0 prediction ->
         bicubic 224 + CLIP image tower -> q/|q| -> sharded kNN -> gather -> cross-attention K/V re-projection (ddim.py:355-380).

  python tools/cfg_bench.py --cfg 3 [--rows 20927907] [--batch 64] [--steps 250]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/cfg_bench.py --cfg 4
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "retrieval-augmented-diffusion-models_b200")]
import bench  # noqa: E402  (UNET architecture, make_weights, peaks)

VQ_F8 = dict(embed_dim=4, n_embed=16384, ddconfig=dict(double_z=False, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                                                        ch_mult=[1, 2, 2, 4], num_res_blocks=2, attn_resolutions=[32], dropout=0.0))
"""VQ-f8 first stage for the 32x32x4 latent of BASELINE.json (ldm vq-f8 layout: 4 levels, attention at resolution 32)."""


class Timer:
    def __init__(self):
        self.ms = {}

    def run(self, name, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        e1.synchronize()
        self.ms[name] = self.ms.get(name, 0.0) + e0.elapsed_time(e1)
        return out


def synthetic_db(rows_total, lo, hi, dev, chunk=1 << 20):
    """Rows [lo, hi) of a deterministic fp16 database: chunk c is N(0,1) from seed 6000 + c, whatever the sharding."""
    db = torch.empty((hi - lo, 512), dtype=torch.float16, device=dev)
    g = torch.Generator(device=dev)
    for c in range(lo // chunk, (hi + chunk - 1) // chunk):
        a, b = max(lo, c * chunk), min(hi, (c + 1) * chunk)
        g.manual_seed(6000 + c)
        block = torch.randn((min(chunk, rows_total - c * chunk), 512), generator=g, device=dev, dtype=torch.float32)
        db[a - lo:b - lo] = block[a - c * chunk:b - c * chunk].to(torch.float16)
    return db


def synthetic_prompts(B, seed=4):
    g = torch.Generator().manual_seed(seed)
    tok = torch.zeros(B, 77, dtype=torch.long)
    for b in range(B):
        L = int(torch.randint(8, 21, (1,), generator=g))
        tok[b, 0] = 49406
        tok[b, 1:1 + L] = torch.randint(1, 49406, (L,), generator=g)
        tok[b, 1 + L] = 49407
    return tok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, required=True, choices=[3, 4])
    ap.add_argument("--rows", type=int, default=20_927_907)
    ap.add_argument("--batch", type=int, default=None, help="GLOBAL batch (default 64 for cfg 3, 32 for cfg 4)")
    ap.add_argument("--steps", type=int, default=None, help="DDIM steps (default 250 / 100)")
    ap.add_argument("--mode", default=bench.DEFAULT_MODE)
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")
    from oracle import clip as oclip          # random-init weights of the named architecture only (the oracle computes nothing here)
    from rdm_b200 import sampler
    from rdm_b200.clip import VIT_B32, B200Clip
    from rdm_b200.knn import B200Searcher, ShardedSearcher, shard_range
    from rdm_b200.unet import B200UNet
    B_glob = a.batch or (64 if a.cfg == 3 else 32)
    S = a.steps or (250 if a.cfg == 3 else 100)
    k = 4 if a.cfg == 3 else 8
    assert B_glob % world == 0
    B = B_glob // world
    modes = {"fp32": 0, "bf16x3": 1, "bf16": 2, "fp16x2": 3, "fp16": 4}
    t_setup = time.time()
    lo, hi = shard_range(a.rows, rank, world)
    db = synthetic_db(a.rows, lo, hi, dev)
    local_s = B200Searcher(db, device=dev, idx_base=lo)
    searcher = ShardedSearcher(local_s, validate=False) if world > 1 else local_s
    net = B200UNet(dev, **bench.UNET)
    net.load_state_dict(bench.make_weights())
    net.set_mode(modes[a.mode])
    clip = B200Clip(dev, **VIT_B32)
    clip.load_state_dict(oclip.random_state_dict(**VIT_B32, seed=5))
    tb = sampler.make_ddim_tables(sampler.alphas_cumprod_linear(), S, 0.0, device=dev)
    x_T = torch.randn(B, 4, 32, 32, generator=torch.Generator().manual_seed(100 + rank)).to(dev)
    setup_s = time.time() - t_setup
    T = Timer()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.time()
    if a.cfg == 3:
        tok = synthetic_prompts(B_glob)[rank * B:(rank + 1) * B].to(dev)          # every rank encodes ITS prompts; the sharded searcher exchanges the query rows
        q = T.run("clip_text", lambda: clip.encode_text(tok).float())
        nns, _ = T.run("knn", lambda: searcher.search_raw_device(q, k))          # q / ||q|| inside the library (ddpm.py:907), all_gather / all_to_all when sharded
        rows = T.run("gather", lambda: searcher.gather_device(nns))              # neighbour rows from their owner ranks
        cond = torch.cat([q[:, None], rows[:, :k - 1]], dim=1)                   # the query itself is neighbour 0 (ddpm.py:775)
        T.run("kv_projection", lambda: net.set_context(torch.cat([cond, torch.zeros_like(cond)])))
        x = T.run("ddim", lambda: net.ddim_sample(x_T, tb["timesteps"], tb["coef"], cfg_scale=2.0))
    else:
        from rdm_b200.vqdecoder import B200VQDecoder
        from oracle import vqdecoder as ovq
        dec = B200VQDecoder(dev, VQ_F8["embed_dim"], VQ_F8["n_embed"], VQ_F8["ddconfig"])
        dec.load_state_dict(ovq.randomize_(ovq.VQModelInterface(**VQ_F8), 8).state_dict())
        cond = torch.randn(B, k, 512, generator=torch.Generator().manual_seed(7)).to(dev)   # first conditioning: noise of shape r_shape (ddim.py:300-305)
        x = x_T
        for i in range(S):
            T.run("kv_projection", lambda: net.set_context(torch.cat([cond, torch.zeros_like(cond)])))
            x, p0 = T.run("unet_step", lambda: net.ddim_sample(x, tb["timesteps"], tb["coef"], cfg_scale=2.0, first_step=i, num_steps=1, want_pred_x0=True))
            img = T.run("vq_decode", lambda: dec.decode(p0, force_not_quantize=True))
            q = T.run("clip_image", lambda: clip.encode_image(clip.preprocess(img)).float())
            nns, _ = T.run("knn", lambda: searcher.search_raw_device(q, k))      # per-rank queries; exchange inside the sharded searcher
            cond = T.run("gather", lambda: searcher.gather_device(nns))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.time() - t0
    t = torch.tensor([wall], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        pk = bench.peaks()
        knn_calls = 1 if a.cfg == 3 else S
        nq = B_glob                                                             # queries every shard scans per search (the union of all ranks' queries)
        knn_ms = T.ms["knn"] / knn_calls
        print(json.dumps({"config": f"BASELINE cfg{a.cfg}", "n_gpus": world, "global_batch": B_glob, "ddim_steps": S, "k_nn": k, "db_rows": a.rows,
                          "unet_mode": a.mode, "images_per_s": B_glob / float(t), "wall_s": float(t), "setup_s": setup_s,
                          "stage_ms_rank0": {n: round(v, 3) for n, v in T.ms.items()},
                          "knn": {"queries": nq, "ms_per_search": knn_ms, "rows_per_gpu": hi - lo, "gbs_per_gpu": (hi - lo) * 512 * 2 / (knn_ms * 1e-3) / 1e9,
                                  "frac_hbm": (hi - lo) * 512 * 2 / (knn_ms * 1e-3) / 1e9 / pk["hbm_gbs"]},
                          "finite": bool(torch.isfinite(x).all())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
