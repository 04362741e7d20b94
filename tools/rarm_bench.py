"""RARM decode throughput: the 256-token sampling loop of scripts/rarm_sample.py (batch 4, k = 4 neighbours, top-k 256) on the
ImageNet-size decoder with random weights.  Prints one JSON line: tokens/s, images/s, us per step and the weight-streaming roofline of the
GEMV kernels (algorithmic bytes per step = sum over dense layers of N*K*sizeof(weight) + the mean live KV cache).

    python tools/rarm_bench.py [--batch 4] [--guidance 1.0] [--mode fp16|fp32] [--reps 5]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "retrieval-augmented-diffusion-models_b200"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--guidance", type=float, default=1.0)
    ap.add_argument("--mode", default="fp16", choices=["fp16", "fp32"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-graph", action="store_true")
    a = ap.parse_args()
    import ref_weights
    from rdm_b200 import _lib
    from rdm_b200.rarm import MODE_FP16, MODE_FP32, RARM_IMAGENET, B200Rarm
    dev = torch.device("cuda:0")
    net = B200Rarm(dev, **RARM_IMAGENET)
    net.load_state_dict(ref_weights.state_dict_for(net.shapes.items(), 31))
    net.set_mode(MODE_FP16 if a.mode == "fp16" else MODE_FP32)
    net.set_graph(not a.no_graph)
    B, steps = a.batch, 256
    B2 = 2 * B if a.guidance > 1.0 else B
    g = torch.Generator().manual_seed(0)
    ctx = torch.randn(B, 4, 512, generator=g)
    r = torch.cat([ctx, torch.zeros_like(ctx)]) if a.guidance > 1.0 else ctx
    sos = torch.full((B, 1), 16385)
    u = torch.rand(steps, B, generator=g).to(dev)
    times = []
    for rep in range(a.reps + 2):
        net.set_context(r)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        e0.record()
        toks = net.sample(sos, steps, temperature=1.0, top_k=256, guidance_scale=a.guidance, uniforms=u)
        e1.record()
        torch.cuda.synchronize()
        if rep >= 2:
            times.append(e0.elapsed_time(e1))
        launches = _lib.launch_count() - l0
    ms = sorted(times)[len(times) // 2]
    C, L, X, V = 768, 18, 512, 16384
    wsize = 2 if a.mode == "fp16" else 4
    w_elems = L * (3 * C * C + C * C + C * C + C * C + 8 * C * C + 4 * C * C) + V * C           # the dense layers of one step (context K/V is per sample)
    kv_bytes = 2 * L * B2 * (steps / 2) * C * 4                                                  # mean live cache per step
    step_bytes = w_elems * wsize + kv_bytes
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6458.7))
    gbs = step_bytes / (ms / steps * 1e-3) / 1e9
    print(json.dumps({"metric": "RARM tokens/sec (256-token loop, top-k 256)", "value": B * steps / (ms * 1e-3), "unit": "tokens/s", "images_per_s": B / (ms * 1e-3),
                      "ms_per_batch": ms, "us_per_step": ms / steps * 1e3, "batch": B, "rows": B2, "mode": a.mode, "graph": not a.no_graph,
                      "gpu_launches": int(launches), "tokens_in_range": bool(int(toks[:, 1:].max()) < V),
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                   "bytes_per_step": step_bytes, "weight_bytes_per_step": w_elems * wsize},
                      "reference_published": "7.46 s per batch of 4 (scripts/demo_rarm.ipynb, no KV cache, unknown GPU)"}))


if __name__ == "__main__":
    main()
