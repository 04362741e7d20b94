"""Graph-replayed full-arch U-Net forward time (B2 = 32, 32x32x4) against the number of concurrent batch chains.  python tools/chains_sweep.py [modes] [chains]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "retrieval-augmented-diffusion-models_b200")]
import torch
import bench
from rdm_b200.unet import B200UNet
modes = [int(m) for m in sys.argv[1].split(",")] if len(sys.argv) > 1 else [3, 4]
chains = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8]
shape = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [16, 4, 32]      # B, C, H
dev = torch.device("cuda:0")
cfg = dict(bench.UNET, in_channels=shape[1], out_channels=shape[1], image_size=shape[2])
net = B200UNet(dev, **cfg)
sd = bench.make_weights() if shape[1] == 4 else None
if sd is None:
    from oracle import unet as ounet
    sd = ounet.randomize_(ounet.UNetModel(**cfg), 3).state_dict()
net.load_state_dict(sd)
g = torch.Generator(device=dev).manual_seed(0)
B = shape[0]
x = torch.randn(B, shape[1], shape[2], shape[2], generator=g, device=dev); t = torch.full((2 * B,), 501, device=dev)
ctx = torch.randn(2 * B, 4, 512, generator=g, device=dev)
out = {}
for mode in modes:
    net.set_mode(mode)
    base = None
    for ch in chains:
        net.set_chains(ch)
        net.set_context(ctx)
        y = net.forward(x, t)
        for _ in range(5):
            net.forward(x, t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            net.forward(x, t)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        if base is None:
            base = y.clone()
        err = float((y - base).norm() / base.norm())
        out[f"mode{mode}_chains{ch}"] = dict(ms=round(ms, 4), rel_vs_first=err)
        print(f"mode {mode} chains {ch}: {ms:.3f} ms   (rel diff vs first {err:.1e})", flush=True)
print(json.dumps(out))
