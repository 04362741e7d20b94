"""Where does a multi-chain forward differ from the single-chain one?  per-sample rel diff, modes 0/1/3, chains 1 -> 2 -> 1."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "retrieval-augmented-diffusion-models_b200")]
import torch
import bench
from rdm_b200.unet import B200UNet
dev = torch.device("cuda:0")
net = B200UNet(dev, **bench.UNET); net.load_state_dict(bench.make_weights())
g = torch.Generator(device=dev).manual_seed(0)
B = 8
x = torch.randn(B, 4, 32, 32, generator=g, device=dev); t = torch.full((2 * B,), 501, device=dev)
ctx = torch.randn(2 * B, 4, 512, generator=g, device=dev)
for mode in (0, 1, 3):
    net.set_mode(mode)
    outs = []
    for ch in (1, 2, 1, 4):
        net.set_chains(ch); net.set_context(ctx)
        y = net.forward(x, t).clone(); y2 = net.forward(x, t).clone()
        outs.append(y)
        print(f"mode {mode} chains {ch}: eager-vs-replay {float((y - y2).norm() / y.norm()):.2e}", flush=True)
    ref = outs[0]
    for name, o in zip(("2", "1again", "4"), outs[1:]):
        per = [(float((o[i] - ref[i]).norm() / ref[i].norm())) for i in range(2 * B)]
        print(f"mode {mode} chains {name} vs 1: total {float((o - ref).norm() / ref.norm()):.2e} per-sample " + " ".join(f"{p:.1e}" for p in per), flush=True)
