"""Final-latent error of each engine mode vs the torch-CPU fp32 oracle, full cfg2 architecture.  python tools/ddim_error.py [S] [B]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "retrieval-augmented-diffusion-models_b200")]
import torch
import bench
from oracle import ddim as oddim, unet as ounet
from rdm_b200 import sampler
from rdm_b200.unet import B200UNet
S = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
modes = [int(m) for m in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2]
dev = torch.device("cuda:0")
sd = bench.make_weights()
ref = ounet.UNetModel(**bench.UNET).eval(); ref.load_state_dict(sd)
g = torch.Generator().manual_seed(31)
x_T = torch.randn(B, 4, 32, 32, generator=g)
cond, unc = torch.randn(B, 4, 512, generator=g) * 3, torch.zeros(B, 4, 512)
t0 = time.time(); want = oddim.ddim_sample(ref, x_T, cond, unc, S=S, scale=2.0); print(f"oracle DDIM-{S} B={B}: {time.time()-t0:.1f}s", flush=True)
net = B200UNet(dev, **bench.UNET); net.load_state_dict(sd)
tb = sampler.make_ddim_tables(sampler.alphas_cumprod_linear(), S, 0.0, device=dev)
for mode in modes:
    net.set_mode(mode)
    net.set_context(torch.cat([cond, unc]).to(dev))
    got = net.ddim_sample(x_T.to(dev), tb["timesteps"], tb["coef"], cfg_scale=2.0)
    err = float((got.cpu().double() - want.double()).norm() / want.double().norm())
    print(f"mode {mode}: DDIM-{S} final-latent rel-L2 = {err:.3e}", flush=True)
