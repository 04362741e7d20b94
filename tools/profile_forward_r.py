"""Graph-replayed forward time of the reference's shipped shape (64x64x3 latent, B2 = 32).  python tools/profile_forward_r.py [mode] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "retrieval-augmented-diffusion-models_b200")]
import torch
import bench
from rdm_b200.unet import B200UNet
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
net = B200UNet(dev, **bench.UNET_R); net.load_state_dict(bench.make_weights(cfg=bench.UNET_R)); net.set_mode(mode)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(16, 3, 64, 64, generator=g, device=dev); t = torch.full((32,), 501, device=dev)
net.set_context(torch.randn(32, 4, 512, generator=g, device=dev))
for _ in range(2):
    net.forward(x, t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(nrep):
    net.forward(x, t)
e1.record(); torch.cuda.synchronize()
print(f"R-shape graph forward: {e0.elapsed_time(e1)/nrep:.3f} ms")
