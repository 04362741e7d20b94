"""Per-GEMM timing of one full-arch U-Net forward (B2=32, 32x32x4).  python tools/profile_forward.py [mode] [graph_forwards]"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "retrieval-augmented-diffusion-models_b200")]
import torch
import bench
from rdm_b200 import _lib
from rdm_b200.unet import B200UNet
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda:0")
net = B200UNet(dev, **bench.UNET); net.load_state_dict(bench.make_weights()); net.set_mode(mode)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(16, 4, 32, 32, generator=g, device=dev); t = torch.full((32,), 501, device=dev)
net.set_context(torch.randn(32, 4, 512, generator=g, device=dev))
net.forward(x, t); torch.cuda.synchronize()
if nrep:                      # plain graph-replayed forwards (for ncu / timing)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(nrep):
        net.forward(x, t)
    e1.record(); torch.cuda.synchronize()
    print(f"graph forward: {e0.elapsed_time(e1)/nrep:.3f} ms")
    sys.exit(0)
p = net.profile_forward(x, t); p = net.profile_forward(x, t)
print(p)
lines = _lib.lib().rdm_unet_profile_text(net._h).decode().strip().split("\n")
agg = collections.OrderedDict()
for l in lines:
    key = l.split(" ms=")[0]; ms = float(l.split(" ms=")[1].split()[0])
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(v[1] for v in agg.values())
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    f = dict(kv.split("=") for kv in k.split()[1:] if "=" in kv and "x" not in kv.split("=")[1])
    fl = 2.0 * int(f["M"]) * int(f["N"]) * int(f["K"]) * c
    print(f"{ms:8.3f} ms {100*ms/tot:5.1f}%  x{c:3d}  {fl/(ms*1e-3)/1e12:7.1f} TF/s  {k}")
print("total gemm ms", tot)
