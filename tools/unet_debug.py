"""Per-layer comparison of the CUDA U-Net against the oracle (debug aid).  python tools/unet_debug.py [tiny|full]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "retrieval-augmented-diffusion-models_b200")]
import torch
from oracle import unet as ounet
from rdm_b200 import _lib
from rdm_b200.unet import B200UNet

which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cfg = ounet.TINY_UNET if which == "tiny" else ounet.BASELINE_UNET
H = 16 if which == "tiny" else 32
dev = torch.device("cuda:0")
ref = ounet.randomize_(ounet.UNetModel(**cfg), 1).eval()
net = B200UNet(dev, **cfg); net.load_state_dict(ref.state_dict()); net.set_mode(mode)
g = torch.Generator().manual_seed(0)
x = torch.randn(2, cfg["in_channels"], H, H, generator=g); t = torch.tensor([991, 17]); c = torch.randn(2, 4, 512, generator=g) * 3
rows = []
blocks = list(ref.input_blocks) + [ref.middle_block] + list(ref.output_blocks)
for bi, blk in enumerate(blocks):
    for li, layer in enumerate(blk):
        layer.register_forward_hook(lambda m, i, o, bi=bi, li=li: rows.append((bi, li, type(m).__name__, float(o.mean()), float(o.abs().mean()))))
with torch.no_grad():
    want = ref(x, t, c)
_lib.lib().rdm_unet_set_debug(net._h, 1)
net.set_context(c.to(dev)); got = net.forward(x.to(dev), t.to(dev))
log = _lib.lib().rdm_unet_debug_log(net._h).decode().strip().split("\n")
got_rows = {(int(a), int(b)): (float(m), float(am)) for tag, a, b, m, am in (l.split() for l in log) if tag == "layer"}
for l in log:
    if not l.startswith("layer"):
        print(l)
for bi, li, name, m, am in rows:
    gm, gam = got_rows.get((bi, li), (float("nan"),) * 2)
    flag = "" if abs(gam - am) <= 1e-4 * am + 1e-7 else "   <<<<<< MISMATCH"
    print(f"block {bi:2d} layer {li} {name:20s} ref mean {m:+.6f} |x| {am:.6f}   cuda mean {gm:+.6f} |x| {gam:.6f}{flag}")
print("final rel-L2:", float((got.cpu().double() - want.double()).norm() / want.double().norm()))
