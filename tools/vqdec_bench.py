"""Throughput of the first-stage decode at the shipped VQ-f4 configuration (models/rdm/imagenet/config.yaml:60-80):
B latents 64x64x3 -> B images 256x256x3, random-init weights.  python tools/vqdec_bench.py [B] [mode]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "retrieval-augmented-diffusion-models_b200")]
import torch
from oracle import vqdecoder as ovq          # configuration + random-init weights only (the oracle forward is not run here)
from rdm_b200.vqdecoder import B200VQDecoder
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
cfg = ovq.RDM_VQ_F4
ref = ovq.randomize_(ovq.VQModelInterface(**cfg), 0)
dec = B200VQDecoder(dev, cfg["embed_dim"], cfg["n_embed"], cfg["ddconfig"]); dec.load_state_dict(ref.state_dict()); dec.set_mode(mode)
z = torch.randn(B, 3, 64, 64, generator=torch.Generator().manual_seed(0)).to(dev)
for _ in range(3):
    dec.decode(z)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(5):
    out = dec.decode(z)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(json.dumps({"batch": B, "mode": mode, "ms_per_batch": ms, "images_per_s": B / ms * 1e3, "out_shape": list(out.shape),
                  "finite": bool(torch.isfinite(out).all())}))
