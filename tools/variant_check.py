"""Forward parity of one engine-configuration variant (selected by environment variables that the library reads once at load time:
RDM_TC_CLUSTER, RDM_SKIP, RDM_PDL, RDM_TC_NOSPLIT) against the torch-CPU fp32 oracle.  Prints one JSON line {mode: rel-L2, ...}.
python tools/variant_check.py [modes, default 0,1,3,4]      (tests/test_variants_gpu.py runs this in sub-processes)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "retrieval-augmented-diffusion-models_b200")]
import torch
from oracle import unet as ounet
from rdm_b200.unet import B200UNet
modes = [int(m) for m in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 1, 3, 4]
dev = torch.device("cuda:0")
ref = ounet.randomize_(ounet.UNetModel(**ounet.BASELINE_UNET), 3).eval()
net = B200UNet(dev, **ounet.BASELINE_UNET); net.load_state_dict(ref.state_dict())
g = torch.Generator().manual_seed(23)
x = torch.randn(2, 4, 32, 32, generator=g); t = torch.randint(0, 1000, (2,), generator=g); c = torch.randn(2, 4, 512, generator=g) * 3
with torch.no_grad():
    want = ref(x, t, c).double()
out = {}
for m in modes:
    net.set_mode(m); net.set_context(c.to(dev))
    got = net.forward(x.to(dev), t.to(dev)).double().cpu()
    got2 = net.forward(x.to(dev), t.to(dev)).double().cpu()          # second call = CUDA-graph replay
    out[str(m)] = max(float((got - want).norm() / want.norm()), float((got2 - want).norm() / want.norm()))
print(json.dumps(out))
