"""Generates tests/golden/clip_small.npz by running the REFERENCE's own vendored CLIP implementation
(/root/reference/rdm/modules/custom_clip/model.py -- torch-only, importable in the build container) on a small seeded
configuration.  Run once in the build container:  python tests/golden/make_golden.py
The fixture pins oracle/clip.py (tests/test_oracle_clip.py); nothing on the GPU box reads /root/reference.
"""
import importlib.util
import os

import numpy as np
import torch

REF = "/root/reference/rdm/modules/custom_clip/model.py"
CFG = dict(embed_dim=64, image_resolution=32, vision_layers=2, vision_width=64, vision_patch_size=16, context_length=16, vocab_size=64,
           transformer_width=64, transformer_heads=1, transformer_layers=2)


def main():
    spec = importlib.util.spec_from_file_location("ref_clip_model", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(0)
    model = mod.CLIP(**CFG).float().eval()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in model.named_parameters():          # informative, well-scaled weights (the reference init leaves some tensors empty/tiny)
            if p.ndim >= 2:
                p.copy_(torch.randn(p.shape, generator=g) / (p.shape[-1] if n.endswith("proj") or n == "text_projection" else p[0].numel()) ** 0.5)
            elif "ln" in n and n.endswith("weight"):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
            elif p.ndim == 1:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    img = torch.randn(3, 3, 32, 32, generator=g)
    tok = torch.randint(1, 60, (4, 16), generator=g)
    for b, L in enumerate((5, 9, 16, 12)):             # EOT = the largest id sits at different positions; padding zeros after it
        tok[b, L - 1] = 63
        tok[b, L:] = 0
    with torch.no_grad():
        fi, ft = model.encode_image(img), model.encode_text(tok)
    out = {"cfg_keys": np.array(list(CFG.keys())), "cfg_vals": np.array(list(CFG.values())), "image": img.numpy(), "tokens": tok.numpy(),
           "image_features": fi.numpy(), "text_features": ft.numpy()}
    for k, v in model.state_dict().items():
        out["sd:" + k] = v.numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "clip_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", sum(v.numel() for v in model.state_dict().values()), "values")


if __name__ == "__main__":
    main()
