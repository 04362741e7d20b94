"""Fixtures that pin the BENCHMARKED configuration and the reference's SHIPPED shape (VERDICT r1, items 1b / missing 1).

Writes, from the torch-CPU fp32 oracle (oracle/unet.py, oracle/ddim.py -- themselves pinned to the reference's own modules by
ref_unet_tiny.npz / ref_ddim_tiny.npz):

  ddim100_cfg2.npz   cfg2-B: bench.py's U-Net (bench.UNET, bench.make_weights()) on the 32x32x4 latent, DDIM-100, CFG 2.0, zeros uncond,
                     batch 16 (= bench.BATCH), seeds 0..NSEEDS-1: final latents [NSEEDS, 16, 4, 32, 32]
  rshape_imagenet.npz  R: models/rdm/imagenet/config.yaml:14-59 (64x64x3 latent, in_channels 3), weights randomize_(seed 3):
                     one forward at B2 = 2 and DDIM-20 (CFG 2.0) of one image
  rshape_imagenet_ddim100.npz  the same model, DDIM-100 (CFG 2.0) of two images

Inputs are regenerated from seeds by `inputs_cfg2` / `inputs_rshape` below (the GPU tests import them), only the oracle's outputs are stored.
    python tests/golden/make_golden_ddim100.py [cfg2|rshape|all] [nseeds]
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "retrieval-augmented-diffusion-models_b200")]

import numpy as np
import torch

NSEEDS, BATCH, S_DDIM, SCALE = 5, 16, 100, 2.0


def inputs_cfg2(seed, batch=BATCH):
    g = torch.Generator().manual_seed(1000 + seed)
    x_T = torch.randn(batch, 4, 32, 32, generator=g)
    cond = torch.randn(batch, 4, 512, generator=g) * 3
    return x_T, cond, torch.zeros_like(cond)


def inputs_rshape():
    g = torch.Generator().manual_seed(77)
    x = torch.randn(2, 3, 64, 64, generator=g)
    t = torch.tensor([991, 301])
    c = torch.cat([torch.randn(1, 4, 512, generator=g) * 3, torch.zeros(1, 4, 512)])
    x_T = torch.randn(1, 3, 64, 64, generator=g)
    cond = torch.randn(1, 4, 512, generator=g) * 3
    return x, t, c, x_T, cond, torch.zeros_like(cond)


def inputs_rshape100():
    g = torch.Generator().manual_seed(78)
    x_T = torch.randn(2, 3, 64, 64, generator=g)
    cond = torch.randn(2, 4, 512, generator=g) * 3
    return x_T, cond, torch.zeros_like(cond)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    nseeds = int(sys.argv[2]) if len(sys.argv) > 2 else NSEEDS
    import bench
    from oracle import ddim as oddim, unet as ounet
    torch.set_num_threads(os.cpu_count())
    if what in ("rshape", "all"):
        ref = ounet.randomize_(ounet.UNetModel(**ounet.IMAGENET_UNET), 3).eval()
        x, t, c, x_T, cond, unc = inputs_rshape()
        t0 = time.time()
        with torch.no_grad():
            fwd = ref(x, t, c)
        lat = oddim.ddim_sample(ref, x_T, cond, unc, S=20, scale=SCALE)
        np.savez(os.path.join(HERE, "rshape_imagenet.npz"), forward=fwd.numpy(), ddim20=lat.numpy())
        print(f"rshape: {time.time() - t0:.1f}s", flush=True)
    if what in ("rshape100", "all"):
        # the shipped shape at the headline step count: DDIM-100, CFG 2.0, two images (inputs_rshape100)
        ref = ounet.randomize_(ounet.UNetModel(**ounet.IMAGENET_UNET), 3).eval()
        x_T, cond, unc = inputs_rshape100()
        t0 = time.time()
        lat = oddim.ddim_sample(ref, x_T, cond, unc, S=S_DDIM, scale=SCALE)
        np.savez(os.path.join(HERE, "rshape_imagenet_ddim100.npz"), ddim100=lat.numpy())
        print(f"rshape100: {time.time() - t0:.1f}s", flush=True)
    if what in ("cfg2", "all"):
        ref = ounet.UNetModel(**bench.UNET).eval()
        ref.load_state_dict(bench.make_weights())
        out = []
        for s in range(nseeds):
            t0 = time.time()
            x_T, cond, unc = inputs_cfg2(s)
            out.append(oddim.ddim_sample(ref, x_T, cond, unc, S=S_DDIM, scale=SCALE).numpy())
            print(f"cfg2 seed {s}: {time.time() - t0:.1f}s", flush=True)
            np.savez(os.path.join(HERE, "ddim100_cfg2.npz"), latents=np.stack(out))


if __name__ == "__main__":
    main()
