"""Golden vectors from the REFERENCE's own hot-path code, run on CPU in the build container.

    python tests/golden/make_golden_ref.py        (writes tests/golden/ref_*.npz; needs /root/reference)

The reference's modules are imported UNMODIFIED from /root/reference; the un-vendored packages they import are replaced by the
stand-ins of tests/golden/ref_stubs.py (see its header for what that covers).  What the fixtures pin:

  ref_unet_tiny.npz   rdm/modules/diffusionmodules/openaimodel.py  UNetModel.__init__/forward (:66-317, :335-371),
                      TimestepEmbedSequential (:17-33); rdm/modules/attention.py SpatialTransformer (:122-196),
                      BasicTransformerBlock (:77-96), CrossAttention (:20-74)                          -> oracle/unet.py
  ref_ddim_tiny.npz   rdm/models/diffusion/ddim.py DDIMSampler.make_schedule/sample/ddim_sampling/p_sample_ddim (:27-268) with
                      classifier-free guidance, eta = 0 and eta = 0.5, over the same U-Net                -> oracle/ddim.py
  ref_rarm_small.npz  rdm/modules/attention.py RetrievalPatchTransformer (:199-272; discrete tokens, positional encodings, causal
                      self-attention, cross-attention to the retrieved vectors) and the sampling arithmetic of
                      rdm/models/autoregression/transformer.py LatentImageRETRO.sample (:224-270)         -> oracle/rarm.py

The only reference method replaced is DDIMSampler.register_buffer (ddim.py:21-25), which force-moves every buffer to "cuda"
(SURVEY F6); the override keeps them on the CPU and changes no arithmetic.  tests/test_oracle_ref_golden.py checks the oracle
against these files; the GPU tests check the CUDA path against the same files.  Nothing on the GPU box reads /root/reference.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_stubs  # noqa: E402
import ref_weights  # noqa: E402

UNET_CFG = dict(image_size=16, in_channels=4, out_channels=4, model_channels=64, attention_resolutions=[2, 4], num_res_blocks=1,
                channel_mult=[1, 2, 3], num_head_channels=32, use_spatial_transformer=True, transformer_depth=1, context_dim=512)
# = oracle.unet.TINY_UNET (every layer kind, non-aligned concat GroupNorm groups) + the reference's use_spatial_transformer switch
RARM_CFG = dict(in_channels=50, n_heads=2, d_head=64, depth=2, context_dim=128, positional_encodings=True, sequence_length=12,
                out_channels=48, cross_attend=True, causal=True, continuous=False)


def save(name, out):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def unet_and_ddim():
    from rdm.modules.diffusionmodules.openaimodel import UNetModel          # the reference's
    from rdm.models.diffusion.ddim import DDIMSampler                        # the reference's
    assert UNetModel.__module__.startswith("rdm.") and "/root/reference" in sys.modules[UNetModel.__module__].__file__
    net = ref_weights.fill_(UNetModel(**UNET_CFG), 11).eval()
    g = torch.Generator().manual_seed(12)
    x = torch.randn(4, 4, 16, 16, generator=g)
    t = torch.tensor([1, 501, 991, 250])
    ctx = torch.randn(4, 3, 512, generator=g)
    with torch.no_grad():
        y = net(x, t, context=[ctx])                                         # the list form RETRODiffusionWrapper passes (ddpm.py:128-131)
    out = {"cfg_json": np.array(repr(UNET_CFG)), "x": x.numpy(), "t": t.numpy(), "context": ctx.numpy(), "out": y.numpy(), "weight_seed": np.int64(11),
           "sd_keys": np.array(list(net.state_dict().keys())), "n_params": np.int64(sum(p.numel() for p in net.parameters()))}
    save("ref_unet_tiny.npz", out)

    class CpuDDIMSampler(DDIMSampler):
        def register_buffer(self, name, attr):                               # ddim.py:21-25 without the forced .to("cuda")
            setattr(self, name, attr)

    class Model:                                                             # what DDIMSampler touches of LatentDiffusion
        num_timesteps = 1000
        device = torch.device("cpu")
        parameterization = "eps"

        def __init__(self):
            betas = ref_stubs.make_beta_schedule("linear", 1000, linear_start=0.0015, linear_end=0.0195)      # config.yaml:7-11
            ac = np.cumprod(1.0 - betas, axis=0)
            f32 = lambda a: torch.tensor(a, dtype=torch.float32)             # ldm DDPM.register_schedule: to_torch = float32
            self.betas, self.alphas_cumprod, self.alphas_cumprod_prev = f32(betas), f32(ac), f32(np.append(1.0, ac[:-1]))

        def apply_model(self, x_noisy, t, cond):                             # ddpm.py:445-458 -> :128-131
            return net(x_noisy, t, context=[cond])

    xT = torch.randn(2, 4, 16, 16, generator=g)
    c, uc = torch.randn(2, 3, 512, generator=g), torch.zeros(2, 3, 512)
    out = {"x_T": xT.numpy(), "cond": c.numpy(), "uncond": uc.numpy(), "scale": np.float32(2.0), "S": np.int64(5)}
    for tag, eta, seed in (("eta0", 0.0, 0), ("eta05", 0.5, 7)):
        torch.manual_seed(seed)
        s = CpuDDIMSampler(Model())
        samples, inter = s.sample(5, 2, (4, 16, 16), conditioning=c, eta=eta, x_T=xT.clone(), verbose=False, log_every_t=1,
                                  unconditional_guidance_scale=2.0, unconditional_conditioning=uc)
        out[f"{tag}:samples"] = samples.numpy()
        out[f"{tag}:x_inter"] = torch.stack(inter["x_inter"][1:]).numpy()
        out[f"{tag}:pred_x0"] = torch.stack(inter["pred_x0"][1:]).numpy()
        out[f"{tag}:seed"] = np.int64(seed)
        if eta == 0.0:
            out["ddim_timesteps"] = np.asarray(s.ddim_timesteps)
            out["ddim_alphas"] = np.asarray(s.ddim_alphas)
            out["ddim_alphas_prev"] = np.asarray(s.ddim_alphas_prev)
        else:
            out["eta05:ddim_sigmas"] = np.asarray(s.ddim_sigmas)
    # plain conditional sampling (scale 1: no batch doubling, ddim.py:239-240)
    torch.manual_seed(0)
    s = CpuDDIMSampler(Model())
    samples, _ = s.sample(4, 2, (4, 16, 16), conditioning=c, eta=0.0, x_T=xT.clone(), verbose=False)
    out["noguid:samples"] = samples.numpy()
    save("ref_ddim_tiny.npz", out)


def rarm():
    from rdm.modules.attention import RetrievalPatchTransformer                # the reference's
    net = ref_weights.fill_(RetrievalPatchTransformer(**RARM_CFG), 21).eval()
    g = torch.Generator().manual_seed(22)
    tok = torch.randint(0, 50, (3, 12), generator=g)
    tok[:, 0] = 49                                                             # sos id = last vocabulary entry (config.yaml: sos_token 16385 of 16386)
    ctx = torch.randn(3, 4, 128, generator=g)
    with torch.no_grad():
        logits = net(tok, context=ctx)
        logits_short = net(tok[:, :5], context=ctx)                            # a prefix: what step 4 of the sampling loop evaluates
        logits_uncond = net(tok, context=torch.zeros_like(ctx))
    out = {"cfg_json": np.array(repr(RARM_CFG)), "tokens": tok.numpy(), "context": ctx.numpy(), "logits": logits.numpy(),
           "logits_prefix5": logits_short.numpy(), "logits_uncond": logits_uncond.numpy(), "weight_seed": np.int64(21),
           "sd_keys": np.array(list(net.state_dict().keys())), "n_params": np.int64(sum(p.numel() for p in net.parameters()))}
    save("ref_rarm_small.npz", out)


if __name__ == "__main__":
    ref_stubs.install()
    unet_and_ddim()
    rarm()
