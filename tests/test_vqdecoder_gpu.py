"""GPU parity of the first-stage decode path: librdm_b200's VQ decoder (rdm_vqdec_decode through the C ABI)
vs the torch-CPU fp32 oracle (oracle/vqdecoder.py) on identical weights and latents."""
import pytest
import torch

from oracle import vqdecoder as ovq

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


def _pair(cfg, seed, cuda):
    from rdm_b200.vqdecoder import B200VQDecoder
    ref = ovq.randomize_(ovq.VQModelInterface(**cfg), seed).eval()
    dec = B200VQDecoder(cuda, cfg["embed_dim"], cfg["n_embed"], cfg["ddconfig"])
    assert sorted(dec.names) == sorted(ref.state_dict().keys()), "parameter inventory of the C++ decoder and the oracle differ"
    dec.load_state_dict(ref.state_dict())
    return ref, dec


@pytest.mark.parametrize("mode,tol", [(3, 3e-3), (4, 6e-3)])
def test_tiny_decoder_without_quantisation(cuda, mode, tol):
    ref, dec = _pair(ovq.TINY_VQ, 1, cuda)
    dec.set_mode(mode)
    z = torch.randn(3, 3, 32, 32, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        want = ref.decode(z, force_not_quantize=True)
    got = dec.decode(z.to(cuda), force_not_quantize=True)
    again = dec.decode(z.to(cuda), force_not_quantize=True)                 # CUDA-graph replay
    assert got.shape == want.shape == (3, 3, 64, 64)
    assert rel_l2(got, want) < tol and rel_l2(again, want) < tol


def test_codebook_lookup_is_exact(cuda):
    """Latents built from codebook rows plus a small perturbation: the lookup must return exactly those rows (index-exact), so the decode
    equals the un-quantised decode of the clean rows."""
    ref, dec = _pair(ovq.TINY_VQ, 2, cuda)
    g = torch.Generator().manual_seed(9)
    idx = torch.randint(0, ovq.TINY_VQ["n_embed"], (2, 32, 32), generator=g)
    clean = ref.quantize.embedding.weight.detach()[idx].permute(0, 3, 1, 2).contiguous()
    z = clean + 1e-3 * torch.randn(clean.shape, generator=g)
    with torch.no_grad():
        assert torch.equal(ref.quantize(z)[2][2].view(2, 32, 32), idx)       # the perturbation does not change the oracle's argmin
    got_q = dec.decode(z.to(cuda))
    got_clean = dec.decode(clean.to(cuda), force_not_quantize=True)
    assert torch.equal(got_q, got_clean)            # bit-identical: every reduction on this path has a fixed order or runs in fp64


def test_f4_shaped_decoder_256px(cuda):
    """VQ-f4 layout of models/rdm/imagenet/config.yaml:60-80 at half width (ch 64): 64x64x3 latent -> 256x256x3, i.e. 4096-token mid
    attention on the tensor-core engine and 3x3 convolutions over 128- and 256-pixel rows."""
    cfg = dict(ovq.RDM_VQ_F4, n_embed=1024, ddconfig=dict(ovq.RDM_VQ_F4["ddconfig"], ch=64))
    ref, dec = _pair(cfg, 3, cuda)
    z = torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(11))
    with torch.no_grad():
        want = ref.decode(z)
    got = dec.decode(z.to(cuda))
    assert got.shape == want.shape == (1, 3, 256, 256)
    err = rel_l2(got, want)
    assert err < 5e-3, f"rel-L2 {err:.2e}"


def test_shim_routes_cuda_tensors_to_the_kernel(cuda):
    import rdm  # noqa: F401  (puts shims/ on sys.path when ldm is not installed)
    from ldm.models.autoencoder import VQModelInterface
    m = VQModelInterface(**ovq.TINY_VQ).eval()
    ref = ovq.randomize_(ovq.VQModelInterface(**ovq.TINY_VQ), 4).eval()
    m.load_state_dict(ref.state_dict())
    z = torch.randn(1, 3, 32, 32, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        want = ref.decode(z, force_not_quantize=True)
    got = m.to(cuda).decode(z.to(cuda), force_not_quantize=True)
    assert got.is_cuda and rel_l2(got, want) < 3e-3
