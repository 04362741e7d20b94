"""CPU: the first-stage oracle (oracle/vqdecoder.py) -- codebook lookup against a brute-force numpy nearest neighbour, decoder shapes of the
shipped VQ-f4 configuration (models/rdm/imagenet/config.yaml:60-80), checkpoint key layout the C++ decoder registers."""
import numpy as np
import torch

from oracle import vqdecoder as ovq


def test_lookup_is_the_nearest_codebook_row():
    m = ovq.randomize_(ovq.VQModelInterface(**ovq.TINY_VQ), 0).eval()
    z = torch.randn(2, 3, 8, 8, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        zq, _, (_, _, idx) = m.quantize(z)
    E = m.quantize.embedding.weight.detach().numpy().astype(np.float64)
    zf = z.permute(0, 2, 3, 1).reshape(-1, 3).numpy().astype(np.float64)
    want = ((zf[:, None, :] - E[None]) ** 2).sum(-1).argmin(1)
    assert np.array_equal(idx.numpy(), want)
    assert torch.equal(zq, m.quantize.embedding.weight.detach()[idx].view(2, 8, 8, 3).permute(0, 3, 1, 2))


def test_f4_configuration_shapes_and_keys():
    m = ovq.VQModelInterface(**ovq.RDM_VQ_F4)
    keys = set(m.state_dict().keys())
    for k in ("decoder.conv_in.weight", "decoder.mid.attn_1.q.weight", "decoder.mid.attn_1.proj_out.bias", "decoder.up.2.upsample.conv.weight",
              "decoder.up.1.block.0.nin_shortcut.weight", "decoder.up.0.block.2.conv2.bias", "decoder.norm_out.weight", "decoder.conv_out.weight",
              "quantize.embedding.weight", "post_quant_conv.weight"):
        assert k in keys, k
    assert "decoder.up.0.upsample.conv.weight" not in keys and "decoder.up.2.block.0.nin_shortcut.weight" not in keys
    assert m.quantize.embedding.weight.shape == (8192, 3) and m.decoder.conv_in.weight.shape == (512, 3, 3, 3)
    tiny = ovq.randomize_(ovq.VQModelInterface(**ovq.TINY_VQ), 2).eval()
    with torch.no_grad():
        assert tiny.decode(torch.randn(1, 3, 16, 16)).shape == (1, 3, 32, 32)          # f = 2^(len(ch_mult) - 1)
