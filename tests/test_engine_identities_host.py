"""CPU: the two algebraic identities the tensor-core executor relies on since round 2, restated in NumPy / torch exactly the way the CUDA
code indexes them, against torch's own convolution.  (The CUDA path itself is held to the oracle by tests/test_unet_gpu.py and
tests/test_variants_gpu.py, which run both forms of each layer; these tests pin the index arithmetic where it can be read.)

* Upsample fold (csrc/kernels.cu fold_up_weights_kernel, csrc/gemm_tc.cu TcKernelParams::ups): a 3x3 / pad 1 conv over the 2x NEAREST
  upsampled image equals, for each output parity (py, px), a 2x2-tap conv over the SOURCE image with summed weights; tap (ty, tx) reads
  source pixel (y + ty + py - 1, x + tx + px - 1) (zero outside), and GEMM row (b, y, x) is output pixel (b, 2y + py, 2x + px).
  Reference op: ldm Upsample.forward (interpolate nearest x2, then conv) as used by openaimodel.py's output blocks.
* K extension (csrc/gemm_tc.cuh TcA::hi2): conv2(h) + skip_connection(x) of a ResBlock is ONE contraction over the concatenated K axis
  [9 * Cout taps of h | Cin channels of x] with the summed bias.

The last section runs the CUDA kernels of csrc/kernels.cu that prepare those operands (fold_up_weights_kernel, add_vec_kernel) and the
register-tiled first convolution (conv_first_kernel) AS WRITTEN under the host emulation of tests/emu, through test-only entry points
(tests/emu/tc_stubs.cpp)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def fold_up_weights(w):
    """w [N, C, 3, 3] (torch layout) -> [4, N, 2, 2, C]; the kernel's rule: rows that land on source row y + ty + py - 1 are
    py = 0: ty 0 <- {0}, ty 1 <- {1, 2};  py = 1: ty 0 <- {0, 1}, ty 1 <- {2} (same for columns)."""
    N, C = w.shape[:2]
    out = np.zeros((4, N, 2, 2, C), w.dtype)
    rows = {(0, 0): (0, 0), (0, 1): (1, 2), (1, 0): (0, 1), (1, 1): (2, 2)}
    for q in range(4):
        py, px = q >> 1, q & 1
        for ty in range(2):
            for tx in range(2):
                ky0, ky1 = rows[(py, ty)]
                kx0, kx1 = rows[(px, tx)]
                acc = np.zeros((N, C), w.dtype)
                for ky in range(ky0, ky1 + 1):                      # same summation order as the kernel (ky, then kx; fp32 there)
                    for kx in range(kx0, kx1 + 1):
                        acc = acc + w[:, :, ky, kx]
                out[q, :, ty, tx, :] = acc
    return out


@pytest.mark.parametrize("B,C,N,H,W", [(2, 8, 6, 4, 4), (1, 5, 7, 3, 6), (3, 4, 4, 1, 1), (1, 3, 2, 8, 2)])
def test_upsample_then_conv_equals_four_parity_convs_with_folded_weights(B, C, N, H, W):
    g = torch.Generator().manual_seed(B * 100 + C)
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(N, C, 3, 3, generator=g, dtype=torch.float64)
    b = torch.randn(N, generator=g, dtype=torch.float64)
    want = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, b, padding=1)           # the reference's Upsample + conv
    wf = fold_up_weights(w.numpy())
    xp = np.zeros((B, C, H + 2, W + 2)); xp[:, :, 1:-1, 1:-1] = x.numpy()                        # zero fill outside = TMA out-of-bounds boxes
    got = np.zeros((B, N, 2 * H, 2 * W))
    for q in range(4):
        py, px = q >> 1, q & 1
        for y in range(H):
            for xx in range(W):
                acc = b.numpy().copy()
                for ty in range(2):
                    for tx in range(2):
                        sy, sx = y + ty + py - 1, xx + tx + px - 1                              # source pixel of this tap
                        acc = acc + np.einsum("bc,nc->bn", xp[:, :, sy + 1, sx + 1], wf[q, :, ty, tx, :])
                got[:, :, 2 * y + py, 2 * xx + px] = acc                                        # epilogue row remap
    np.testing.assert_allclose(got, want.numpy(), rtol=1e-12, atol=1e-12)


def test_folded_weights_do_4_ninths_of_the_multiplies_and_cover_every_tap_once():
    w = np.arange(2 * 3 * 9, dtype=np.float32).reshape(2, 3, 3, 3) + 1
    wf = fold_up_weights(w)
    assert wf.shape == (4, 2, 2, 2, 3)
    # every original tap is used exactly once per output parity: the folded weights of a parity sum to the sum of the 3x3 kernel
    for q in range(4):
        np.testing.assert_allclose(wf[q].sum(axis=(1, 2)), w.sum(axis=(2, 3)))
    assert wf[0].size * 4 * 9 == w.size * 16                                                     # 16 / 9 of the weights, 4 / 9 of the MACs per output


@pytest.mark.parametrize("Cin,Cout,H", [(6, 4, 5), (3, 8, 2)])
def test_resblock_tail_is_one_contraction_over_taps_plus_skip_channels(Cin, Cout, H):
    g = torch.Generator().manual_seed(Cin)
    h = torch.randn(2, Cout, H, H, generator=g, dtype=torch.float64)                             # SiLU(GN(...)) operand of out_layers' conv
    x = torch.randn(2, Cin, H, H, generator=g, dtype=torch.float64)                              # the block's input (skip_connection operand)
    w2 = torch.randn(Cout, Cout, 3, 3, generator=g, dtype=torch.float64); b2 = torch.randn(Cout, generator=g, dtype=torch.float64)
    ws = torch.randn(Cout, Cin, 1, 1, generator=g, dtype=torch.float64); bs = torch.randn(Cout, generator=g, dtype=torch.float64)
    want = F.conv2d(x, ws, bs) + F.conv2d(h, w2, b2, padding=1)                                  # openaimodel.py ResBlock: skip_connection(x) + h
    # one GEMM: A = [im2col(h) (tap-major K) | x], W = [w2 as [N][tap][c] | ws], bias = b2 + bs
    cols = F.unfold(h, 3, padding=1).reshape(2, Cout, 9, H * H).permute(0, 3, 2, 1).reshape(2 * H * H, 9 * Cout)
    A = torch.cat([cols, x.permute(0, 2, 3, 1).reshape(2 * H * H, Cin)], dim=1)
    Wm = torch.cat([w2.permute(0, 2, 3, 1).reshape(Cout, 9 * Cout), ws.reshape(Cout, Cin)], dim=1)
    got = (A @ Wm.T + (b2 + bs)).reshape(2, H, H, Cout).permute(0, 3, 1, 2)
    torch.testing.assert_close(got, want, rtol=1e-12, atol=1e-12)


# ---- the CUDA kernels themselves (csrc/kernels.cu as written) under the host emulation of tests/emu --------------------------------------
@pytest.fixture(scope="module")
def emu():
    import ctypes, os, shutil, sys
    if shutil.which("g++") is None:
        pytest.skip("needs g++ to build the emulated library")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
    import build_emu
    return ctypes.CDLL(build_emu.build())


def _p(a):
    import ctypes
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


@pytest.mark.parametrize("N,C", [(4, 8), (7, 3), (32, 64)])
def test_fold_up_weights_kernel_is_the_restated_rule_bit_for_bit(emu, N, C):
    """fold_up_weights_kernel: input [N][9][C] (the packed conv layout of unet.cu), output [4][N][2*2][C]; fp32 sums in (ky, kx) order."""
    rng = np.random.default_rng(N * 100 + C)
    w = rng.standard_normal((N, C, 3, 3)).astype(np.float32)                                      # torch layout
    packed = np.ascontiguousarray(w.transpose(0, 2, 3, 1).reshape(N, 9, C))                      # reg_conv_w: [cout][tap][cin]
    out = np.full((4, N, 4, C), np.nan, np.float32)
    assert emu.emu_fold_up_weights(_p(packed), N, C, _p(out)) == 0
    want = fold_up_weights(w).reshape(4, N, 4, C)
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))


def test_add_vec_kernel(emu):
    a = np.arange(1000, dtype=np.float32) * 0.37; b = np.arange(1000, dtype=np.float32)[::-1].copy() * 1e-3
    out = np.empty(1000, np.float32)
    assert emu.emu_add_vec(_p(a), _p(b), _p(out), 1000) == 0
    assert np.array_equal(out, a + b)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 8, 8, 4, 32), (1, 5, 7, 3, 16), (3, 4, 4, 4, 192), (1, 9, 3, 1, 8), (2, 16, 16, 3, 64)])
def test_first_convolution_kernel_matches_torch(emu, B, H, W, Cin, Cout):
    """conv_first_kernel (register-tiled direct 3x3 conv over 1-4 input channels, openaimodel.py:141-145 input_blocks.0.0): ragged pixel
    counts (the last CTA is partial), K = 27 padded to 28 for three input channels, bias, the zero padding ring."""
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + Cin)
    x = torch.randn(B, Cin, H, W, generator=g); w = torch.randn(Cout, Cin, 3, 3, generator=g) * 0.3; b = torch.randn(Cout, generator=g)
    want = F.conv2d(x.double(), w.double(), b.double(), padding=1).permute(0, 2, 3, 1).reshape(B * H * W, Cout).numpy()
    xn = np.ascontiguousarray(x.permute(0, 2, 3, 1).reshape(B * H * W, Cin).numpy())
    packed = np.ascontiguousarray(w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).numpy())
    out = np.full((B * H * W, Cout), np.nan, np.float32)
    assert emu.emu_conv_first(_p(xn), B, H, W, Cin, _p(packed), _p(b.numpy()), Cout, _p(out)) == 0
    np.testing.assert_allclose(out, want, rtol=2e-5, atol=2e-5)
