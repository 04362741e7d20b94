// tcgen05 GEMM engine interface (gemm_tc.cu).  Operands are bf16 hi/lo planes (x ~= hi + lo).
#pragma once
#include "kernels.cuh"

struct TcA {                       // activation operand: NHWC plane [B*H*W, C] (row stride ld elements)
    const __nv_bfloat16* hi = nullptr; const __nv_bfloat16* lo = nullptr; int ld = 0;
    int B = 1, H = 1, W = 1, C = 0;
    int ksize = 1;                 // 1: plain [M, C] matrix (B*H*W rows); 3: 3x3 / pad 1 / stride 1 conv over the (H, W) grid
    // K extension (one-plane modes): out += A2[M, C2] * W2[N, C2]^T in the SAME accumulator -- the 1x1 skip_connection of a ResBlock rides
    // on its second 3x3 conv (no second GEMM, no fp32 residual round trip)
    const __nv_bfloat16* hi2 = nullptr; int ld2 = 0, C2 = 0;
    // ups = 1 (with ksize = 3): the conv runs over the 2x NEAREST-UPSAMPLED image of this [B, H, W, C] plane (Upsample + conv of the decoder
    // path, openaimodel.py Upsample.forward) without materialising it: output pixel (2y + py, 2x + px) only sees the 2x2 source pixels
    // (y + ty + py - 1, x + tx + px - 1), so each of the 4 output parities is a 2x2-tap conv with pre-summed weights -- 4/9 of the MMA work
    // and no 4x larger operand plane.  Weights: [4 N, 4 C] (parity-major rows, (ty, tx, c) columns: unet.cu fold_up_weights); the output
    // has 4 * B * H * W rows.
    int ups = 0;
};
struct TcW {                       // weights [N, K] K-major, K ordered (tap, cin)
    const __nv_bfloat16* hi = nullptr; const __nv_bfloat16* lo = nullptr; int ld = 0;
    int N = 0, K = 0;
    const __nv_bfloat16* hi2 = nullptr; int ld2 = 0;      // weights of the K extension [N, C2]
    int dynamic = 0;               // 1: this operand is an ACTIVATION written by preceding kernels (attention K / V^T): it must not be prefetched before griddepcontrol.wait
};
bool gemm_tc_supported(const TcA& a);
// nsplit 3: A_hi*W_hi + A_lo*W_hi + A_hi*W_lo (fp32-grade);  nsplit 2: A*W_hi + A*W_lo (single-plane A, split W);  nsplit 1: A*W.
// f16: the 16-bit planes hold IEEE fp16 (11-bit significand) instead of bf16 (8-bit); kind::f16 MMA either way.
// Output either fp32 (e.out) or 16-bit planes (out_hi[/out_lo], same format flag); e.res / e.bias / e.rowvec / e.act as for gemm_simt.
int gemm_tc(const TcA& a, const TcW& w, const GemmEpi& e, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int out_bf_ld, int nsplit, int f16, cudaStream_t st);
// Concurrent batch chains (unet.cu) issue their GEMMs from one host thread, chain after chain: the split-K partial-sum workspace used by
// the launches that follow is slot `slot` (0..7), so that chains running at the same time on different streams never share partials.
void gemm_tc_set_workspace_slot(int slot);
