// tcgen05 GEMM engine interface (gemm_tc.cu).  Operands are bf16 hi/lo planes (x ~= hi + lo).
#pragma once
#include "kernels.cuh"

struct TcA {                       // activation operand: NHWC plane [B*H*W, C] (row stride ld elements)
    const __nv_bfloat16* hi = nullptr; const __nv_bfloat16* lo = nullptr; int ld = 0;
    int B = 1, H = 1, W = 1, C = 0;
    int ksize = 1;                 // 1: plain [M, C] matrix (B*H*W rows); 3: 3x3 / pad 1 / stride 1 conv over the (H, W) grid
};
struct TcW {                       // weights [N, K] K-major, K ordered (tap, cin)
    const __nv_bfloat16* hi = nullptr; const __nv_bfloat16* lo = nullptr; int ld = 0;
    int N = 0, K = 0;
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
