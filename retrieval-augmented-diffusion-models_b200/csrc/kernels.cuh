// Launch wrappers of the U-Net glue kernels (definitions in kernels.cu) and the GEMM engines.
// Activation layout everywhere: NHWC fp32, viewed as a row-major matrix [M = B*H*W, C] with a row
// stride `ld` (in floats) so channel-concatenation is just two producers writing into one buffer.
#pragma once
#include "common.cuh"

struct View {            // [M, C] fp32 matrix view
    float* p = nullptr;
    int ld = 0;          // row stride in floats (multiple of 4)
    int C = 0;
    // optional GroupNorm statistics slab that travels with the buffer: {sum, sum of squares} (fp64) per (image, column), entry (b, c) at
    // st[(b * st_ld + c) * 2]; accumulated by the epilogue of the tcgen05 GEMM that writes the view (zeroed once per forward)
    double* st = nullptr; int st_ld = 0;
    View() {}
    View(float* p_, int ld_, int C_) : p(p_), ld(ld_), C(C_) {}
    View cols(int c0, int n) const { View v(p + c0, ld, n); if (st) { v.st = st + 2 * (size_t)c0; v.st_ld = st_ld; } return v; }
};

// Destination of an element-wise producer: fp32 view and/or bf16 hi(/lo) planes (x ~= hi + lo) for the tcgen05 engine.
struct Out4 {
    float* f = nullptr; int ldf = 0;
    __nv_bfloat16* hi = nullptr; __nv_bfloat16* lo = nullptr; int ldb = 0;      // 16-bit planes: bf16, or IEEE fp16 bits when f16 != 0
    int f16 = 0;
    __host__ __device__ Out4() {}
    Out4(View v) : f(v.p), ldf(v.ld) {}
    Out4(__nv_bfloat16* h, __nv_bfloat16* l, int ld, int f16_ = 0) : hi(h), lo(l), ldb(ld), f16(f16_) {}
    __host__ __device__ bool any() const { return f || hi; }
};

// 16-bit split of an fp32 value: hi = rn16(x), lo = rn16(x - hi); bf16 or fp16 (clamped to the finite fp16 range)
__device__ __forceinline__ void split16(float x, int f16, unsigned short& hi, unsigned short& lo) {
    if (f16) {
        x = fminf(fmaxf(x, -65504.f), 65504.f);
        __half h = __float2half_rn(x); hi = __half_as_ushort(h); lo = __half_as_ushort(__float2half_rn(x - __half2float(h)));
    } else {
        __nv_bfloat16 h = __float2bfloat16_rn(x); hi = __bfloat16_as_ushort(h); lo = __bfloat16_as_ushort(__float2bfloat16_rn(x - __bfloat162float(h)));
    }
}
// two values -> one packed 32-bit word of a single 16-bit plane (first value in the low half); fp16: saturating to the finite range, which is
// what split16's clamp does.  One F2FP instruction on the device; the host emulation (no inline PTX) goes through split16.
__device__ __forceinline__ uint32_t pack16x2(float a, float b, int f16) {
#ifdef __CUDA_ARCH__
    uint32_t u;
    if (f16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(b), "f"(a));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(b), "f"(a));
    return u;
#else
    unsigned short h0, l0, h1, l1; split16(a, f16, h0, l0); split16(b, f16, h1, l1);
    return (uint32_t)h0 | ((uint32_t)h1 << 16);
#endif
}
// 4 consecutive values -> packed hi (and lo) 8-byte stores
__device__ __forceinline__ void store_planes4(__nv_bfloat16* hi, __nv_bfloat16* lo, int f16, float a, float b, float c, float d) {
    if (!lo) { *reinterpret_cast<uint2*>(hi) = make_uint2(pack16x2(a, b, f16), pack16x2(c, d, f16)); return; }
    unsigned short h[4], l[4];
    split16(a, f16, h[0], l[0]); split16(b, f16, h[1], l[1]); split16(c, f16, h[2], l[2]); split16(d, f16, h[3], l[3]);
    *reinterpret_cast<uint2*>(hi) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
    if (lo) *reinterpret_cast<uint2*>(lo) = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
}

// ---- implicit-GEMM description: out[M,N] = epi( A[M,K] * W[N,K]^T ) -----------------------------
struct GemmA {
    const float* x = nullptr; int ld = 0;   // source NHWC view [B*Hs*Ws, Cin]
    int B = 1, Hs = 1, Ws = 1, Cin = 0;
    int Ho = 1, Wo = 1;                     // output grid; M = B*Ho*Wo
    int ksize = 1, stride = 1, ups = 0;     // 3x3 (pad 1) or 1x1; stride 1|2; ups=1: nearest-2x upsample before the conv
    int M() const { return B * Ho * Wo; }
    int K() const { return ksize * ksize * Cin; }
};
enum { ACT_NONE = 0, ACT_SILU = 1, ACT_GEGLU = 2, ACT_QUICKGELU = 3,        // QuickGELU: x * sigmoid(1.702 x) (custom_clip/model.py:159-161)
       ACT_XATTN = 4 };   // tensor-core engine only: the 32-column chunk of a row IS one head's query -> attend to the xk cached context keys/values
struct GemmEpi {
    const float* bias = nullptr;                                   // [N]
    const float* rowvec = nullptr; int rowvec_ld = 0; int rows_per_batch = 1;   // + rowvec[(m / rows_per_batch)*rowvec_ld + n]
    const float* res = nullptr; int res_ld = 0;                    // + res[m*res_ld + n]
    int act = ACT_NONE;                                            // GEGLU: columns (2j,2j+1) = (a_j, gate_j) -> out col j
    // ACT_XATTN (cross-attention fused into the to_q projection, attention.py:46-74 with d_head 32): keys at xkv[(b*xk + j)*xkv_ld + n],
    // values at + xv_off, b = m / rows_per_batch; out[m, head*32 + i] = sum_j softmax_j(q . k_j * xscale) v_j[i]
    const float* xkv = nullptr; int xkv_ld = 0, xv_off = 0, xk = 0; float xscale = 1.f;
    float* out = nullptr; int out_ld = 0;
    // tensor-core engine only: accumulate per-(image, column) {sum, sumsq} of the fp32 result into stats (see View::st); images are
    // stats_hw consecutive rows.  *stats_fused is set to 1 by gemm_tc when the launch does it (not with split-K or odd shapes).
    double* stats = nullptr; int stats_ld = 0, stats_hw = 0; int* stats_fused = nullptr;
};
// fp32 CUDA-core engine (strict mode / fallback).  W: [N, K] row-major, K ordered (tap, cin).
int gemm_simt(const GemmA& a, const float* W, int N, const GemmEpi& e, cudaStream_t st);

// ---- glue kernels -------------------------------------------------------------------------------
// x NCHW [Bsrc,C,H,W] -> NHWC rows [Bout*H*W, C]: output image b reads source image (b + b_off) modulo Bsrc (CFG doubling, ddim.py:233;
// b_off = first batch row of a chain)
int k_nchw_to_nhwc(const float* x, int Bsrc, int Bout, int C, int H, int W, View out, cudaStream_t st, int b_off = 0);
int k_nhwc_to_nchw(View in, int B, int C, int H, int W, float* out, cudaStream_t st);
// timestep_embedding (ldm util): t int64 [B] -> [B, dim] = [cos | sin]
int k_timestep_embedding(const long long* t, int B, int dim, float* out, cudaStream_t st);
// first convolution (3x3, pad 1, stride 1, Cin <= 4): direct CUDA-core form, out = conv(x) + bias in fp32.  w: [Cout][9][Cin] (tap-major K)
bool k_conv_first_supported(int Cin, int Cout);
int k_conv_first(View x, int B, int H, int W, const float* w, const float* bias, int Cout, View out, cudaStream_t st);
// GroupNorm statistics: sums[b][g] = (sum, sumsq) in fp64 (buffer must be zeroed), over x [B*HW, C]
int k_gn_stats(View x, int B, int HW, int groups, double* sums, cudaStream_t st);
// y = (x-mean)*rstd*gamma+beta, optional SiLU; y is a contiguous-or-strided fp32 view
int k_gn_apply(View x, int B, int HW, int groups, const double* sums, float eps, const float* gamma, const float* beta,
               int silu, Out4 y, Out4 raw, cudaStream_t st);      // raw (optional): the un-normalised x re-emitted (skip 1x1 conv operand)
// One-launch GroupNorm for the tensor-core engine modes (gn_fused.cu): the CTAs of an image form a cluster and exchange their partial sums
// through distributed shared memory; or, chan != nullptr, the group sums are folded from per-(image, channel) {sum, sumsq} pairs that the
// producing GEMM's epilogue accumulated (image stride chan_ld channels) and x is read once.
bool k_gn_fused_supported(int C, int HW, int groups, bool pre);
int k_gn_fused(View x, int B, int HW, int groups, const double* chan, int chan_ld, float eps, const float* gamma, const float* beta, int silu,
               Out4 y, Out4 raw, cudaStream_t st);
// LayerNorm over the last dim of [M, C] (eps 1e-5), one warp per row
int k_layernorm(View x, int M, const float* gamma, const float* beta, float eps, Out4 y, cudaStream_t st);
// softmax(q k^T * scale) v for heads of width 32.  q: [B*Nq, heads*32] view, k/v: [B*Nk, heads*32] views.
int k_attention(View q, View k, View v, int B, int Nq, int Nk, int heads, float scale, Out4 out, cudaStream_t st);
// Tensor-core variant for the f16 engine modes (attention_mma.cu): q/k/v are column blocks of an fp16 plane with row stride ld
// (the QKV GEMM writes that plane directly); warp-level m16n8k16 MMAs, fp32 online softmax.  Nq, Nk multiples of 64.
bool k_attention_mma_supported(int Nq, int Nk);
int k_attention_mma(const __half* q, const __half* k, const __half* v, int ld, int B, int Nq, int Nk, int heads, float scale, Out4 out, cudaStream_t st);
// same for heads of width 64 (CLIP), optional causal mask (key j visible to query i iff j <= i; custom_clip/model.py:287-292)
int k_attention_d64(View q, View k, View v, int B, int N, int heads, float scale, int causal, Out4 out, cudaStream_t st);
// first-stage decoder glue: P = softmax(scale * S) rows as an fp16 plane; 16-bit plane transpose (V -> V^T); VQ lookup + post_quant_conv
int k_softmax_rows(const float* s, int ld_s, int M, int N, float scale, __half* out, int ld_o, cudaStream_t st);
int k_transpose_plane(const __half* in, int ld_in, int rows, int cols, __half* out, int ld_out, cudaStream_t st);
int k_vq_quantize(const float* z, int B, int E, int HW, const float* codebook, int n_e, const float* pq_w, const float* pq_b, int Z, int quantize,
                  View out, cudaStream_t st);
// fp32 [M, C] -> bf16 planes
int k_split_planes(View x, long long M, Out4 y, cudaStream_t st);
int k_add_vec(const float* a, const float* b, float* out, int n, cudaStream_t st);
int k_fold_up_weights(const float* w, int N, int C, float* out, cudaStream_t st);
// im2col of a 3x3 / stride 2 / pad 1 conv (ldm Downsample): x NHWC [B,H,W,C] -> [B*Ho*Wo, 9*C] (tap-major), Ho=(H+1)/2
int k_im2col_s2(View x, int B, int H, int W, Out4 y, cudaStream_t st);
// nearest-neighbour 2x upsample (ldm Upsample): x NHWC [B,H,W,C] -> [B,2H,2W,C]
int k_upsample2x(View x, int B, int H, int W, Out4 y, cudaStream_t st);
// out[i] = silu(in[i])
int k_silu(const float* in, float* out, long long n, cudaStream_t st);
// DDIM update with classifier-free guidance (ddim.py:236-238,258-267).  x NHWC [B*HW, C]; eps [2B*HW, C] (cond first) or [B*HW,C] if !cfg.
// coef (device, 8 floats per step): {sqrt_one_minus_at, 1/sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma, 0,0,0}
int k_ddim_update(const float* x, const float* eps, long long n_per_half, int cfg, float scale, const float* coef_dev,
                  const float* noise, float* x_prev, float* pred_x0, cudaStream_t st);
// Table-driven variants for the graph-captured sampling loop: the step index lives in device memory so ONE captured
// graph serves every step.  t_out[0..B2) = timesteps[*step]; coefficients = coef_table + 8*(*step); noise (optional)
// = noise_table + (*step)*n_per_half; k_step_advance increments *step.
int k_fill_timesteps(const long long* timesteps, const int* step, int B2, long long* t_out, cudaStream_t st);
int k_ddim_update_table(const float* x, const float* eps, long long n_per_half, int cfg, float scale, const float* coef_table, const int* step,
                        const float* noise_table, float* x_prev, float* pred_x0, cudaStream_t st);
int k_step_advance(int* step, cudaStream_t st);
