// U-Net glue kernels: layout changes, timestep embedding, GroupNorm, LayerNorm, attention (d_head 32),
// SiLU and the fused CFG + DDIM update.  All fp32 data, fp32 math with fp64 GroupNorm statistics.
// Semantics: SURVEY.md Appendix A (ldm pieces), rdm/modules/attention.py:42-74,92-96,183-196,
// rdm/models/diffusion/ddim.py:232-238,253-267.
#include "kernels.cuh"
#include "ptx.cuh"
#include <math_constants.h>
#include <cstdlib>

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }
__device__ __forceinline__ float silu_fast(float x) { return x * __frcp_rn(1.f + __expf(-x)); }

// 4 consecutive channels of row m starting at column c (c % 4 == 0)
__device__ __forceinline__ void store4(const Out4& o, size_t m, int c, float a, float b, float cc, float d) {
    if (o.f) *reinterpret_cast<float4*>(o.f + m * o.ldf + c) = make_float4(a, b, cc, d);
    if (o.hi) store_planes4(o.hi + m * o.ldb + c, o.lo ? o.lo + m * o.ldb + c : nullptr, o.f16, a, b, cc, d);
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int Bsrc, int Bout, int b_off, int C, int HW, float* __restrict__ out, int ld) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over Bout*HW*C, c fastest
    long long total = (long long)Bout * HW * C;
    if (i >= total) return;
    int c = (int)(i % C);
    long long m = i / C;
    int p = (int)(m % HW), b = ((int)(m / HW) + b_off) % Bsrc;
    out[m * ld + c] = x[((long long)b * C + c) * HW + p];
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, int ld, int B, int C, int HW, float* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over B*C*HW, p fastest
    long long total = (long long)B * C * HW;
    if (i >= total) return;
    int p = (int)(i % HW);
    long long r = i / HW;
    int c = (int)(r % C), b = (int)(r / C);
    out[i] = in[((long long)b * HW + p) * ld + c];
}

__global__ void timestep_embedding_kernel(const long long* __restrict__ t, int B, int dim, float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int half = dim / 2;
    if (i >= B * half) return;
    int b = i / half, j = i % half;
    // freqs = exp(-ln(10000) * j / half) in fp32, args = float(t) * freqs   (same op order as the torch statement)
    float freq = expf(-9.210340371976184f * (float)j / (float)half);
    float arg = (float)t[b] * freq;
    out[(size_t)b * dim + j] = cosf(arg);
    out[(size_t)b * dim + half + j] = sinf(arg);
    if ((dim & 1) && j == 0) out[(size_t)b * dim + dim - 1] = 0.f;
}

// First convolution of the U-Net (openaimodel.py:141-145: conv_nd(dims, in_channels, model_channels, 3, padding=1) with 3 or 4 input
// channels): K = 9 * Cin <= 36 is far too short for either GEMM engine (the generic CUDA-core engine took 65 us for 0.45 GFLOP at 32x32,
// B2 = 32 -- 1.4 % of the forward).  Direct form, register-tiled: a CTA owns TP = PG * 16 consecutive pixels and every output channel.
// Shared memory holds the weights re-laid as [k = tap * Cin + c][Cout (+4 pad)] and the im2col patch of the CTA's pixels [TP][K] (zeros
// for the padding ring); a thread owns 4 output channels x 16 pixels = 16 float4 accumulators, and per 4 k it reads 4 weight float4 and
// 16 patch float4 (broadcast within the lanes of a pixel group) for 256 FMAs -- FMA-bound instead of shared-memory-bound (the first
// version, one pixel x 4 channels per thread, re-read its weights for every pixel: 36 LDS.128 per 144 FMAs, 65 us under ncu).
// Summation order per output: bias, then k ascending -- the order of the previous form and of the reference's fp32 convolution loop nest.
constexpr int CF_PPT = 16;
__global__ void __launch_bounds__(256) conv_first_kernel(const float* __restrict__ x, int ld, int Cin, int B, int H, int W, const float* __restrict__ w,
                                                         const float* __restrict__ bias, int Cout, float* __restrict__ out, int out_ld) {
    extern __shared__ float s_cf[];
    const int K = 9 * Cin, Kp = (K + 3) & ~3, LDW = Cout + 4, G = Cout / 4, PG = blockDim.x / G, TP = PG * CF_PPT;
    float* s_w = s_cf;                                   // [Kp][LDW]
    float* s_x = s_cf + Kp * LDW;                        // [TP][Kp]
    for (int i = threadIdx.x; i < Kp * Cout; i += blockDim.x) { const int o = i / Kp, k = i - o * Kp; s_w[k * LDW + o] = k < K ? w[o * K + k] : 0.f; }
    const long long M = (long long)B * H * W, p0 = (long long)blockIdx.x * TP;
    for (int i = threadIdx.x; i < TP * Kp; i += blockDim.x) {
        const int p = i / Kp, k = i - p * Kp;
        const long long m = p0 + p;
        float v = 0.f;
        if (k < K && m < M) {
            const int tap = k / Cin, c = k - tap * Cin, dy = tap / 3 - 1, dx = tap % 3 - 1;
            const int iy = (int)((m / W) % H) + dy, ix = (int)(m % W) + dx;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[(m + (long long)dy * W + dx) * ld + c];
        }
        s_x[i] = v;
    }
    __syncthreads();
    const int cg = (threadIdx.x % G) * 4, pg = threadIdx.x / G;
    if (pg >= PG) return;
    float4 acc[CF_PPT];
    const float4 b4 = bias ? *reinterpret_cast<const float4*>(bias + cg) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < CF_PPT; p++) acc[p] = b4;
    const float* xp = s_x + pg * CF_PPT * Kp;
    for (int k4 = 0; k4 < Kp; k4 += 4) {
        const float4 w0 = *reinterpret_cast<const float4*>(s_w + (k4 + 0) * LDW + cg), w1 = *reinterpret_cast<const float4*>(s_w + (k4 + 1) * LDW + cg);
        const float4 w2 = *reinterpret_cast<const float4*>(s_w + (k4 + 2) * LDW + cg), w3 = *reinterpret_cast<const float4*>(s_w + (k4 + 3) * LDW + cg);
#pragma unroll
        for (int p = 0; p < CF_PPT; p++) {
            const float4 v = *reinterpret_cast<const float4*>(xp + p * Kp + k4);
            float4 a = acc[p];
            a.x = fmaf(v.x, w0.x, a.x); a.y = fmaf(v.x, w0.y, a.y); a.z = fmaf(v.x, w0.z, a.z); a.w = fmaf(v.x, w0.w, a.w);
            a.x = fmaf(v.y, w1.x, a.x); a.y = fmaf(v.y, w1.y, a.y); a.z = fmaf(v.y, w1.z, a.z); a.w = fmaf(v.y, w1.w, a.w);
            a.x = fmaf(v.z, w2.x, a.x); a.y = fmaf(v.z, w2.y, a.y); a.z = fmaf(v.z, w2.z, a.z); a.w = fmaf(v.z, w2.w, a.w);
            a.x = fmaf(v.w, w3.x, a.x); a.y = fmaf(v.w, w3.y, a.y); a.z = fmaf(v.w, w3.z, a.z); a.w = fmaf(v.w, w3.w, a.w);
            acc[p] = a;
        }
    }
#pragma unroll
    for (int p = 0; p < CF_PPT; p++) {
        const long long m = p0 + pg * CF_PPT + p;
        if (m < M) *reinterpret_cast<float4*>(out + m * out_ld + cg) = acc[p];
    }
}

// grid (row_chunks, B).  Thread t owns the float4 column v = t % V for the whole kernel and walks rows slot, slot+nslots, ...
// (consecutive threads -> consecutive 16-byte pieces of a row: coalesced), keeping its 4 channel sums in fp64 registers; one
// shared-memory atomic per touched group per thread at the end, then fp64 atomics to the global accumulators.
__global__ void gn_stats_kernel(const float* __restrict__ x, int ld, int C, int HW, int groups, int rows_per_cta, double* __restrict__ sums) {
    pdl_wait();
    pdl_launch_dependents();
    // All accumulation in fp64: the variance is E[x^2] - mean^2, and with fp32 partial sums (whose atomics also commit in a varying order)
    // a group whose |mean| is several times its spread lost ~1e-4 of rstd -- run-to-run noise as large as the fp16 operand rounding itself
    // (seen on the first-stage decoder).  The kernel is L2/HBM-bound; 8 DFMA per 16 bytes are free on B200.
    extern __shared__ double s_acc[];          // [2][groups]
    const int b = blockIdx.y, V = C / 4, cpg = C / groups;
    for (int i = threadIdx.x; i < 2 * groups; i += blockDim.x) s_acc[i] = 0.0;
    __syncthreads();
    const int nslots = blockDim.x / V;          // blockDim.x is a multiple of V or smaller threads idle
    const int v = threadIdx.x % V, slot = threadIdx.x / V;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(HW, r0 + rows_per_cta);
    if (slot < nslots) {
        const float* base = x + ((long long)b * HW) * ld + v * 4;
        double s[4] = {0.0, 0.0, 0.0, 0.0}, ss[4] = {0.0, 0.0, 0.0, 0.0};
        for (int r = r0 + slot; r < r1; r += nslots) {
            const float4 q = *reinterpret_cast<const float4*>(base + (long long)r * ld);
            const double q0 = q.x, q1 = q.y, q2 = q.z, q3 = q.w;
            s[0] += q0; ss[0] = fma(q0, q0, ss[0]); s[1] += q1; ss[1] = fma(q1, q1, ss[1]);
            s[2] += q2; ss[2] = fma(q2, q2, ss[2]); s[3] += q3; ss[3] = fma(q3, q3, ss[3]);
        }
        int g = (v * 4) / cpg; double gs = 0.0, gss = 0.0;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            int gt = (v * 4 + t) / cpg;
            if (gt != g) { atomicAdd(&s_acc[g], gs); atomicAdd(&s_acc[groups + g], gss); g = gt; gs = 0.0; gss = 0.0; }
            gs += s[t]; gss += ss[t];
        }
        atomicAdd(&s_acc[g], gs); atomicAdd(&s_acc[groups + g], gss);
    }
    __syncthreads();
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        atomicAdd(&sums[((size_t)b * groups + g) * 2 + 0], s_acc[g]);
        atomicAdd(&sums[((size_t)b * groups + g) * 2 + 1], s_acc[groups + g]);
    }
}

// grid (row_chunks, B), same thread->column mapping as gn_stats: a thread owns float4 column v for the whole kernel (group ids, gamma,
// beta and the finalised mean/rstd are loaded once) and walks rows slot, slot+nslots, ...; no per-element integer division.
__global__ void gn_apply_kernel(const float* __restrict__ x, int ld, int C, int HW, int groups, const double* __restrict__ sums, float eps,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int silu, Out4 y, Out4 raw, int rows_per_cta) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float s_mean[64], s_rstd[64];
    const int b = blockIdx.y, V = C / 4, cpg = C / groups;
    if (threadIdx.x < groups) {
        const double cnt = (double)HW * cpg;
        double s = sums[((size_t)b * groups + threadIdx.x) * 2], ss = sums[((size_t)b * groups + threadIdx.x) * 2 + 1];
        double mean = s / cnt, var = ss / cnt - mean * mean;
        s_mean[threadIdx.x] = (float)mean;
        s_rstd[threadIdx.x] = (float)(1.0 / sqrt((var > 0 ? var : 0.0) + (double)eps));
    }
    __syncthreads();
    const int nslots = blockDim.x / V, v = threadIdx.x % V, slot = threadIdx.x / V;
    if (slot >= nslots) return;
    float sc[4], sh[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const int c = v * 4 + t, g = c / cpg;
        sc[t] = s_rstd[g] * gamma[c];
        sh[t] = beta[c] - s_mean[g] * sc[t];
    }
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(HW, r0 + rows_per_cta);
    const size_t m0 = (size_t)b * HW;
    const bool want_raw = raw.any();
    const bool fast = y.f == nullptr && y.lo == nullptr;   // single 16-bit plane output (2^-11 / 2^-8 rounding follows): MUFU exp + reciprocal are exact enough
    // (no unroll pragma: a thread walks only 4-8 rows; `#pragma unroll 4` here and in gn_stats_kernel measured 4.23 vs 4.16 ms per forward)
    for (int r = r0 + slot; r < r1; r += nslots) {
        const float4 q = *reinterpret_cast<const float4*>(x + (m0 + r) * ld + v * 4);
        float o0 = fmaf(q.x, sc[0], sh[0]), o1 = fmaf(q.y, sc[1], sh[1]), o2 = fmaf(q.z, sc[2], sh[2]), o3 = fmaf(q.w, sc[3], sh[3]);
        if (silu) {
            if (fast) { o0 = silu_fast(o0); o1 = silu_fast(o1); o2 = silu_fast(o2); o3 = silu_fast(o3); }
            else { o0 = silu_f(o0); o1 = silu_f(o1); o2 = silu_f(o2); o3 = silu_f(o3); }
        }
        store4(y, m0 + r, v * 4, o0, o1, o2, o3);
        if (want_raw) store4(raw, m0 + r, v * 4, q.x, q.y, q.z, q.w);
    }
}

// one warp per row; the row is read ONCE into registers (<= 8 float4 per lane, C <= 1024), then mean and the centred variance
// (two-pass formula, fp32) come from registers.  Rows wider than 1024 fall back to re-reading through L1.
__global__ void layernorm_kernel(const float* __restrict__ x, int ld, int C, int M, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float eps, Out4 y) {
    pdl_wait();
    pdl_launch_dependents();
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* xr = x + (size_t)row * ld;
    const int V = C / 4;
    if (V <= 256) {
        float4 q[8];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int v = lane + 32 * i;
            q[i] = v < V ? *reinterpret_cast<const float4*>(xr + v * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            s += (q[i].x + q[i].y) + (q[i].z + q[i].w);
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(FULL, s, m);
        const float mean = s / (float)C;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (lane + 32 * i < V) {
                float a = q[i].x - mean, b = q[i].y - mean, c = q[i].z - mean, d = q[i].w - mean;
                ss += (a * a + b * b) + (c * c + d * d);
            }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) ss += __shfl_xor_sync(FULL, ss, m);
        const float rstd = rsqrtf(ss / (float)C + eps);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int v = lane + 32 * i;
            if (v < V) {
                float4 g = __ldg(reinterpret_cast<const float4*>(gamma + v * 4)), bb = __ldg(reinterpret_cast<const float4*>(beta + v * 4));
                store4(y, (size_t)row, v * 4, (q[i].x - mean) * rstd * g.x + bb.x, (q[i].y - mean) * rstd * g.y + bb.y,
                       (q[i].z - mean) * rstd * g.z + bb.z, (q[i].w - mean) * rstd * g.w + bb.w);
            }
        }
        return;
    }
    float s = 0.f;
    for (int v = lane; v < V; v += 32) { float4 q = *reinterpret_cast<const float4*>(xr + v * 4); s += (q.x + q.y) + (q.z + q.w); }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(FULL, s, m);
    const float mean = s / (float)C;
    float ss = 0.f;
    for (int v = lane; v < V; v += 32) {
        float4 q = *reinterpret_cast<const float4*>(xr + v * 4);
        float a = q.x - mean, b = q.y - mean, c = q.z - mean, d = q.w - mean;
        ss += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) ss += __shfl_xor_sync(FULL, ss, m);
    const float rstd = rsqrtf(ss / (float)C + eps);
    for (int v = lane; v < V; v += 32) {
        float4 q = *reinterpret_cast<const float4*>(xr + v * 4);
        float4 g = *reinterpret_cast<const float4*>(gamma + v * 4), bb = *reinterpret_cast<const float4*>(beta + v * 4);
        store4(y, (size_t)row, v * 4, (q.x - mean) * rstd * g.x + bb.x, (q.y - mean) * rstd * g.y + bb.y,
               (q.z - mean) * rstd * g.z + bb.z, (q.w - mean) * rstd * g.w + bb.w);
    }
}

// Attention for d_head = DH (32: U-Net, 64: CLIP): a thread owns RQ query rows, keys/values of one (batch, head) are staged through
// shared memory in tiles of KT keys and read as broadcasts (RQ rows per thread divide the shared-memory traffic per FMA by RQ);
// online softmax in fp32.  grid (ceil(Nq / (blockDim*RQ)), heads, B).  causal: key j contributes to query i only if j <= i.
constexpr int KT = 64;
template <int DH, int RQ>
__global__ void attention_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk, const float* __restrict__ v, int ldv,
                                 int Nq, int Nk, float scale_log2e, int causal, Out4 out) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int V4 = DH / 4;
    __shared__ float4 sk[KT][V4], sv[KT][V4];
    const int b = blockIdx.z, h = blockIdx.y;
    int qi[RQ]; bool active[RQ];
    float qr[RQ][DH], acc[RQ][DH], mx[RQ], l[RQ];
#pragma unroll
    for (int u = 0; u < RQ; u++) {
        qi[u] = (blockIdx.x * RQ + u) * blockDim.x + threadIdx.x;          // rows of one thread are blockDim apart: coalescing as before
        active[u] = qi[u] < Nq;
        mx[u] = -CUDART_INF_F; l[u] = 0.f;
        if (active[u]) {
            const float4* qp = reinterpret_cast<const float4*>(q + ((size_t)b * Nq + qi[u]) * ldq + h * DH);
#pragma unroll
            for (int i = 0; i < V4; i++) { float4 t = qp[i]; qr[u][4 * i] = t.x * scale_log2e; qr[u][4 * i + 1] = t.y * scale_log2e; qr[u][4 * i + 2] = t.z * scale_log2e; qr[u][4 * i + 3] = t.w * scale_log2e; }
        } else {
#pragma unroll
            for (int i = 0; i < DH; i++) qr[u][i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < DH; i++) acc[u][i] = 0.f;
    }
    int qmax = 0;
#pragma unroll
    for (int u = 0; u < RQ; u++) qmax = max(qmax, qi[u]);
    for (int k0 = 0; k0 < Nk; k0 += KT) {
        const int kn = min(KT, Nk - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < kn * V4; i += blockDim.x) {
            int j = i / V4, c = i % V4;
            sk[j][c] = *reinterpret_cast<const float4*>(k + ((size_t)b * Nk + k0 + j) * ldk + h * DH + c * 4);
            sv[j][c] = *reinterpret_cast<const float4*>(v + ((size_t)b * Nk + k0 + j) * ldv + h * DH + c * 4);
        }
        __syncthreads();
        const int jend = causal ? min(kn, qmax - k0 + 1) : kn;
        for (int j = 0; j < jend; j++) {
            float s[RQ];
#pragma unroll
            for (int u = 0; u < RQ; u++) s[u] = 0.f;
#pragma unroll
            for (int c = 0; c < V4; c++) {
                const float4 kk = sk[j][c];
#pragma unroll
                for (int u = 0; u < RQ; u++) {
                    s[u] = fmaf(qr[u][4 * c], kk.x, s[u]); s[u] = fmaf(qr[u][4 * c + 1], kk.y, s[u]);
                    s[u] = fmaf(qr[u][4 * c + 2], kk.z, s[u]); s[u] = fmaf(qr[u][4 * c + 3], kk.w, s[u]);
                }
            }
            float p[RQ];
#pragma unroll
            for (int u = 0; u < RQ; u++) {                 // s is the logit in log2 units
                const bool vis = !causal || (k0 + j <= qi[u]);
                if (vis && s[u] > mx[u]) {
                    float corr = exp2f(mx[u] - s[u]);
                    l[u] *= corr;
#pragma unroll
                    for (int i = 0; i < DH; i++) acc[u][i] *= corr;
                    mx[u] = s[u];
                }
                p[u] = vis ? exp2f(s[u] - mx[u]) : 0.f;
                l[u] += p[u];
            }
#pragma unroll
            for (int c = 0; c < V4; c++) {
                const float4 vv = sv[j][c];
#pragma unroll
                for (int u = 0; u < RQ; u++) {
                    acc[u][4 * c] = fmaf(p[u], vv.x, acc[u][4 * c]); acc[u][4 * c + 1] = fmaf(p[u], vv.y, acc[u][4 * c + 1]);
                    acc[u][4 * c + 2] = fmaf(p[u], vv.z, acc[u][4 * c + 2]); acc[u][4 * c + 3] = fmaf(p[u], vv.w, acc[u][4 * c + 3]);
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < RQ; u++) {
        if (active[u]) {
            const float inv = 1.f / l[u];
#pragma unroll
            for (int i = 0; i < V4; i++) store4(out, (size_t)b * Nq + qi[u], h * DH + 4 * i, acc[u][4 * i] * inv, acc[u][4 * i + 1] * inv, acc[u][4 * i + 2] * inv, acc[u][4 * i + 3] * inv);
        }
    }
}

__global__ void split_planes_kernel(const float* __restrict__ x, int ld, int C, long long total4, Out4 y) {
    pdl_wait();
    pdl_launch_dependents();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int V = C / 4; int v = (int)(i % V); long long m = i / V;
    float4 q = *reinterpret_cast<const float4*>(x + m * ld + v * 4);
    store4(y, (size_t)m, v * 4, q.x, q.y, q.z, q.w);
}

__global__ void im2col_s2_kernel(const float* __restrict__ x, int ld, int C, int B, int H, int W, int Ho, int Wo, long long total4, Out4 y) {
    pdl_wait();
    pdl_launch_dependents();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // over Mo * 9 * C/4
    if (i >= total4) return;
    const int V = C / 4; int v = (int)(i % V); long long r = i / V;
    int tap = (int)(r % 9); long long mo = r / 9;
    int ox = (int)(mo % Wo); long long t = mo / Wo; int oy = (int)(t % Ho), b = (int)(t / Ho);
    int iy = oy * 2 + tap / 3 - 1, ix = ox * 2 + tap % 3 - 1;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) q = *reinterpret_cast<const float4*>(x + ((size_t)(b * H + iy) * W + ix) * ld + v * 4);
    store4(y, (size_t)mo, tap * C + v * 4, q.x, q.y, q.z, q.w);
}

__global__ void upsample2x_kernel(const float* __restrict__ x, int ld, int C, int B, int H, int W, long long total4, Out4 y) {
    pdl_wait();
    pdl_launch_dependents();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // over B*2H*2W * C/4
    if (i >= total4) return;
    const int V = C / 4; int v = (int)(i % V); long long mo = i / V;
    int ox = (int)(mo % (2 * W)); long long t = mo / (2 * W); int oy = (int)(t % (2 * H)), b = (int)(t / (2 * H));
    float4 q = *reinterpret_cast<const float4*>(x + ((size_t)(b * H + (oy >> 1)) * W + (ox >> 1)) * ld + v * 4);
    store4(y, (size_t)mo, v * 4, q.x, q.y, q.z, q.w);
}

__global__ void silu_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = silu_f(in[i]);
}

// Same operation order and rounding as the reference's float32 tensor expressions (no FMA contraction).
__global__ void ddim_update_kernel(const float* __restrict__ x, const float* __restrict__ eps, long long n, int cfg, float scale,
                                   const float* __restrict__ coef, const float* __restrict__ noise, float* __restrict__ x_prev, float* __restrict__ pred_x0) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float s1m = coef[0], sqrt_at = coef[1], sqrt_aprev = coef[2], dirc = coef[3], sigma = coef[4];
    float e = eps[i];
    if (cfg) { float eu = eps[n + i]; e = __fadd_rn(eu, __fmul_rn(scale, __fsub_rn(e, eu))); }      // e_u + s*(e_c - e_u), ddim.py:238
    float xv = x[i];
    float p0 = __fdiv_rn(__fsub_rn(xv, __fmul_rn(s1m, e)), sqrt_at);                                 // ddim.py:259
    float xp = __fadd_rn(__fmul_rn(sqrt_aprev, p0), __fmul_rn(dirc, e));                             // ddim.py:263,267
    if (noise) xp = __fadd_rn(xp, __fmul_rn(sigma, noise[i]));
    x_prev[i] = xp;
    if (pred_x0) pred_x0[i] = p0;
}

__global__ void fill_timesteps_kernel(const long long* __restrict__ timesteps, const int* __restrict__ step, int B2, long long* __restrict__ t_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B2) t_out[i] = timesteps[*step];
}
__global__ void ddim_update_table_kernel(const float* __restrict__ x, const float* __restrict__ eps, long long n, int cfg, float scale,
                                         const float* __restrict__ coef_table, const int* __restrict__ step, const float* __restrict__ noise_table,
                                         float* __restrict__ x_prev, float* __restrict__ pred_x0) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = *step;
    const float* coef = coef_table + 8 * s;
    const float s1m = coef[0], sqrt_at = coef[1], sqrt_aprev = coef[2], dirc = coef[3], sigma = coef[4];
    float e = eps[i];
    if (cfg) { float eu = eps[n + i]; e = __fadd_rn(eu, __fmul_rn(scale, __fsub_rn(e, eu))); }
    float xv = x[i];
    float p0 = __fdiv_rn(__fsub_rn(xv, __fmul_rn(s1m, e)), sqrt_at);
    float xp = __fadd_rn(__fmul_rn(sqrt_aprev, p0), __fmul_rn(dirc, e));
    if (noise_table) xp = __fadd_rn(xp, __fmul_rn(sigma, noise_table[(size_t)s * n + i]));
    x_prev[i] = xp;
    if (pred_x0) pred_x0[i] = p0;
}
__global__ void step_advance_kernel(int* step) { *step += 1; }

inline int blocks_for(long long n, int t) { return (int)((n + t - 1) / t); }

}  // namespace


// ---- first-stage (VQ) decoder glue: single-head attention over all pixels runs as two tensor-core GEMMs with these in between ------------
// P[r, :] = softmax(scale * S[r, :]) as an fp16 plane; one CTA per row, the row lives in registers (N <= 256 threads * 16 values)
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, int ld_s, int N, float scale_log2e, __half* __restrict__ out, int ld_o) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float red[8];
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float4* src = reinterpret_cast<const float4*>(s + (size_t)row * ld_s);
    const int V = N >> 2;
    float4 q[4];
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int v = tid + 256 * i;
        q[i] = v < V ? src[v] : make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
        mx = fmaxf(mx, fmaxf(fmaxf(q[i].x, q[i].y), fmaxf(q[i].z, q[i].w)));
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, m));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 8; w++) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        q[i].x = exp2f((q[i].x - mx) * scale_log2e); q[i].y = exp2f((q[i].y - mx) * scale_log2e);
        q[i].z = exp2f((q[i].z - mx) * scale_log2e); q[i].w = exp2f((q[i].w - mx) * scale_log2e);
        sum += (q[i].x + q[i].y) + (q[i].z + q[i].w);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) sum += __shfl_xor_sync(FULL, sum, m);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) sum += red[w];
    const float inv = 1.f / sum;
    __half* dst = out + (size_t)row * ld_o;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int v = tid + 256 * i;
        if (v < V) {
            const __half2 a = __floats2half2_rn(q[i].x * inv, q[i].y * inv), b = __floats2half2_rn(q[i].z * inv, q[i].w * inv);
            *reinterpret_cast<uint2*>(dst + v * 4) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
        }
    }
}

// out[c][r] = in[r][c] for a 16-bit plane (V -> V^T: the B operand of the P V product must be K-major); 32 x 32 tiles through shared memory
__global__ void transpose_plane_kernel(const unsigned short* __restrict__ in, int ld_in, int rows, int cols, unsigned short* __restrict__ out, int ld_out) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ unsigned short tile[32][34];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 256 threads: 8 rows per pass
    for (int j = ty; j < 32; j += 8) if (r0 + j < rows && c0 + tx < cols) tile[j][tx] = in[(size_t)(r0 + j) * ld_in + c0 + tx];
    __syncthreads();
    for (int j = ty; j < 32; j += 8) if (c0 + j < cols && r0 + tx < rows) out[(size_t)(c0 + j) * ld_out + r0 + tx] = tile[tx][j];
}

// VectorQuantizer lookup + post_quant_conv (taming VectorQuantizer2 / ldm VQModelInterface.decode): for every pixel
// of z NCHW [B, E, H, W] (E <= 4) the nearest codebook row by d = |z|^2 + |e|^2 - 2 z.e (first minimum wins), then the 1x1 post_quant_conv
// [Z, E] (+bias) -> NHWC rows [B*H*W, ld] (ld >= Z).  quantize == 0 skips the lookup (force_not_quantize).  The codebook is streamed through
// shared memory in chunks that every thread of the CTA scans.
__global__ void __launch_bounds__(256) vq_quantize_kernel(const float* __restrict__ z, int B, int E, int HW, const float* __restrict__ codebook, int n_e,
                                                          const float* __restrict__ pq_w, const float* __restrict__ pq_b, int Z, int quantize,
                                                          float* __restrict__ out, int ld) {
    __shared__ float4 s_code[1024];
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x, total = (long long)B * HW;
    const bool ok = pix < total;
    float zv[4] = {0.f, 0.f, 0.f, 0.f};
    if (ok) { const long long b = pix / HW, p = pix % HW; for (int e = 0; e < E; e++) zv[e] = z[(b * E + e) * HW + p]; }
    float zq[4] = {zv[0], zv[1], zv[2], zv[3]};
    if (quantize) {
        const float zz = ((zv[0] * zv[0] + zv[1] * zv[1]) + zv[2] * zv[2]) + zv[3] * zv[3];
        float best = CUDART_INF_F; int bi = 0;
        for (int c0 = 0; c0 < n_e; c0 += 1024) {
            __syncthreads();
            for (int i = threadIdx.x; i < 1024 && c0 + i < n_e; i += blockDim.x) {
                float4 e4 = make_float4(0.f, 0.f, 0.f, 0.f);
                const float* cr = codebook + (size_t)(c0 + i) * E;
                e4.x = cr[0]; if (E > 1) e4.y = cr[1]; if (E > 2) e4.z = cr[2]; if (E > 3) e4.w = cr[3];
                s_code[i] = e4;
            }
            __syncthreads();
            const int cn = n_e - c0 < 1024 ? n_e - c0 : 1024;
            for (int i = 0; i < cn; i++) {
                const float4 e4 = s_code[i];
                const float ee = ((e4.x * e4.x + e4.y * e4.y) + e4.z * e4.z) + e4.w * e4.w;
                const float dot = ((zv[0] * e4.x + zv[1] * e4.y) + zv[2] * e4.z) + zv[3] * e4.w;
                const float d = (zz + ee) - 2.f * dot;
                if (d < best) { best = d; bi = c0 + i; }
            }
        }
        const float* cr = codebook + (size_t)bi * E;
        for (int e = 0; e < E; e++) zq[e] = cr[e];
    }
    if (!ok) return;
    for (int o = 0; o < Z; o++) {
        float acc = pq_b ? pq_b[o] : 0.f;
        for (int e = 0; e < E; e++) acc = fmaf(pq_w[o * E + e], zq[e], acc);
        out[pix * ld + o] = acc;
    }
    for (int o = Z; o < ld; o++) out[pix * ld + o] = 0.f;
}

#define LAUNCH_CHECK() do { RDM_COUNT_LAUNCH(); RDM_CHECK_CUDA(cudaGetLastError()); } while (0)

int k_nchw_to_nhwc(const float* x, int Bsrc, int Bout, int C, int H, int W, View out, cudaStream_t st, int b_off) {
    long long n = (long long)Bout * H * W * C;
    nchw_to_nhwc_kernel<<<blocks_for(n, 256), 256, 0, st>>>(x, Bsrc, Bout, b_off, C, H * W, out.p, out.ld);
    LAUNCH_CHECK(); return RDM_OK;
}
int k_nhwc_to_nchw(View in, int B, int C, int H, int W, float* out, cudaStream_t st) {
    long long n = (long long)B * H * W * C;
    nhwc_to_nchw_kernel<<<blocks_for(n, 256), 256, 0, st>>>(in.p, in.ld, B, C, H * W, out);
    LAUNCH_CHECK(); return RDM_OK;
}
int k_timestep_embedding(const long long* t, int B, int dim, float* out, cudaStream_t st) {
    timestep_embedding_kernel<<<blocks_for((long long)B * (dim / 2), 128), 128, 0, st>>>(t, B, dim, out);
    LAUNCH_CHECK(); return RDM_OK;
}
// block = G * PG threads (G = Cout / 4 channel groups, PG pixel groups of 16 pixels); shared memory: weights + patch
static int conv_first_pg(int Cout) { const int G = Cout / 4; int pg = 256 / G; return pg > 4 ? 4 : pg; }
static size_t conv_first_smem(int Cin, int Cout) { const int Kp = (9 * Cin + 3) & ~3; return ((size_t)Kp * (Cout + 4) + (size_t)conv_first_pg(Cout) * CF_PPT * Kp) * sizeof(float); }
bool k_conv_first_supported(int Cin, int Cout) { return Cin >= 1 && Cin <= 4 && Cout % 4 == 0 && Cout >= 4 && Cout / 4 <= 256 && conv_first_smem(Cin, Cout) <= 48 * 1024; }
int k_conv_first(View x, int B, int H, int W, const float* w, const float* bias, int Cout, View out, cudaStream_t st) {
    RDM_REQUIRE(k_conv_first_supported(x.C, Cout) && out.ld % 4 == 0, RDM_ERR_UNSUPPORTED, "conv_first: Cin=%d Cout=%d", x.C, Cout);
    const int PG = conv_first_pg(Cout), TP = PG * CF_PPT;
    const long long M = (long long)B * H * W;
    conv_first_kernel<<<(unsigned)((M + TP - 1) / TP), (Cout / 4) * PG, conv_first_smem(x.C, Cout), st>>>(x.p, x.ld, x.C, B, H, W, w, bias, Cout, out.p, out.ld);
    LAUNCH_CHECK(); return RDM_OK;
}
int k_gn_stats(View x, int B, int HW, int groups, double* sums, cudaStream_t st) {
    RDM_REQUIRE(x.C % 4 == 0 && x.C % groups == 0 && x.ld % 4 == 0, RDM_ERR_ARG, "gn_stats: C=%d ld=%d groups=%d", x.C, x.ld, groups);
    RDM_REQUIRE(groups <= 64, RDM_ERR_UNSUPPORTED, "gn_stats: groups > 64");
    const int V = x.C / 4;
    int threads = V >= 256 ? ((V + 31) / 32) * 32 : (256 / V) * V;             // whole rows of float4 columns per pass
    if (threads > 1024) threads = 1024;
    RDM_REQUIRE(V <= 1024, RDM_ERR_UNSUPPORTED, "gn_stats: C=%d too wide", x.C);
    // enough CTAs to fill the machine, at least ~8 rows per thread slot
    int nslots = threads / V; if (nslots < 1) nslots = 1;
    static const int rows_pt = getenv("RDM_GN_ROWS_STATS") ? atoi(getenv("RDM_GN_ROWS_STATS")) : 8, cta_cap = getenv("RDM_GN_CTAS") ? atoi(getenv("RDM_GN_CTAS")) : 8;
    int rows_per_cta = nslots * rows_pt;
    int chunks = (HW + rows_per_cta - 1) / rows_per_cta;
    while (chunks * B > 148 * cta_cap && rows_per_cta < HW) { rows_per_cta *= 2; chunks = (HW + rows_per_cta - 1) / rows_per_cta; }
    RDM_CHECK_CUDA(launch_pdl(gn_stats_kernel, dim3(chunks, B), dim3(threads), 2 * groups * sizeof(double), st, (const float*)x.p, x.ld, x.C, HW, groups, rows_per_cta, sums));
    LAUNCH_CHECK(); return RDM_OK;
}
int k_gn_apply(View x, int B, int HW, int groups, const double* sums, float eps, const float* gamma, const float* beta, int silu, Out4 y, Out4 raw, cudaStream_t st) {
    RDM_REQUIRE(x.C % 4 == 0 && y.ldf % 4 == 0 && y.ldb % 4 == 0, RDM_ERR_ARG, "gn_apply: alignment");
    RDM_REQUIRE(groups <= 64 && x.C / 4 <= 1024, RDM_ERR_UNSUPPORTED, "gn_apply: groups > 64 or C too wide");
    const int V = x.C / 4;
    int threads = V >= 256 ? ((V + 31) / 32) * 32 : (256 / V) * V;
    if (threads < groups) threads = ((groups + 31) / 32) * 32;
    int nslots = threads / V; if (nslots < 1) nslots = 1;
    static const int rows_pt = getenv("RDM_GN_ROWS_APPLY") ? atoi(getenv("RDM_GN_ROWS_APPLY")) : 4, cta_cap = getenv("RDM_GN_CTAS_APPLY") ? atoi(getenv("RDM_GN_CTAS_APPLY")) : 16;
    int rows_per_cta = nslots * rows_pt;
    int chunks = (HW + rows_per_cta - 1) / rows_per_cta;
    while (chunks * B > 148 * cta_cap && rows_per_cta < HW) { rows_per_cta *= 2; chunks = (HW + rows_per_cta - 1) / rows_per_cta; }
    RDM_CHECK_CUDA(launch_pdl(gn_apply_kernel, dim3(chunks, B), dim3(threads), 0, st, (const float*)x.p, x.ld, x.C, HW, groups, sums, eps, gamma, beta, silu, y, raw, rows_per_cta));
    LAUNCH_CHECK(); return RDM_OK;
}
int k_layernorm(View x, int M, const float* gamma, const float* beta, float eps, Out4 y, cudaStream_t st) {
    RDM_REQUIRE(x.C % 4 == 0, RDM_ERR_ARG, "layernorm: C %% 4");
    RDM_CHECK_CUDA(launch_pdl(layernorm_kernel, dim3(blocks_for(M, 8)), dim3(256), 0, st, (const float*)x.p, x.ld, x.C, M, gamma, beta, eps, y));
    LAUNCH_CHECK(); return RDM_OK;
}
int k_attention(View q, View k, View v, int B, int Nq, int Nk, int heads, float scale, Out4 out, cudaStream_t st) {
    RDM_REQUIRE(q.C == heads * 32, RDM_ERR_UNSUPPORTED, "attention: only d_head=32 is implemented (C=%d heads=%d)", q.C, heads);
    if (Nq >= 256 && Nk >= 64) {                     // self-attention at the large levels: 2 query rows per thread
        dim3 grid((Nq + 255) / 256, heads, B);
        RDM_CHECK_CUDA(launch_pdl(attention_kernel<32, 2>, grid, dim3(128), 0, st, (const float*)q.p, q.ld, (const float*)k.p, k.ld, (const float*)v.p, v.ld, Nq, Nk, scale * 1.4426950408889634f, 0, out));
    } else {
        int threads = Nq >= 128 ? 128 : ((Nq + 31) / 32) * 32;
        dim3 grid((Nq + threads - 1) / threads, heads, B);
        RDM_CHECK_CUDA(launch_pdl(attention_kernel<32, 1>, grid, dim3(threads), 0, st, (const float*)q.p, q.ld, (const float*)k.p, k.ld, (const float*)v.p, v.ld, Nq, Nk, scale * 1.4426950408889634f, 0, out));
    }
    LAUNCH_CHECK(); return RDM_OK;
}
int k_attention_d64(View q, View k, View v, int B, int N, int heads, float scale, int causal, Out4 out, cudaStream_t st) {
    RDM_REQUIRE(q.C == heads * 64, RDM_ERR_UNSUPPORTED, "attention_d64: C=%d heads=%d", q.C, heads);
    int threads = N >= 128 ? 128 : ((N + 31) / 32) * 32;
    dim3 grid((N + threads - 1) / threads, heads, B);
    RDM_CHECK_CUDA(launch_pdl(attention_kernel<64, 1>, grid, dim3(threads), 0, st, (const float*)q.p, q.ld, (const float*)k.p, k.ld, (const float*)v.p, v.ld, N, N, scale * 1.4426950408889634f, causal, out));
    LAUNCH_CHECK(); return RDM_OK;
}
int k_softmax_rows(const float* s, int ld_s, int M, int N, float scale, __half* out, int ld_o, cudaStream_t st) {
    RDM_REQUIRE(N % 4 == 0 && N <= 4096 && ld_s % 4 == 0 && ld_o % 4 == 0, RDM_ERR_UNSUPPORTED, "softmax_rows: N=%d (multiple of 4, <= 4096)", N);
    RDM_CHECK_CUDA(launch_pdl(softmax_rows_kernel, dim3(M), dim3(256), 0, st, s, ld_s, N, scale * 1.4426950408889634f, out, ld_o));
    LAUNCH_CHECK(); return RDM_OK;
}
int k_transpose_plane(const __half* in, int ld_in, int rows, int cols, __half* out, int ld_out, cudaStream_t st) {
    RDM_CHECK_CUDA(launch_pdl(transpose_plane_kernel, dim3((cols + 31) / 32, (rows + 31) / 32), dim3(256), 0, st,
                              reinterpret_cast<const unsigned short*>(in), ld_in, rows, cols, reinterpret_cast<unsigned short*>(out), ld_out));
    LAUNCH_CHECK(); return RDM_OK;
}
int k_vq_quantize(const float* z, int B, int E, int HW, const float* codebook, int n_e, const float* pq_w, const float* pq_b, int Z, int quantize,
                  View out, cudaStream_t st) {
    RDM_REQUIRE(E >= 1 && E <= 4 && Z >= 1 && Z <= out.ld, RDM_ERR_UNSUPPORTED, "vq_quantize: embed_dim=%d z_channels=%d", E, Z);
    const long long total = (long long)B * HW;
    vq_quantize_kernel<<<blocks_for(total, 256), 256, 0, st>>>(z, B, E, HW, codebook, n_e, pq_w, pq_b, Z, quantize, out.p, out.ld);
    LAUNCH_CHECK(); return RDM_OK;
}
__global__ void add_vec_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}
// out = a + b (weight-load time: the summed bias of a ResBlock whose skip_connection rides on its second conv, unet.cu)
int k_add_vec(const float* a, const float* b, float* out, int n, cudaStream_t st) {
    add_vec_kernel<<<blocks_for(n, 256), 256, 0, st>>>(a, b, out, n);
    LAUNCH_CHECK(); return RDM_OK;
}
// w [N][9][C] (3x3 taps, row-major) -> out [4][N][4][C]: parity q = (py, px) of the output pixel, 2x2 taps (ty, tx) over the source image.
// Kernel rows that land on source row y + ty + py - 1:  py = 0: ty 0 <- {0}, ty 1 <- {1, 2};   py = 1: ty 0 <- {0, 1}, ty 1 <- {2}  (same for columns).
__global__ void fold_up_weights_kernel(const float* __restrict__ w, int N, int C, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, total = (long long)16 * N * C;
    if (i >= total) return;
    const int c = (int)(i % C); long long t = i / C;
    const int tap = (int)(t % 4); t /= 4;
    const int n = (int)(t % N), q = (int)(t / N);
    const int py = q >> 1, px = q & 1, ty = tap >> 1, tx = tap & 1;
    const int ky0 = py == 0 ? (ty == 0 ? 0 : 1) : (ty == 0 ? 0 : 2), ky1 = py == 0 ? (ty == 0 ? 0 : 2) : (ty == 0 ? 1 : 2);
    const int kx0 = px == 0 ? (tx == 0 ? 0 : 1) : (tx == 0 ? 0 : 2), kx1 = px == 0 ? (tx == 0 ? 0 : 2) : (tx == 0 ? 1 : 2);
    float s = 0.f;
    for (int ky = ky0; ky <= ky1; ky++)
        for (int kx = kx0; kx <= kx1; kx++) s += w[((size_t)n * 9 + ky * 3 + kx) * C + c];
    out[i] = s;
}
int k_fold_up_weights(const float* w, int N, int C, float* out, cudaStream_t st) {
    fold_up_weights_kernel<<<blocks_for((long long)16 * N * C, 256), 256, 0, st>>>(w, N, C, out);
    LAUNCH_CHECK(); return RDM_OK;
}
int k_split_planes(View x, long long M, Out4 y, cudaStream_t st) {
    RDM_REQUIRE(x.C % 4 == 0 && x.ld % 4 == 0, RDM_ERR_ARG, "split_planes: alignment");
    long long total4 = M * (x.C / 4);
    RDM_CHECK_CUDA(launch_pdl(split_planes_kernel, dim3(blocks_for(total4, 256)), dim3(256), 0, st, (const float*)x.p, x.ld, x.C, total4, y));
    LAUNCH_CHECK(); return RDM_OK;
}
int k_im2col_s2(View x, int B, int H, int W, Out4 y, cudaStream_t st) {
    RDM_REQUIRE(x.C % 4 == 0 && x.ld % 4 == 0, RDM_ERR_ARG, "im2col_s2: alignment");
    int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    long long total4 = (long long)B * Ho * Wo * 9 * (x.C / 4);
    RDM_CHECK_CUDA(launch_pdl(im2col_s2_kernel, dim3(blocks_for(total4, 256)), dim3(256), 0, st, (const float*)x.p, x.ld, x.C, B, H, W, Ho, Wo, total4, y));
    LAUNCH_CHECK(); return RDM_OK;
}
int k_upsample2x(View x, int B, int H, int W, Out4 y, cudaStream_t st) {
    RDM_REQUIRE(x.C % 4 == 0 && x.ld % 4 == 0, RDM_ERR_ARG, "upsample2x: alignment");
    long long total4 = (long long)B * 4 * H * W * (x.C / 4);
    RDM_CHECK_CUDA(launch_pdl(upsample2x_kernel, dim3(blocks_for(total4, 256)), dim3(256), 0, st, (const float*)x.p, x.ld, x.C, B, H, W, total4, y));
    LAUNCH_CHECK(); return RDM_OK;
}
int k_silu(const float* in, float* out, long long n, cudaStream_t st) {
    silu_kernel<<<blocks_for(n, 256), 256, 0, st>>>(in, out, n);
    LAUNCH_CHECK(); return RDM_OK;
}
int k_ddim_update(const float* x, const float* eps, long long n, int cfg, float scale, const float* coef, const float* noise,
                  float* x_prev, float* pred_x0, cudaStream_t st) {
    ddim_update_kernel<<<blocks_for(n, 256), 256, 0, st>>>(x, eps, n, cfg, scale, coef, noise, x_prev, pred_x0);
    LAUNCH_CHECK(); return RDM_OK;
}

int k_fill_timesteps(const long long* timesteps, const int* step, int B2, long long* t_out, cudaStream_t st) {
    fill_timesteps_kernel<<<blocks_for(B2, 128), 128, 0, st>>>(timesteps, step, B2, t_out);
    LAUNCH_CHECK(); return RDM_OK;
}
int k_ddim_update_table(const float* x, const float* eps, long long n, int cfg, float scale, const float* coef_table, const int* step,
                        const float* noise_table, float* x_prev, float* pred_x0, cudaStream_t st) {
    ddim_update_table_kernel<<<blocks_for(n, 256), 256, 0, st>>>(x, eps, n, cfg, scale, coef_table, step, noise_table, x_prev, pred_x0);
    LAUNCH_CHECK(); return RDM_OK;
}
int k_step_advance(int* step, cudaStream_t st) {
    step_advance_kernel<<<1, 1, 0, st>>>(step);
    LAUNCH_CHECK(); return RDM_OK;
}
