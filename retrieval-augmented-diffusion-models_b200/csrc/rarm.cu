// RARM decoder with a key/value cache.
//
// Replaces, for sampling, rdm/modules/attention.py RetrievalPatchTransformer (:199-272, configured as in
// models/rarm/imagenet/*/config.yaml:14-27: discrete tokens, positional encodings, 18 x BasicTransformerBlock :77-96 with causal
// self-attention, cross-attention to the k retrieved CLIP vectors and a GEGLU feed-forward, Conv1d(inner, 16384, 1) head) and the loop
// LatentImageRETRO.sample (rdm/models/autoregression/transformer.py:224-270).  The reference re-runs the whole prefix for every new
// token (256 forwards over up to 256 positions, batch doubled under guidance); here every layer keeps its keys/values, so a step is
// ONE new row per sequence: M = B2 <= 8 rows against ~204 M weights.  That is a weight-streaming (HBM-bound) problem, not a tensor-core
// one -- 2*M flop per weight element at M = 8 sits below the fp32 ridge -- so the dense layers are warp-per-column GEMV kernels:
// the <= 8 activation rows live in shared memory (LayerNorm fused into that prologue), every warp streams whole weight rows with
// 8/16-byte coalesced loads (fp16 weights by default: 408 MB per step), fp32 accumulation, transposing shuffle reduction, fused
// bias / residual / GEGLU epilogue.  The step {embed, 18 layers, head, guided top-k sampling, position++} is captured in ONE CUDA
// graph with the position in device memory and replayed 256 times without host synchronisation.
// Algorithmic bytes per step: sum over layers of N*K*sizeof(weight) + the live part of the KV cache (2*L*B2*(pos+1)*C*4 B).
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/rdm_b200.h"
#include <math_constants.h>
#include <string.h>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int GV_ROWS = 8;            // activation rows per CTA (one M chunk)
constexpr int DH = 64;                // d_head of the shipped RARM configs

// ---- warp helpers -------------------------------------------------------------------------------------------
// Transposing reduction: R per-lane partial sums -> lane L holds the complete sum of row row_of_lane<R>(L) (R = 32: row L; R = 16: row L / 2).
template <int R> __device__ __forceinline__ float reduce_rows(float (&acc)[R], int lane) {
    int width = 16;
#pragma unroll
    for (int cnt = R; cnt > 1; cnt >>= 1, width >>= 1) {
        const bool upper = (lane & width) != 0;
#pragma unroll
        for (int i = 0; i < cnt / 2; i++) {
            float send = upper ? acc[i] : acc[i + cnt / 2];
            float keep = upper ? acc[i + cnt / 2] : acc[i];
            acc[i] = keep + __shfl_xor_sync(FULL, send, width);
        }
    }
    float t = acc[0];
    for (; width >= 1; width >>= 1) t += __shfl_xor_sync(FULL, t, width);
    return t;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, m));
    return v;
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

// 4 consecutive weights of one row as floats
__device__ __forceinline__ float4 load_w4(const float* w) { return __ldg(reinterpret_cast<const float4*>(w)); }
__device__ __forceinline__ float4 load_w4(const __half* w) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(w));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

// 8 consecutive weights of one row as floats: ONE 16-byte load for fp16 weights (two for fp32)
__device__ __forceinline__ void load_w8(const float* w, float (&o)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(w)), b = __ldg(reinterpret_cast<const float4*>(w) + 1);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
__device__ __forceinline__ void load_w8(const __half* w, float (&o)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(w));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&u.z)), d = __half22float2(*reinterpret_cast<const __half2*>(&u.w));
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; o[4] = c.x; o[5] = c.y; o[6] = d.x; o[7] = d.y;
}

// The kernels of a decode step are chained with programmatic dependent launch: a kernel may be staged (and request its STATIC weights)
// while its predecessor still runs; it executes griddepcontrol.wait before it touches anything a predecessor wrote and before its own
// first global store.  RDM_PDL=0 switches the attribute off (plain stream order).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_dep(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_rdm_use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- dense layers: out[M, N'] = epi(LN?(x)[M, K] * W[N, K]^T) ------------------------------------------------
struct GemvP {
    const float* x; int ldx;            // activations, fp32 [M, K]
    int M, K, N;                        // N = weight rows
    const float* bias;                  // [N] or null
    const float* ln_g; const float* ln_b;   // LayerNorm over K fused into the prologue (eps 1e-5), or null
    const float* res; int ldres;        // + res[m, n'] or null
    float* out; int ldo;
};
// CPW weight rows (output columns) per warp pass; GEGLU: rows (2j, 2j+1) = (value_j, gate_j) -> output column j (needs CPW == 4).
// MR: activation rows per CTA (4 for batches of <= 4 rows -- an unguided batch of 4 --, else 8 = GV_ROWS): the FMA loop, the staged tile and the
// accumulators scale with it
template <typename WT, int CPW, int WARPS, bool GEGLU, int MR>
__global__ void __launch_bounds__(WARPS * 32, 2) rarm_gemv_kernel(GemvP p, const WT* __restrict__ w) {
    extern __shared__ float sx[];                          // [MR][K]
    const int K = p.K, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m0 = blockIdx.y * MR, rows = min(MR, p.M - m0);
    constexpr int R = CPW * MR, LPR = 32 / R;               // LPR: lanes that hold the same (column, row) sum after the reduction
    const int n0 = (blockIdx.x * WARPS + warp) * CPW;
    pdl_launch_dependents();
    // Weight rows are static: the first PF 256-element blocks of this warp's rows are requested BEFORE the predecessor kernel is waited for
    // and before the activation prologue (staging + LayerNorm), so the DRAM round trip of the weights overlaps both.  A lane owns 8
    // consecutive elements of every block: one 16-byte load per row and block (fp16).
    constexpr int PF = CPW >= 4 ? 2 : 4;                   // 64 registers of weights in flight per lane either way (two CTAs per SM)
    const int nblk = K >> 8;                               // whole 256-element blocks; K % 256 == 128 leaves one 128-element tail
    const WT* wr[CPW];
#pragma unroll
    for (int c = 0; c < CPW; c++) wr[c] = w + (size_t)min(n0 + c, p.N - 1) * K;      // clamped: the duplicate column is not stored
    float wpre[PF][CPW][8];
    if (n0 < p.N) {
#pragma unroll
        for (int b = 0; b < PF; b++)
            if (b < nblk) {
#pragma unroll
                for (int c = 0; c < CPW; c++) load_w8(wr[c] + (b << 8) + lane * 8, wpre[b][c]);
            }
    }
    pdl_wait();
    for (int i = threadIdx.x; i < MR * (K / 4); i += WARPS * 32) {
        const int r = i / (K / 4), c = i % (K / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows) v = *reinterpret_cast<const float4*>(p.x + (size_t)(m0 + r) * p.ldx + c * 4);
        *reinterpret_cast<float4*>(sx + r * K + c * 4) = v;
    }
    __syncthreads();
    if (p.ln_g) {                                          // nn.LayerNorm (attention.py:84-86): two-pass mean / variance in fp32
        for (int r = warp; r < rows; r += WARPS) {
            float* xr = sx + r * K;
            float s = 0.f;
            for (int c = lane; c < K; c += 32) s += xr[c];
            const float mean = warp_sum(s) / (float)K;
            float ss = 0.f;
            for (int c = lane; c < K; c += 32) { const float d = xr[c] - mean; ss += d * d; }
            const float rstd = rsqrtf(warp_sum(ss) / (float)K + 1e-5f);
            for (int c = lane; c < K; c += 32) xr[c] = (xr[c] - mean) * rstd * p.ln_g[c] + p.ln_b[c];
        }
        __syncthreads();
    }
    if (n0 >= p.N) return;                                 // no block-wide barrier below this line
    float acc[R];
#pragma unroll
    for (int i = 0; i < R; i++) acc[i] = 0.f;
    auto fma_block = [&](const float (&wv)[CPW][8], int kc) {
#pragma unroll
        for (int m = 0; m < MR; m++) {
            const float4 a0 = *reinterpret_cast<const float4*>(sx + m * K + kc + lane * 8), a1 = *reinterpret_cast<const float4*>(sx + m * K + kc + lane * 8 + 4);
#pragma unroll
            for (int c = 0; c < CPW; c++) {
                float t = acc[c * MR + m];
                t = fmaf(a0.x, wv[c][0], t); t = fmaf(a0.y, wv[c][1], t); t = fmaf(a0.z, wv[c][2], t); t = fmaf(a0.w, wv[c][3], t);
                t = fmaf(a1.x, wv[c][4], t); t = fmaf(a1.y, wv[c][5], t); t = fmaf(a1.z, wv[c][6], t); t = fmaf(a1.w, wv[c][7], t);
                acc[c * MR + m] = t;
            }
        }
    };
#pragma unroll
    for (int b = 0; b < PF; b++)
        if (b < nblk) fma_block(wpre[b], b << 8);
    // rows longer than PF blocks (the 4C-wide feed-forward input): PF blocks of loads in flight per round
    for (int b0 = PF; b0 < nblk; b0 += PF) {
#pragma unroll
        for (int b = 0; b < PF; b++)
            if (b0 + b < nblk) {
#pragma unroll
                for (int c = 0; c < CPW; c++) load_w8(wr[c] + ((b0 + b) << 8) + lane * 8, wpre[b][c]);
            }
#pragma unroll
        for (int b = 0; b < PF; b++)
            if (b0 + b < nblk) fma_block(wpre[b], (b0 + b) << 8);
    }
    if (K & 128) {                                         // 128-element tail: 4 elements per lane
        const int kc = nblk << 8;
        float4 wv[CPW];
#pragma unroll
        for (int c = 0; c < CPW; c++) wv[c] = load_w4(wr[c] + kc + lane * 4);
#pragma unroll
        for (int m = 0; m < MR; m++) {
            const float4 a = *reinterpret_cast<const float4*>(sx + m * K + kc + lane * 4);
#pragma unroll
            for (int c = 0; c < CPW; c++) {
                float t = acc[c * MR + m];
                t = fmaf(a.x, wv[c].x, t); t = fmaf(a.y, wv[c].y, t); t = fmaf(a.z, wv[c].z, t); t = fmaf(a.w, wv[c].w, t);
                acc[c * MR + m] = t;
            }
        }
    }
    float tot = reduce_rows<R>(acc, lane);
    const int r = lane / LPR;                              // row_of_lane<R>
    const int c = r / MR, m = r % MR, n = n0 + c;
    const bool col_ok = n < p.N;
    if (p.bias && col_ok) tot += p.bias[n];
    if (GEGLU) {
        static_assert(!GEGLU || CPW == 4, "GEGLU pairs need 4 columns per warp");
        const float gate = __shfl_sync(FULL, tot, (lane + MR * LPR) & 31);       // same row m of column c + 1
        if ((c & 1) == 0 && (lane % LPR) == 0 && n + 1 < p.N && m < rows) {
            const int no = n >> 1;
            float t = tot * gelu_erf(gate);
            if (p.res) t += p.res[(size_t)(m0 + m) * p.ldres + no];
            p.out[(size_t)(m0 + m) * p.ldo + no] = t;
        }
    } else {
        const bool writer = (lane % LPR) == 0;
        if (writer && col_ok && m < rows) {
            float t = tot;
            if (p.res) t += p.res[(size_t)(m0 + m) * p.ldres + n];
            p.out[(size_t)(m0 + m) * p.ldo + n] = t;
        }
    }
}

// ---- token embedding + positional encoding (attention.py:255-259): x[b,:] = emb[tok[b % B][pos]] + pos_t[pos] --------
__global__ void rarm_embed_kernel(const long long* __restrict__ tokens, int B, int tcap, const int* __restrict__ pos_dev, const float* __restrict__ emb,
                                  const float* __restrict__ pos_t, int vocab, int C, float* __restrict__ x) {
    const int b = blockIdx.x, pos = *pos_dev;
    long long id = tokens[(size_t)(b % B) * tcap + pos];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    for (int c = threadIdx.x; c < C; c += blockDim.x) x[(size_t)b * C + c] = emb[(size_t)id * C + c] + pos_t[(size_t)pos * C + c];
}

// ---- attention of ONE new query row per (batch, head) over cached keys/values (attention.py:42-74, d_head 64) ---------
// Self-attention (knew != null): the new key/value rows are first appended to the cache at position *pos_dev, then keys 0..pos are
// attended (exactly the rows the reference's causal mask leaves visible to the last query, :58-65).  Cross-attention
// (knew == null): nk_fixed context rows, no mask.  Row j of batch b: kc + (b*tk + j)*ldk + head*64.
__global__ void __launch_bounds__(256) rarm_attn_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ knew, const float* __restrict__ vnew, int ldn,
                                                        float* kc, float* vc, int tk, int ldk, const int* __restrict__ pos_dev, int nk_fixed,
                                                        float scale_log2e, float* __restrict__ out, int ldo) {
    extern __shared__ float sm[];                          // q[64] | p[nk] | red[4*64]
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();
    pdl_wait();                                            // q / knew / vnew come from the preceding GEMV; the cache append below is this kernel's first store
    int nk = nk_fixed;
    if (knew) {
        const int pos = *pos_dev;
        nk = pos + 1;
        if (tid < DH) kc[((size_t)b * tk + pos) * ldk + h * DH + tid] = knew[(size_t)b * ldn + h * DH + tid];
        else if (tid < 2 * DH) vc[((size_t)b * tk + pos) * ldk + h * DH + tid - DH] = vnew[(size_t)b * ldn + h * DH + tid - DH];
    }
    float* sq = sm; float* sp = sm + DH; float* red = sm + DH + ((nk + 3) & ~3);
    __shared__ float s_warp[8];
    if (tid < DH) sq[tid] = q[(size_t)b * ldq + h * DH + tid] * scale_log2e;
    __syncthreads();                                       // orders the cache append (global) and sq (shared) for the whole CTA
    float mx = -CUDART_INF_F;
    for (int j = tid; j < nk; j += 256) {
        const float4* kr = reinterpret_cast<const float4*>(kc + ((size_t)b * tk + j) * ldk + h * DH);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < DH / 4; i++) { const float4 kk = kr[i]; s = fmaf(sq[4 * i], kk.x, s); s = fmaf(sq[4 * i + 1], kk.y, s); s = fmaf(sq[4 * i + 2], kk.z, s); s = fmaf(sq[4 * i + 3], kk.w, s); }
        sp[j] = s; mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    if (lane == 0) s_warp[warp] = mx;
    __syncthreads();
    mx = s_warp[0];
#pragma unroll
    for (int i = 1; i < 8; i++) mx = fmaxf(mx, s_warp[i]);
    __syncthreads();                                       // s_warp is reused for the sum
    float l = 0.f;
    for (int j = tid; j < nk; j += 256) { const float e = exp2f(sp[j] - mx); sp[j] = e; l += e; }
    l = warp_sum(l);
    if (lane == 0) s_warp[warp] = l;
    __syncthreads();
    l = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) l += s_warp[i];
    // out[i] = sum_j p_j v_j[i] / l: 4 groups of 64 threads take every 4th key, then a shared-memory reduction
    const int i = tid & (DH - 1), g = tid >> 6;
    float acc = 0.f;
    for (int j = g; j < nk; j += 4) acc = fmaf(sp[j], vc[((size_t)b * tk + j) * ldk + h * DH + i], acc);
    red[g * DH + i] = acc;
    __syncthreads();
    if (tid < DH) out[(size_t)b * ldo + h * DH + tid] = (red[tid] + red[DH + tid] + red[2 * DH + tid] + red[3 * DH + tid]) / l;
}

// ---- guided top-k sampling of one token per sequence (transformer.py:249-266) ----------------------------------------------
// logits [B2, V] (rows [cond | uncond] when guided).  One CTA per sequence: logit = (lu + s*(lc - lu)) / T in the reference's fp32
// operation order; k-th largest by a 4-pass radix select (taming top_k_logits keeps everything >= the k-th value); softmax with
// float64 sums; draw = first index whose cumulative probability exceeds u * total (u < 0: argmax, lowest index).  The token is
// written to tokens[b, pos + 1] when that position is not part of the given prefix (pos + 1 >= n_fixed).
__device__ __forceinline__ unsigned key_of(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
constexpr int SAMP_THREADS = 1024;
__global__ void __launch_bounds__(SAMP_THREADS) rarm_sample_kernel(const float* __restrict__ logits, int B, int V, int guided, float scale, float temperature, int top_k,
                                                                   const float* __restrict__ uniforms, int greedy, const int* __restrict__ pos_dev, int n_fixed, int tcap,
                                                                   long long* __restrict__ tokens, float* __restrict__ probs_out) {
    extern __shared__ float sl[];                          // [V]
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_remaining;
    __shared__ float s_red[32];
    __shared__ double s_dsum[32];
    __shared__ unsigned long long s_best;
    __shared__ int s_idx, s_last;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pos = *pos_dev;
    const float* lc = logits + (size_t)b * V;
    const float* lu = logits + (size_t)(b + B) * V;
    float mx = -CUDART_INF_F;
    for (int i = tid; i < V; i += SAMP_THREADS) {
        float e = lc[i];
        if (guided) { const float u = lu[i]; e = __fadd_rn(u, __fmul_rn(scale, __fsub_rn(e, u))); }
        e = __fdiv_rn(e, temperature) + 0.f;               // + 0: -0 -> +0 so the integer key order equals the float order
        sl[i] = e; mx = fmaxf(mx, e);
    }
    mx = warp_max(mx);
    if (lane == 0) s_red[warp] = mx;
    if (tid == 0) { s_prefix = 0u; s_remaining = (unsigned)top_k; s_best = 0ull; s_idx = -1; s_last = -1; }
    __syncthreads();
    mx = s_red[0];
    for (int i = 1; i < SAMP_THREADS / 32; i++) mx = fmaxf(mx, s_red[i]);
    unsigned thr_key = 0u;                                 // keep everything
    if (top_k > 0 && top_k < V) {
        unsigned mask = 0u;
        for (int shift = 24; shift >= 0; shift -= 8) {
            if (tid < 256) hist[tid] = 0u;
            __syncthreads();
            const unsigned prefix = s_prefix;
            for (int i = tid; i < V; i += SAMP_THREADS) { const unsigned k = key_of(sl[i]); if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u); }
            __syncthreads();
            if (tid == 0) {
                unsigned cum = 0u, rem = s_remaining; int bin = 255;
                for (; bin > 0; bin--) { if (cum + hist[bin] >= rem) break; cum += hist[bin]; }
                s_remaining = rem - cum; s_prefix = prefix | ((unsigned)bin << shift);
            }
            mask |= 255u << shift;
            __syncthreads();
        }
        thr_key = s_prefix;
    }
    // probabilities of the kept entries; thread t owns the contiguous range [t*E, (t+1)*E) so that a block scan gives the CDF in index order
    const int E = (V + SAMP_THREADS - 1) / SAMP_THREADS, i0 = tid * E, i1 = min(V, i0 + E);
    double local = 0.0; int last_kept = -1;
    for (int i = i0; i < i1; i++) {
        const float v = sl[i];
        float pr = 0.f;
        if (key_of(v) >= thr_key) { pr = expf(v - mx); last_kept = i; }
        sl[i] = pr; local += (double)pr;
    }
    if (greedy) {
        unsigned long long best = 0ull;
        for (int i = i0; i < i1; i++) { const unsigned long long k = ((unsigned long long)__float_as_uint(sl[i]) << 32) | (unsigned)(0xffffffffu - (unsigned)i); best = k > best ? k : best; }
        atomicMax(&s_best, best);                          // probabilities are >= 0: their bit patterns order like the values; ties -> lowest index
    }
    if (last_kept >= 0) atomicMax(&s_last, last_kept);
    // inclusive scan of the per-thread sums
    double incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const double o = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) s_dsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        double w = s_dsum[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const double o = __shfl_up_sync(FULL, w, d); if (lane >= d) w += o; }
        s_dsum[lane] = w;                                  // inclusive over warps
    }
    __syncthreads();
    const double total = s_dsum[SAMP_THREADS / 32 - 1];
    const double excl = (warp ? s_dsum[warp - 1] : 0.0) + incl - local;
    if (!greedy) {
        const double target = (double)uniforms[(size_t)max(pos + 1 - n_fixed, 0) * B + b] * total;
        if (local > 0.0 && excl <= target && target < excl + local) {
            double cum = excl; int pick = last_kept;
            for (int i = i0; i < i1; i++) { cum += (double)sl[i]; if (cum > target) { pick = i; break; } }
            s_idx = pick;
        }
    }
    __syncthreads();
    if (probs_out) { const float inv = (float)(1.0 / total); for (int i = tid; i < V; i += SAMP_THREADS) probs_out[(size_t)b * V + i] = sl[i] * inv; }
    if (tid == 0) {
        int idx = greedy ? (int)(0xffffffffu - (unsigned)(s_best & 0xffffffffull)) : (s_idx >= 0 ? s_idx : s_last);
        if (pos + 1 >= n_fixed && pos + 1 < tcap) tokens[(size_t)b * tcap + pos + 1] = idx;
    }
}

__global__ void rarm_advance_kernel(int* pos) { if (threadIdx.x == 0 && blockIdx.x == 0) *pos += 1; }

// fp32 -> fp16 weight plane (round to nearest, clamped to the finite range)
__global__ void rarm_to_half_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2half_rn(fminf(fmaxf(in[i], -65504.f), 65504.f));
}

// ---- host side ------------------------------------------------------------------------------------------------
struct Slot {
    size_t numel = 0; float* dst = nullptr; bool loaded = false;
    int kind = 0;              // 0 plain; 1 transpose [rows, cols] -> [cols, rows]; 2 rows scattered (dst_row0, step); 3 GEGLU interleave (half_rows, row_len)
    int rows = 0, row_len = 0, dst_row0 = 0, dst_step = 1;
};
struct LayerW {
    float *ln1g, *ln1b, *qkv, *o1w, *o1b, *ln2g, *ln2b, *q2, *kv2, *o2w, *o2b, *ln3g, *ln3b, *ff1w, *ff1b, *ff2w, *ff2b;
};

}  // namespace

struct rdm_rarm {
    int device = 0; rdm_rarm_cfg cfg{}; int mode = RDM_UNET_MODE_TC_FP16;
    int C = 0;
    float* wbase = nullptr; size_t wfloats = 0, woff = 0;
    __half* wh = nullptr; bool half_dirty = true;
    std::unordered_map<std::string, Slot> params; std::vector<std::string> order;
    std::vector<LayerW> layers;
    float *emb = nullptr, *pos_t = nullptr, *headw = nullptr, *headb = nullptr;
    // state of the current context / sequences
    int B2 = 0, kctx = 0;
    float *ctx_kv = nullptr;      // [L][B2*kctx][2C]
    float *kcache = nullptr, *vcache = nullptr;   // [L][B2][T][C]
    float *x0 = nullptr, *x1 = nullptr, *qkv = nullptr, *att = nullptr, *qx = nullptr, *ffh = nullptr, *logits = nullptr;
    int cap_B2 = 0, cap_k = 0;
    int* pos_dev = nullptr;
    long long* tok_buf = nullptr; float* u_buf = nullptr; int tok_cap_B = 0;
    cudaStream_t cap_stream = nullptr;
    cudaGraphExec_t step_exec = nullptr; unsigned long long step_kernels = 0; long long step_key[10] = {0};
    bool use_graph = true;
};

namespace {
typedef rdm_rarm Rarm;
#define LAUNCH_CHECK() do { RDM_COUNT_LAUNCH(); RDM_CHECK_CUDA(cudaGetLastError()); } while (0)

float* walloc(Rarm* n, size_t floats) { floats = (floats + 63) & ~(size_t)63; float* p = n->wbase ? n->wbase + n->woff : (float*)nullptr + n->woff; n->woff += floats; return p; }
float* reg(Rarm* n, const std::string& name, float* dst, size_t numel, int kind = 0, int rows = 0, int row_len = 0, int dst_row0 = 0, int dst_step = 1) {
    Slot s; s.numel = numel; s.dst = dst; s.kind = kind; s.rows = rows; s.row_len = row_len; s.dst_row0 = dst_row0; s.dst_step = dst_step;
    n->params[name] = s; n->order.push_back(name);
    return dst;
}
// names in the reference's registration order (attention.py:224-245; BasicTransformerBlock :79-87: attn1, ff, attn2, norm1..3)
void build(Rarm* n) {
    const rdm_rarm_cfg& c = n->cfg; const int C = c.n_heads * c.d_head, X = c.context_dim, T = c.sequence_length;
    n->C = C; n->woff = 0; n->params.clear(); n->order.clear(); n->layers.clear();
    n->pos_t = walloc(n, (size_t)T * C); reg(n, "positional_encoding", n->pos_t, (size_t)C * T, 1, C, T);            // [C, T] -> [T, C]
    n->emb = walloc(n, (size_t)c.in_channels * C); reg(n, "proj_in.weight", n->emb, (size_t)c.in_channels * C);
    for (int i = 0; i < c.depth; i++) {
        const std::string p = "transformer_blocks." + std::to_string(i) + ".";
        LayerW l{};
        l.qkv = walloc(n, (size_t)3 * C * C);
        reg(n, p + "attn1.to_q.weight", l.qkv, (size_t)C * C, 2, C, C, 0, 1);
        reg(n, p + "attn1.to_k.weight", l.qkv, (size_t)C * C, 2, C, C, C, 1);
        reg(n, p + "attn1.to_v.weight", l.qkv, (size_t)C * C, 2, C, C, 2 * C, 1);
        l.o1w = walloc(n, (size_t)C * C); reg(n, p + "attn1.to_out.0.weight", l.o1w, (size_t)C * C);
        l.o1b = walloc(n, C); reg(n, p + "attn1.to_out.0.bias", l.o1b, C);
        l.ff1w = walloc(n, (size_t)8 * C * C); reg(n, p + "ff.net.0.proj.weight", l.ff1w, (size_t)8 * C * C, 3, 4 * C, C);
        l.ff1b = walloc(n, (size_t)8 * C); reg(n, p + "ff.net.0.proj.bias", l.ff1b, (size_t)8 * C, 3, 4 * C, 1);
        l.ff2w = walloc(n, (size_t)4 * C * C); reg(n, p + "ff.net.2.weight", l.ff2w, (size_t)4 * C * C);
        l.ff2b = walloc(n, C); reg(n, p + "ff.net.2.bias", l.ff2b, C);
        l.q2 = walloc(n, (size_t)C * C); reg(n, p + "attn2.to_q.weight", l.q2, (size_t)C * C);
        l.kv2 = walloc(n, (size_t)2 * C * X);
        reg(n, p + "attn2.to_k.weight", l.kv2, (size_t)C * X, 2, C, X, 0, 1);
        reg(n, p + "attn2.to_v.weight", l.kv2, (size_t)C * X, 2, C, X, C, 1);
        l.o2w = walloc(n, (size_t)C * C); reg(n, p + "attn2.to_out.0.weight", l.o2w, (size_t)C * C);
        l.o2b = walloc(n, C); reg(n, p + "attn2.to_out.0.bias", l.o2b, C);
        l.ln1g = walloc(n, C); reg(n, p + "norm1.weight", l.ln1g, C); l.ln1b = walloc(n, C); reg(n, p + "norm1.bias", l.ln1b, C);
        l.ln2g = walloc(n, C); reg(n, p + "norm2.weight", l.ln2g, C); l.ln2b = walloc(n, C); reg(n, p + "norm2.bias", l.ln2b, C);
        l.ln3g = walloc(n, C); reg(n, p + "norm3.weight", l.ln3g, C); l.ln3b = walloc(n, C); reg(n, p + "norm3.bias", l.ln3b, C);
        n->layers.push_back(l);
    }
    n->headw = walloc(n, (size_t)c.out_channels * C); reg(n, "proj_out.weight", n->headw, (size_t)c.out_channels * C);      // Conv1d [out, C, 1]
    n->headb = walloc(n, c.out_channels); reg(n, "proj_out.bias", n->headb, c.out_channels);
}
int64_t missing(const Rarm* n) { int64_t m = 0; for (auto& kv : n->params) if (!kv.second.loaded) m++; return m; }

int ensure_half(Rarm* n, cudaStream_t st) {
    if (n->mode == RDM_UNET_MODE_FP32 || !n->half_dirty) return RDM_OK;
    if (!n->wh) RDM_CHECK_CUDA(cudaMalloc((void**)&n->wh, n->wfloats * sizeof(__half)));
    rarm_to_half_kernel<<<(unsigned)((n->wfloats + 255) / 256), 256, 0, st>>>(n->wbase, n->wh, n->wfloats);
    LAUNCH_CHECK();
    n->half_dirty = false;
    return RDM_OK;
}

template <typename WT, int CPW, int WARPS, bool GEGLU, int MR>
int launch_gemv_mr(const GemvP& p, const WT* w, cudaStream_t st) {
    const size_t smem = (size_t)MR * p.K * sizeof(float);
    static size_t configured = 0;                          // per instantiation; all handles of the process share the device function
    if (smem > 48 * 1024 && smem > configured) {
        RDM_CHECK_CUDA(cudaFuncSetAttribute(rarm_gemv_kernel<WT, CPW, WARPS, GEGLU, MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = 200 * 1024;
    }
    dim3 grid((unsigned)ceil_div(p.N, CPW * WARPS), (unsigned)ceil_div(p.M, MR));
    RDM_CHECK_CUDA(launch_dep(rarm_gemv_kernel<WT, CPW, WARPS, GEGLU, MR>, grid, dim3(WARPS * 32), smem, st, p, w));
    LAUNCH_CHECK();
    return RDM_OK;
}
template <typename WT, int CPW, int WARPS, bool GEGLU>
int launch_gemv(const GemvP& p, const WT* w, cudaStream_t st) {
    return p.M <= 4 ? launch_gemv_mr<WT, CPW, WARPS, GEGLU, 4>(p, w, st) : launch_gemv_mr<WT, CPW, WARPS, GEGLU, GV_ROWS>(p, w, st);
}
// out[M, N or N/2] = epi(LN?(x) * W^T): W is an fp32 pointer into the weight arena; the fp16 modes read the same offset of the half plane
int gemv(Rarm* n, const float* x, int ldx, int M, int K, const float* w, int N, const float* bias, const float* ln_g, const float* ln_b,
         const float* res, int ldres, bool geglu, float* out, int ldo, cudaStream_t st) {
    RDM_REQUIRE(K % 128 == 0 && (size_t)GV_ROWS * K * 4 <= 200 * 1024 && ldx % 4 == 0, RDM_ERR_UNSUPPORTED, "rarm gemv: K = %d (multiple of 128, <= 6400)", K);
    RDM_REQUIRE(!geglu || N % 4 == 0, RDM_ERR_UNSUPPORTED, "rarm gemv: GEGLU needs N %% 4 == 0 (N = %d)", N);
    GemvP p{x, ldx, M, K, N, bias, ln_g, ln_b, res, ldres, out, ldo};
    const bool f16 = n->mode != RDM_UNET_MODE_FP32;
    const __half* wh = f16 ? n->wh + (w - n->wbase) : nullptr;
    if (geglu) return f16 ? launch_gemv<__half, 4, 8, true>(p, wh, st) : launch_gemv<float, 4, 8, true>(p, w, st);
    if (N <= 1024) return f16 ? launch_gemv<__half, 2, 4, false>(p, wh, st) : launch_gemv<float, 2, 4, false>(p, w, st);
    return f16 ? launch_gemv<__half, 4, 8, false>(p, wh, st) : launch_gemv<float, 4, 8, false>(p, w, st);
}

int attn(const float* q, int ldq, const float* knew, const float* vnew, int ldn, float* kc, float* vc, int tk, int ldk, const int* pos_dev, int nk_fixed,
         int nk_max, int heads, int B2, float* out, int ldo, cudaStream_t st) {
    const size_t smem = (size_t)(DH + ((nk_max + 3) & ~3) + 4 * DH) * sizeof(float);
    RDM_REQUIRE(smem <= 48 * 1024, RDM_ERR_UNSUPPORTED, "rarm attention: %d keys", nk_max);
    RDM_CHECK_CUDA(launch_dep(rarm_attn_kernel, dim3(heads, B2), dim3(256), smem, st, q, ldq, knew, vnew, ldn, kc, vc, tk, ldk, pos_dev, nk_fixed, 0.125f * 1.4426950408889634f, out, ldo));
    LAUNCH_CHECK();
    return RDM_OK;
}

int ensure_state(Rarm* n, int B2, int k) {
    const int C = n->C, L = n->cfg.depth, T = n->cfg.sequence_length, V = n->cfg.out_channels;
    if (B2 <= n->cap_B2 && k <= n->cap_k) return RDM_OK;
    const int cb = B2 > n->cap_B2 ? B2 : n->cap_B2, ck = k > n->cap_k ? k : n->cap_k;
    for (void* p : {(void*)n->ctx_kv, (void*)n->kcache, (void*)n->vcache, (void*)n->x0, (void*)n->x1, (void*)n->qkv, (void*)n->att, (void*)n->qx, (void*)n->ffh, (void*)n->logits})
        if (p) cudaFree(p);
    n->ctx_kv = n->kcache = n->vcache = n->x0 = n->x1 = n->qkv = n->att = n->qx = n->ffh = n->logits = nullptr; n->cap_B2 = n->cap_k = 0;
    if (n->step_exec) { cudaGraphExecDestroy(n->step_exec); n->step_exec = nullptr; }       // the captured pointers are gone
    RDM_CHECK_CUDA(cudaMalloc((void**)&n->ctx_kv, (size_t)L * cb * ck * 2 * C * 4));
    RDM_CHECK_CUDA(cudaMalloc((void**)&n->kcache, (size_t)L * cb * T * C * 4));
    RDM_CHECK_CUDA(cudaMalloc((void**)&n->vcache, (size_t)L * cb * T * C * 4));
    RDM_CHECK_CUDA(cudaMalloc((void**)&n->x0, (size_t)cb * C * 4));
    RDM_CHECK_CUDA(cudaMalloc((void**)&n->x1, (size_t)cb * C * 4));
    RDM_CHECK_CUDA(cudaMalloc((void**)&n->qkv, (size_t)cb * 3 * C * 4));
    RDM_CHECK_CUDA(cudaMalloc((void**)&n->att, (size_t)cb * C * 4));
    RDM_CHECK_CUDA(cudaMalloc((void**)&n->qx, (size_t)cb * C * 4));
    RDM_CHECK_CUDA(cudaMalloc((void**)&n->ffh, (size_t)cb * 4 * C * 4));
    RDM_CHECK_CUDA(cudaMalloc((void**)&n->logits, (size_t)cb * V * 4));
    n->cap_B2 = cb; n->cap_k = ck;
    return RDM_OK;
}

// token / uniform staging owned by the handle (stable addresses for the captured graph): tok_buf int64 [rows, T+1], u_buf float [T+1, rows]
int ensure_tok(Rarm* n, int rows) {
    if (n->tok_cap_B >= rows) return RDM_OK;
    const int T = n->cfg.sequence_length;
    if (n->tok_buf) cudaFree(n->tok_buf);
    if (n->u_buf) cudaFree(n->u_buf);
    n->tok_buf = nullptr; n->u_buf = nullptr; n->tok_cap_B = 0;
    if (n->step_exec) { cudaGraphExecDestroy(n->step_exec); n->step_exec = nullptr; }
    RDM_CHECK_CUDA(cudaMalloc((void**)&n->tok_buf, (size_t)rows * (T + 1) * sizeof(long long)));
    RDM_CHECK_CUDA(cudaMalloc((void**)&n->u_buf, (size_t)rows * (T + 1) * sizeof(float)));
    n->tok_cap_B = rows;
    return RDM_OK;
}

// One position of every sequence: x = embed(tokens[:, pos]) -> layers -> logits [B2, V]; the position is read from n->pos_dev.
// `tokens` int64 [B, tcap] (device); rows b >= B reuse the tokens of row b - B (guidance doubling, transformer.py:247-248).
int step_forward(Rarm* n, const long long* tokens, int B, int tcap, cudaStream_t st) {
    const rdm_rarm_cfg& c = n->cfg; const int C = n->C, B2 = n->B2, T = c.sequence_length, H = c.n_heads;
    rarm_embed_kernel<<<B2, 256, 0, st>>>(tokens, B, tcap, n->pos_dev, n->emb, n->pos_t, c.in_channels, C, n->x0);
    LAUNCH_CHECK();
    float* x = n->x0; float* y = n->x1;
    for (int i = 0; i < c.depth; i++) {
        const LayerW& l = n->layers[i];
        float* kc = n->kcache + (size_t)i * n->cap_B2 * T * C; float* vc = n->vcache + (size_t)i * n->cap_B2 * T * C;
        float* ckv = n->ctx_kv + (size_t)i * n->cap_B2 * n->cap_k * 2 * C;
        // x = attn1(norm1(x)) + x   (attention.py:93)
        RDM_TRY(gemv(n, x, C, B2, C, l.qkv, 3 * C, nullptr, l.ln1g, l.ln1b, nullptr, 0, false, n->qkv, 3 * C, st));
        RDM_TRY(attn(n->qkv, 3 * C, n->qkv + C, n->qkv + 2 * C, 3 * C, kc, vc, T, C, n->pos_dev, 0, T, H, B2, n->att, C, st));
        RDM_TRY(gemv(n, n->att, C, B2, C, l.o1w, C, l.o1b, nullptr, nullptr, x, C, false, y, C, st));
        std::swap(x, y);
        // x = attn2(norm2(x), context) + x   (:94)
        RDM_TRY(gemv(n, x, C, B2, C, l.q2, C, nullptr, l.ln2g, l.ln2b, nullptr, 0, false, n->qx, C, st));
        RDM_TRY(attn(n->qx, C, nullptr, nullptr, 0, ckv, ckv + C, n->kctx, 2 * C, n->pos_dev, n->kctx, n->kctx, H, B2, n->att, C, st));
        RDM_TRY(gemv(n, n->att, C, B2, C, l.o2w, C, l.o2b, nullptr, nullptr, x, C, false, y, C, st));
        std::swap(x, y);
        // x = ff(norm3(x)) + x   (:95; GEGLU)
        RDM_TRY(gemv(n, x, C, B2, C, l.ff1w, 8 * C, l.ff1b, l.ln3g, l.ln3b, nullptr, 0, true, n->ffh, 4 * C, st));
        RDM_TRY(gemv(n, n->ffh, 4 * C, B2, 4 * C, l.ff2w, C, l.ff2b, nullptr, nullptr, x, C, false, y, C, st));
        std::swap(x, y);
    }
    RDM_TRY(gemv(n, x, C, B2, C, n->headw, c.out_channels, n->headb, nullptr, nullptr, nullptr, 0, false, n->logits, c.out_channels, st));
    return RDM_OK;
}

int sample_launch(Rarm* n, const float* logits, int B, int V, int guided, float scale, float temperature, int top_k, const float* uniforms, int greedy,
                  int n_fixed, int tcap, long long* tokens, float* probs_out, cudaStream_t st) {
    const size_t smem = (size_t)V * sizeof(float);
    RDM_REQUIRE(V >= 1 && smem <= 160 * 1024, RDM_ERR_UNSUPPORTED, "rarm sampler: vocabulary %d", V);
    static bool configured = false;
    if (smem > 48 * 1024 && !configured) {
        RDM_CHECK_CUDA(cudaFuncSetAttribute(rarm_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        configured = true;
    }
    rarm_sample_kernel<<<B, SAMP_THREADS, smem, st>>>(logits, B, V, guided, scale, temperature, top_k, uniforms, greedy, n->pos_dev, n_fixed, tcap, tokens, probs_out);
    LAUNCH_CHECK();
    return RDM_OK;
}

template <typename F> int capture_graph(Rarm* n, F body) {
    if (n->step_exec) { cudaGraphExecDestroy(n->step_exec); n->step_exec = nullptr; }
    RDM_CHECK_CUDA(cudaStreamBeginCapture(n->cap_stream, cudaStreamCaptureModeThreadLocal));
    const unsigned long long before = g_rdm_launches;
    int rc = body(n->cap_stream);
    n->step_kernels = g_rdm_launches - before; g_rdm_launches = before;
    cudaGraph_t g = nullptr;
    cudaError_t ce = cudaStreamEndCapture(n->cap_stream, &g);
    if (rc != RDM_OK) { if (g) cudaGraphDestroy(g); return rc; }
    RDM_REQUIRE(ce == cudaSuccess && g, RDM_ERR_CUDA, "rarm: graph capture failed: %s", cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(&n->step_exec, g, 0);
    cudaGraphDestroy(g);
    RDM_REQUIRE(ce == cudaSuccess, RDM_ERR_CUDA, "rarm: cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
    return RDM_OK;
}

}  // namespace

extern "C" {

int rdm_rarm_create(rdm_rarm_t** out, const rdm_rarm_cfg* c, int32_t device) {
    RDM_REQUIRE(out && c, RDM_ERR_ARG, "rdm_rarm_create: null argument");
    RDM_REQUIRE(c->d_head == DH, RDM_ERR_UNSUPPORTED, "rdm_rarm_create: d_head %d (the shipped RARM configs use 64)", c->d_head);
    RDM_REQUIRE(c->n_heads >= 1 && c->depth >= 1 && c->in_channels >= 1 && c->out_channels >= 1 && c->sequence_length >= 1, RDM_ERR_ARG, "rdm_rarm_create: bad sizes");
    RDM_REQUIRE((c->n_heads * c->d_head) % 128 == 0 && c->context_dim % 128 == 0, RDM_ERR_UNSUPPORTED,
                "rdm_rarm_create: inner width %d and context_dim %d must be multiples of 128", c->n_heads * c->d_head, c->context_dim);
    RDM_REQUIRE(c->sequence_length <= 4096 && (size_t)c->out_channels * 4 <= 160 * 1024, RDM_ERR_UNSUPPORTED, "rdm_rarm_create: sequence_length %d / out_channels %d too large",
                c->sequence_length, c->out_channels);
    DeviceGuard guard(device);
    RDM_REQUIRE(guard.ok, RDM_ERR_CUDA, "rdm_rarm_create: cannot select device %d", device);
    rdm_rarm* n = new rdm_rarm(); n->device = device; n->cfg = *c;
    build(n); n->wfloats = n->woff;
    if (cudaMalloc((void**)&n->wbase, n->wfloats * 4) != cudaSuccess || cudaMalloc((void**)&n->pos_dev, sizeof(int)) != cudaSuccess ||
        cudaStreamCreateWithFlags(&n->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
        rdm_set_error("rdm_rarm_create: allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        rdm_rarm_destroy(n); return RDM_ERR_CUDA;
    }
    cudaMemset(n->wbase, 0, n->wfloats * 4);
    cudaMemset(n->pos_dev, 0, sizeof(int));
    build(n);
    *out = n; return RDM_OK;
}
void rdm_rarm_destroy(rdm_rarm_t* n) {
    if (!n) return;
    DeviceGuard guard(n->device);
    if (n->step_exec) cudaGraphExecDestroy(n->step_exec);
    if (n->cap_stream) cudaStreamDestroy(n->cap_stream);
    for (void* p : {(void*)n->wbase, (void*)n->wh, (void*)n->ctx_kv, (void*)n->kcache, (void*)n->vcache, (void*)n->x0, (void*)n->x1, (void*)n->qkv, (void*)n->att,
                    (void*)n->qx, (void*)n->ffh, (void*)n->logits, (void*)n->pos_dev, (void*)n->tok_buf, (void*)n->u_buf})
        if (p) cudaFree(p);
    delete n;
}
int64_t rdm_rarm_num_params(const rdm_rarm_t* n) { return n ? (int64_t)n->order.size() : 0; }
const char* rdm_rarm_param_name(const rdm_rarm_t* n, int64_t i) { return (n && i >= 0 && i < (int64_t)n->order.size()) ? n->order[i].c_str() : nullptr; }
int64_t rdm_rarm_param_numel(const rdm_rarm_t* n, const char* name) { if (!n || !name) return -1; auto it = n->params.find(name); return it == n->params.end() ? -1 : (int64_t)it->second.numel; }
int64_t rdm_rarm_missing(const rdm_rarm_t* n) { return n ? missing(n) : -1; }
int rdm_rarm_set_mode(rdm_rarm_t* n, int32_t mode) {
    RDM_REQUIRE(n && (mode == RDM_UNET_MODE_FP32 || mode == RDM_UNET_MODE_TC_FP16), RDM_ERR_ARG,
                "rdm_rarm_set_mode: RDM_UNET_MODE_FP32 (fp32 weights) or RDM_UNET_MODE_TC_FP16 (fp16 weights, fp32 accumulation)");
    if (mode != n->mode && n->step_exec) { cudaGraphExecDestroy(n->step_exec); n->step_exec = nullptr; }
    n->mode = mode; return RDM_OK;
}
int rdm_rarm_set_graph(rdm_rarm_t* n, int32_t on) { RDM_REQUIRE(n, RDM_ERR_ARG, "rdm_rarm_set_graph: null handle"); n->use_graph = on != 0; return RDM_OK; }

int rdm_rarm_load(rdm_rarm_t* n, const char* name, const float* host, int64_t numel) {
    RDM_REQUIRE(n && name && host, RDM_ERR_ARG, "rdm_rarm_load: null argument");
    auto it = n->params.find(name);
    RDM_REQUIRE(it != n->params.end(), RDM_ERR_ARG, "rdm_rarm_load: unknown parameter '%s'", name);
    Slot& s = it->second;
    RDM_REQUIRE((size_t)numel == s.numel, RDM_ERR_ARG, "rdm_rarm_load: '%s' has %lld elements, expected %zu", name, (long long)numel, s.numel);
    DeviceGuard guard(n->device);
    if (s.kind == 0) {
        RDM_CHECK_CUDA(cudaMemcpy(s.dst, host, (size_t)numel * 4, cudaMemcpyHostToDevice));
    } else if (s.kind == 1) {                       // [rows, cols] -> [cols, rows]
        std::vector<float> t((size_t)numel);
        for (int r = 0; r < s.rows; r++) for (int c = 0; c < s.row_len; c++) t[(size_t)c * s.rows + r] = host[(size_t)r * s.row_len + c];
        RDM_CHECK_CUDA(cudaMemcpy(s.dst, t.data(), (size_t)numel * 4, cudaMemcpyHostToDevice));
    } else if (s.kind == 2) {                       // rows r -> destination rows dst_row0 + r*dst_step (concatenated projections)
        RDM_CHECK_CUDA(cudaMemcpy2D(s.dst + (size_t)s.dst_row0 * s.row_len, (size_t)s.dst_step * s.row_len * 4, host, (size_t)s.row_len * 4, (size_t)s.row_len * 4, s.rows,
                                    cudaMemcpyHostToDevice));
    } else {                                        // GEGLU: value rows j -> 2j, gate rows half + j -> 2j + 1
        std::vector<float> t((size_t)numel);
        for (int j = 0; j < s.rows; j++) {
            memcpy(&t[(size_t)(2 * j) * s.row_len], &host[(size_t)j * s.row_len], (size_t)s.row_len * 4);
            memcpy(&t[(size_t)(2 * j + 1) * s.row_len], &host[(size_t)(s.rows + j) * s.row_len], (size_t)s.row_len * 4);
        }
        RDM_CHECK_CUDA(cudaMemcpy(s.dst, t.data(), (size_t)numel * 4, cudaMemcpyHostToDevice));
    }
    s.loaded = true; n->half_dirty = true;
    return RDM_OK;
}

int rdm_rarm_set_context(rdm_rarm_t* n, const float* ctx_dev, int32_t B2, int32_t k, void* stream) {
    RDM_REQUIRE(n && ctx_dev && B2 >= 1 && k >= 1, RDM_ERR_ARG, "rdm_rarm_set_context: bad argument");
    RDM_REQUIRE(k <= 1024, RDM_ERR_UNSUPPORTED, "rdm_rarm_set_context: %d context rows", k);
    RDM_REQUIRE(missing(n) == 0, RDM_ERR_STATE, "rdm_rarm_set_context: %lld parameters not loaded", (long long)missing(n));
    DeviceGuard guard(n->device);
    cudaStream_t st = (cudaStream_t)stream;
    RDM_TRY(ensure_state(n, B2, k));
    RDM_TRY(ensure_half(n, st));
    n->B2 = B2; n->kctx = k;
    const int C = n->C, X = n->cfg.context_dim;
    for (int i = 0; i < n->cfg.depth; i++)          // k = to_k(context), v = to_v(context) (attention.py:47-48): step-invariant, projected once
        RDM_TRY(gemv(n, ctx_dev, X, B2 * k, X, n->layers[i].kv2, 2 * C, nullptr, nullptr, nullptr, nullptr, 0, false,
                     n->ctx_kv + (size_t)i * n->cap_B2 * n->cap_k * 2 * C, 2 * C, st));
    return RDM_OK;
}

int rdm_rarm_forward_token(rdm_rarm_t* n, const int64_t* tokens_dev, int32_t B, int32_t pos, float* logits_out_dev, void* stream) {
    RDM_REQUIRE(n && tokens_dev && logits_out_dev, RDM_ERR_ARG, "rdm_rarm_forward_token: null argument");
    RDM_REQUIRE(n->B2 > 0 && (B == n->B2 || 2 * B == n->B2), RDM_ERR_STATE, "rdm_rarm_forward_token: context holds %d rows, tokens %d", n->B2, B);
    RDM_REQUIRE(pos >= 0 && pos < n->cfg.sequence_length, RDM_ERR_ARG, "rdm_rarm_forward_token: position %d outside [0, %d)", pos, n->cfg.sequence_length);
    DeviceGuard guard(n->device);
    cudaStream_t st = (cudaStream_t)stream;
    RDM_TRY(ensure_half(n, st));
    RDM_CHECK_CUDA(cudaMemcpyAsync(n->pos_dev, &pos, sizeof(int), cudaMemcpyHostToDevice, st));
    // the ids are staged in column `pos` of the handle's token buffer (the position also selects the positional encoding and the cache row)
    const int T = n->cfg.sequence_length;
    RDM_TRY(ensure_tok(n, n->B2));
    RDM_CHECK_CUDA(cudaMemcpy2DAsync(n->tok_buf + pos, (size_t)(T + 1) * sizeof(long long), tokens_dev, sizeof(long long), sizeof(long long), B, cudaMemcpyDeviceToDevice, st));
    RDM_TRY(step_forward(n, n->tok_buf, B, T + 1, st));
    RDM_CHECK_CUDA(cudaMemcpyAsync(logits_out_dev, n->logits, (size_t)n->B2 * n->cfg.out_channels * 4, cudaMemcpyDeviceToDevice, st));
    return RDM_OK;
}

int rdm_rarm_sample_step(rdm_rarm_t* n, const float* logits_dev, int32_t B, int32_t V, int32_t guided, float guidance_scale, float temperature, int32_t top_k,
                         const float* uniforms_dev, int64_t* token_out_dev, float* probs_out_dev, void* stream) {
    RDM_REQUIRE(n && logits_dev && token_out_dev && B >= 1 && temperature > 0.f, RDM_ERR_ARG, "rdm_rarm_sample_step: bad argument");
    DeviceGuard guard(n->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int zero = 0;
    RDM_CHECK_CUDA(cudaMemcpyAsync(n->pos_dev, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
    // pos = 0, n_fixed = 1, tcap = 2: the token lands in column 1 of a [B, 2] view of the staging buffer
    RDM_TRY(ensure_tok(n, B));
    RDM_TRY(sample_launch(n, logits_dev, B, V, guided, guidance_scale, temperature, top_k, uniforms_dev, uniforms_dev ? 0 : 1, 1, 2, n->tok_buf, probs_out_dev, st));
    RDM_CHECK_CUDA(cudaMemcpy2DAsync(token_out_dev, sizeof(long long), n->tok_buf + 1, 2 * sizeof(long long), sizeof(long long), B, cudaMemcpyDeviceToDevice, st));
    return RDM_OK;
}

int rdm_rarm_sample(rdm_rarm_t* n, int64_t* tokens_dev, int32_t B, int32_t n_prefix, int32_t steps, float temperature, int32_t top_k, float guidance_scale,
                    const float* uniforms_dev, void* stream) {
    RDM_REQUIRE(n && tokens_dev && B >= 1 && n_prefix >= 1 && steps >= 0 && temperature > 0.f, RDM_ERR_ARG, "rdm_rarm_sample: bad argument");
    const int guided = guidance_scale > 1.f ? 1 : 0, B2 = guided ? 2 * B : B, T = n->cfg.sequence_length, total = n_prefix + steps;
    RDM_REQUIRE(n->B2 == B2, RDM_ERR_STATE, "rdm_rarm_sample: context was set for %d rows, need %d ([r | zeros] when guidance_scale > 1)", n->B2, B2);
    RDM_REQUIRE(total - 1 <= T, RDM_ERR_ARG, "rdm_rarm_sample: %d positions exceed sequence_length %d", total - 1, T);
    if (steps == 0) return RDM_OK;
    DeviceGuard guard(n->device);
    cudaStream_t st = (cudaStream_t)stream;
    RDM_TRY(ensure_half(n, st));
    RDM_TRY(ensure_tok(n, B2));
    const int tcap = T + 1, greedy = uniforms_dev ? 0 : 1;
    // staging: tokens [B, total] -> tok_buf [B, tcap]; uniforms [steps, B] -> u_buf
    RDM_CHECK_CUDA(cudaMemcpy2DAsync(n->tok_buf, (size_t)tcap * sizeof(long long), tokens_dev, (size_t)total * sizeof(long long), (size_t)n_prefix * sizeof(long long), B,
                                     cudaMemcpyDeviceToDevice, st));
    if (uniforms_dev) RDM_CHECK_CUDA(cudaMemcpyAsync(n->u_buf, uniforms_dev, (size_t)steps * B * sizeof(float), cudaMemcpyDeviceToDevice, st));
    const int zero = 0;
    RDM_CHECK_CUDA(cudaMemcpyAsync(n->pos_dev, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
    // one step = {logits = decoder(token at pos); token[pos + 1] = draw (unless given); pos++}
    auto body = [&](cudaStream_t cs) -> int {
        RDM_TRY(step_forward(n, n->tok_buf, B, tcap, cs));
        RDM_TRY(sample_launch(n, n->logits, B, n->cfg.out_channels, guided, guidance_scale, temperature, top_k, n->u_buf, greedy, n_prefix, tcap, n->tok_buf, nullptr, cs));
        rarm_advance_kernel<<<1, 32, 0, cs>>>(n->pos_dev);
        LAUNCH_CHECK();
        return RDM_OK;
    };
    const int iters = total - 1;                    // positions 0 .. total-2 are fed; the last one produces token total-1
    int done = 0;
    if (n->use_graph) {
        long long key[10] = {B, B2, n->kctx, n->mode, guided, (long long)(guidance_scale * 1e6), (long long)(temperature * 1e6), top_k, n_prefix, greedy};
        if (!n->step_exec || memcmp(key, n->step_key, sizeof(key)) != 0) {
            RDM_TRY(body(st));                      // eager first step: sets the kernel attributes, doubles as warm-up
            RDM_CHECK_CUDA(cudaStreamSynchronize(st));
            done = 1;
            RDM_TRY(capture_graph(n, body));
            memcpy(n->step_key, key, sizeof(key));
        }
        for (int i = done; i < iters; i++) { RDM_CHECK_CUDA(cudaGraphLaunch(n->step_exec, st)); g_rdm_launches += n->step_kernels; }
    } else {
        for (int i = 0; i < iters; i++) RDM_TRY(body(st));
    }
    RDM_CHECK_CUDA(cudaMemcpy2DAsync(tokens_dev, (size_t)total * sizeof(long long), n->tok_buf, (size_t)tcap * sizeof(long long), (size_t)total * sizeof(long long), B,
                                     cudaMemcpyDeviceToDevice, st));
    return RDM_OK;
}

}  // extern "C"
