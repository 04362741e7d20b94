// tcgen05 tensor-core implicit GEMM for the U-Net contractions (sm_100a only).
//
//   out[M,N] = epi( A[M,K] * W[N,K]^T ),   A / W stored as bf16 "hi" (+ "lo") planes:
//       x ~= hi + lo,  hi = bf16_rn(x), lo = bf16_rn(x - hi)            (2^-17 relative)
//   NSPLIT = 3:  acc += A_hi*W_hi + A_lo*W_hi + A_hi*W_lo   (fp32-grade contraction on the bf16 tensor pipe)
//   NSPLIT = 1:  acc += A_hi*W_hi                            (plain bf16)
//
// One 128 x BN output tile per work item (persistent CTAs, one per SM), K swept in 64-element blocks; 12 warps:
//   warp 0   : TMA producer -- cp.async.bulk.tensor.4d into 128B-swizzled shared-memory stages; a 3x3 conv is
//              9 shifted box loads of the NHWC plane (TMA out-of-bounds zero fill *is* the conv padding), so no
//              im2col buffer exists; 1x1 convs / Linear layers use the same path with a degenerate box.  Two
//              extensions of the k-block sequence (round 2): a second activation / weight pair appended behind the
//              taps (TcA::hi2: the ResBlock skip_connection), and the 2x2-tap parity form of Upsample + conv (TcA::ups)
//   warp 1   : allocates TMEM, one elected lane issues tcgen05.mma (kind::f16, M=128, N=BN, K=16) with the
//              fp32 accumulator double-buffered in TMEM, tcgen05.commit releases stages / signals the epilogue
//   warps 2-3: idle (setmaxnreg works on whole warpgroups: warpgroup 0 hands its registers to the epilogue)
//   warps 4-11: epilogue, two warps per TMEM lane quarter -- tcgen05.ld 32 lanes x 32 columns, bias / per-sample
//              row vector (time embedding) / SiLU / GEGLU / fused cross-attention / residual, fp32 or 16-bit-plane
//              stores; operands prefetched one chunk ahead
// full/empty mbarrier ring between producer and MMA issuer; tmem_full / tmem_empty mbarriers between issuer and epilogue.
// gemm_tc2_kernel below is the same pipeline on a CTA pair (tcgen05.mma.cta_group::2, M = 256).
#include "kernels.cuh"
#include "gemm_tc.cuh"
#include "ptx.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdlib>

namespace {

// Developer build (tools/ab_build.sh ... -DRDM_AB_TIMING): CTA 0 prints SM-clock offsets of its pipeline milestones.
#ifdef RDM_AB_TIMING
__shared__ long long g_ts[24];
#define TSTAMP(i) do { if (blockIdx.x == 0) g_ts[i] = clock64(); } while (0)
#define TSTAMP_EPI(i) do { if (blockIdx.x == 0 && threadIdx.x == 128 && g_ts[i] == 0) g_ts[i] = clock64(); } while (0)     /* first epilogue warp, first time */
#else
#define TSTAMP(i) do { } while (0)
#define TSTAMP_EPI(i) do { } while (0)
#endif

constexpr int BM = 128, BK = 64, EPI_WARPS = 8, EPI_WARP0 = 4, TC_THREADS = (EPI_WARP0 + EPI_WARPS) * 32;   // warpgroup 0: warp 0 TMA, warp 1 MMA, warps 2-3 idle; warpgroups 1-2: epilogue

// ---- PTX wrappers (mbarrier / TMA helpers live in ptx.cuh) -----------------------------------------------
// erf-form GELU (F.gelu default, ldm FeedForward GEGLU): GELU(x) = x * Phi(x) = max(x, 0) - a * Q(a), a = |x|, with the normal tail
// Q(a) = erfc(a / sqrt 2) / 2 evaluated as 2^-(a p(a) + 1): p is a degree-5 fit of -log2(erfc(a / sqrt 2)) / a on [0, 6] (tools/fit_gelu.py,
// weighted for the absolute error of a Q(a); Q(6) = 1e-9, so a is clamped there).  Branch-free, ONE MUFU.EX2 + 9 FMA-class instructions
// against ~35 for erff with its two divergent ranges -- ncu's source view had erff and the fp16 conversions at half of the GEGLU kernel's
// issue slots.  Max |error| vs the fp64 erf form: 1.2e-7 absolute (1.2e-6 relative above 0.01), i.e. the rounding level of erff itself.
__device__ __forceinline__ float gelu_erf(float x) {
    const float a = fminf(fabsf(x), 6.f);
    float q = -2.992859299411066e-05f;
    q = fmaf(q, a, 0.0007399106398224831f); q = fmaf(q, a, -0.00797757226973772f); q = fmaf(q, a, 0.053238335996866226f);
    q = fmaf(q, a, 0.4589155912399292f); q = fmaf(q, a, 1.1511471271514893f);
    float h;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(fmaf(-a, q, -1.f)));
    return fmaf(-a, h, fmaxf(x, 0.f));
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }

struct TcKernelParams {
    int M, N;                  // valid rows / columns of the output
    int taps, kb_per_tap;      // K blocks: taps * kb_per_tap, each BK=64 wide
    int kb2;                   // K extension (TcA::hi2): kb2 more k-blocks whose A tile comes from the SECOND activation (same pixels, no shift) and whose
                               // weights come from the second weight matrix -- both through the otherwise unused `lo` tensor maps (NSPLIT = 1 only)
    int ksize;                 // 1 or 3 (tap -> (dy,dx)); 2: the 2x2 taps of one output parity of an upsample-folded conv (ups)
    int ups;                   // 1: 3x3 conv over the 2x nearest-upsampled image, folded (TcA::ups): 4 x the work items, one quarter per output
                               // parity (py, px); taps (ty, tx) read source pixel (y + ty + py - 1, x + tx + px - 1); weights of parity q are rows
                               // [q * N, (q + 1) * N) of the [4 N, 4 C] matrix; GEMM row (b, y, x) is stored to output row (b, 2 y + py, 2 x + px)
    int plain;                 // 1: A is a plain [M, C] matrix encoded as (c, m, 1, 1): tile -> x0 = mt*128
    int bw, bh, bb;            // box extents (x, y, batch): bw*bh*bb == 128
    int H, W;                  // logical output grid per image (for tile -> (b,y,x))
    int splits, kb_per_split;  // split-K: work item = (tile, split); partial fp32 accumulators go to `part` [splits][M][N]
    float* part;
    int cluster;               // split-K inside a thread-block cluster: CTA rank = split, partial tiles reduced through distributed shared memory
    int pdl_off;               // the B operand is an activation of a preceding kernel: no prefetch ahead of griddepcontrol.wait
    int pdl;                   // launched with programmatic stream serialization: prefetch weights before griddepcontrol.wait
    // epilogue
    const float* bias; const float* rowvec; int rowvec_ld; int rows_per_batch;
    const float* res; int res_ld; int act;
    const float* xkv; int xkv_ld, xv_off, xk; float xscale_log2e;  // ACT_XATTN (see GemmEpi)
    float* out; int out_ld;                                // fp32 output (or null)
    __nv_bfloat16* out_hi; __nv_bfloat16* out_lo; int out_bf_ld;   // 16-bit-plane output (or null)
    int f16;                                                       // planes are IEEE fp16 instead of bf16
    double* stats; int stats_ld, stats_hw;                         // GroupNorm statistics of the result (GemmEpi::stats), or null
    int geglu_rows;                                                // GEGLU in row form (no residual / row vector, aligned outputs): see geglu_row_form
};

template <int BN, int NSPLIT, int STAGES>
struct TcSmem {
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = (NSPLIT == 3 ? 2 : 1) * A_BYTES + (NSPLIT >= 2 ? 2 : 1) * B_BYTES;   // [A_hi][B_hi][B_lo?][A_lo?]
    static constexpr int EPI_BYTES = EPI_WARPS * 32 * 32 * 4;      // per-epilogue-warp transpose tiles
    static constexpr int STAT_BYTES = 2 * BN * 4;                  // GroupNorm column sums of a tile: [sum | sumsq][BN] fp32
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_BYTES + STAT_BYTES;
    static constexpr int TMEM_COLS = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
    static_assert(2 * BN <= 512, "two accumulator buffers must fit TMEM");
    static_assert(TOTAL <= 232448, "shared memory budget");
};

// ---- epilogue ------------------------------------------------------------------------------------------
// Phase 1: a thread owns one accumulator ROW (TMEM lane) and 32 consecutive columns per tcgen05.ld (the fused cross-attention works on
// that view: 32 columns = one head).  The raw 32x32 block is transposed through a per-warp shared-memory tile.
// Phase 2: lane (r0 = lane / 8, cg = 4 * (lane % 8)) owns columns cg..cg+3 of rows r0, r0+4, ..., r0+28, so every global access of the
// warp covers whole 128-byte lines (64-byte after GEGLU), and bias / time-embedding row vector / residual are ONE float4 per lane (per
// row).  These operands do not depend on the accumulator: they are prefetched into registers (EpiPre) before the accumulator is
// waited for / while the previous chunk is stored -- with ~200 KB of the SM carved out as shared memory there is next to no L1, every
// such load is an L2 round trip, and the epilogue of a short GEMM is latency-, not bandwidth-bound.
constexpr int EPI_WARP_FLOATS = 32 * 32;            // per-warp staging tile, float4 column groups XOR-swizzled by (row & 7): conflict-free
__device__ __forceinline__ float* epi_at(float* stage, int row, int col4) { return stage + row * 32 + (((col4 >> 2) ^ (row & 7)) << 2); }

// Epilogue families: the kernel is instantiated once per family so that a launch only carries (and fetches) the epilogue code it can run.
//   EPI_PLAIN  no activation, bias / time-embedding row (images >= 16 pixels) / residual, full 32-column chunks   -- most layers
//   EPI_GEGLU  ff.net.0 of the SpatialTransformers        EPI_XATTN  to_q with the fused cross-attention
//   EPI_ANY    everything (SiLU, QuickGELU, ragged N, tiny images); also the reference point for the other three
enum { EPI_PLAIN = 0, EPI_GEGLU = 1, EPI_XATTN = 2, EPI_ANY = 3 };
template <int EPI> __device__ __forceinline__ bool epi_is_geglu(const TcKernelParams& p) { return EPI == EPI_ANY ? p.act == ACT_GEGLU : EPI == EPI_GEGLU; }
template <int EPI> __device__ __forceinline__ bool epi_is_xattn(const TcKernelParams& p) { return EPI == EPI_ANY ? p.act == ACT_XATTN : EPI == EPI_XATTN; }

struct EpiPre { float4 b[2]; float4 r[8]; };        // bias (+ row vector) for rows 0..15 / 16..31 of the warp; residual per row

// GEGLU without residual / row vector (the feed-forward of every SpatialTransformer): the "row form" of the epilogue -- a thread keeps its
// accumulator ROW (32 columns = 16 (value, gate) pairs -> 16 outputs = 32 contiguous bytes of the fp16 plane), so there is no transpose
// through shared memory at all and the 16 GELUs of a thread are independent (the transposed form left the two epilogue warps of a
// scheduler waiting on shared-memory round trips between short dependent chains).  Its only operand is the chunk's bias: 32 floats, the
// same for every lane (broadcast loads), parked in the registers of EpiPre::r.
__device__ __forceinline__ bool geglu_row_form(const TcKernelParams& p) { return p.geglu_rows != 0; }
template <int EPI>
__device__ __forceinline__ void epi_prefetch(const TcKernelParams& p, int lane, int m_warp0, int nb, EpiPre& pre) {
    if (epi_is_geglu<EPI>(p) && geglu_row_form(p)) {
#pragma unroll
        for (int i = 0; i < 8; i++) pre.r[i] = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + nb) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const int cg = (lane & 7) * 4, r0 = lane >> 3, n = nb + cg;
    float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) bz = __ldg(reinterpret_cast<const float4*>(p.bias + n));
    pre.b[0] = bz; pre.b[1] = bz;
    if (p.rowvec && p.rows_per_batch >= 16) {       // 16 | rows_per_batch (power-of-two grids): rows 0..15 and 16..31 each lie in one image
#pragma unroll
        for (int s = 0; s < 2; s++) {
            int row = m_warp0 + 16 * s; row = row < p.M ? row : p.M - 1;
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.rowvec + (size_t)(row / p.rows_per_batch) * p.rowvec_ld + n));
            pre.b[s].x += t.x; pre.b[s].y += t.y; pre.b[s].z += t.z; pre.b[s].w += t.w;
        }
    }
    // (branches hoisted out of the unrolled row loops: a skipped block inside an unrolled loop breaks the sequential instruction fetch 8 times)
    if (!p.res) {
#pragma unroll
        for (int it = 0; it < 8; it++) pre.r[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else if (epi_is_geglu<EPI>(p)) {
        const int no = n >> 1;
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int mo = m_warp0 + r0 + 4 * it;
            pre.r[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (mo < p.M) { const float2 t = __ldcs(reinterpret_cast<const float2*>(p.res + (size_t)mo * p.res_ld + no)); pre.r[it].x = t.x; pre.r[it].y = t.y; }
        }
    } else {
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int mo = m_warp0 + r0 + 4 * it;
            pre.r[it] = mo < p.M ? __ldcs(reinterpret_cast<const float4*>(p.res + (size_t)mo * p.res_ld + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// ragged last chunk (N not a multiple of 32; only the 4-channel output conv): the row goes through the warp's staging tile so that a
// ROLLED scalar loop can index it (a rolled loop over registers would put them in local memory; an unrolled one is 32x the code)
__device__ __forceinline__ void epilogue_ragged(const TcKernelParams& p, const uint32_t (&r)[32], float* stage, int lane, int m, int nb) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(epi_at(stage, lane, j)) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
    __syncwarp();
    if (m >= p.M) return;
    const int bidx = m / p.rows_per_batch, nv = p.N - nb < 32 ? p.N - nb : 32;
#pragma unroll 1
    for (int j = 0; j < nv; j++) {
        const int n = nb + j;
        float t = epi_at(stage, lane, j & ~3)[j & 3];
        if (p.bias) t += __ldg(p.bias + n);
        if (p.rowvec) t += __ldg(p.rowvec + (size_t)bidx * p.rowvec_ld + n);
        if (p.act == ACT_SILU) t = silu_f(t);
        else if (p.act == ACT_QUICKGELU) t = t / (1.f + expf(-1.702f * t));
        if (p.res) t += p.res[(size_t)m * p.res_ld + n];
        if (p.out) p.out[(size_t)m * p.out_ld + n] = t;
        else {
            unsigned short h, l; split16(t, p.f16, h, l);
            reinterpret_cast<unsigned short*>(p.out_hi)[(size_t)m * p.out_bf_ld + n] = h;
            if (p.out_lo) reinterpret_cast<unsigned short*>(p.out_lo)[(size_t)m * p.out_bf_ld + n] = l;
        }
    }
}

// fused cross-attention on one row's 32-column chunk (= the query of head nb/32 for token m): attend to the (<= 8) retrieved-context keys
// of this sample.  The K/V rows of this head (1-2 images per warp) are staged once per chunk in the warp's transpose tile and read back as
// shared-memory broadcasts.  Online softmax in a deliberately ROLLED key loop: a chunk executes this once, and straight-line code of that
// size would be fetched from L2 every time (instruction-cache misses cost more than the arithmetic).
// The K/V rows of one chunk (one head; the 32 rows of a warp lie in at most TWO images: rows_per_batch is a multiple of 16, checked on the
// host): [image][K | V][xk][32 floats] = at most 2 * 2 * 8 * 8 float4, i.e. 8 per lane.  They do not depend on the accumulator, so they are
// requested a whole chunk ahead (before the accumulator is waited for / while the previous head is processed) and parked in registers.
struct XPre { float4 v[8]; };
__device__ __forceinline__ void xattn_prefetch(const TcKernelParams& p, int lane, int m_warp0, int nb, XPre& xp) {
    const int hw = p.rows_per_batch;
    const int b0 = (m_warp0 < p.M ? m_warp0 : p.M - 1) / hw, b1 = (m_warp0 + 31 < p.M ? m_warp0 + 31 : p.M - 1) / hw;
    const int half_img = p.xk * 8, per_img = 2 * half_img, tot = (b1 - b0 + 1) * per_img;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int i = lane + 32 * j;
        if (i < tot) {
            const int img = i >= per_img ? 1 : 0, r = i - img * per_img, kvsel = r >= half_img ? 1 : 0, r2 = r - kvsel * half_img;
            xp.v[j] = __ldg(reinterpret_cast<const float4*>(p.xkv + ((size_t)(b0 + img) * p.xk + (r2 >> 3)) * p.xkv_ld + nb + kvsel * p.xv_off) + (r2 & 7));
        }
    }
}
__device__ __forceinline__ void epilogue_xattn(const TcKernelParams& p, float (&v)[32], float* stage, int lane, int m_warp0, int nb, const XPre& xp) {
    const int m = m_warp0 + lane, mm = m < p.M ? m : p.M - 1, hw = p.rows_per_batch;
    const int b0 = (m_warp0 < p.M ? m_warp0 : p.M - 1) / hw, b1 = (m_warp0 + 31 < p.M ? m_warp0 + 31 : p.M - 1) / hw;
    const int per_img = 2 * p.xk * 8;                              // float4 pieces per image: [K | V][xk][32 floats]
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int i = lane + 32 * j;
        if (i < (b1 - b0 + 1) * per_img) *reinterpret_cast<float4*>(stage + (size_t)i * 4) = xp.v[j];
    }
    __syncwarp();
    const float* kb = stage + (size_t)(mm / hw - b0) * per_img * 4;
    float o[32], mx = -3.0e38f, l = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i++) o[i] = 0.f;
#pragma unroll 1
    for (int j = 0; j < p.xk; j++) {
        const float4* kr = reinterpret_cast<const float4*>(kb + j * 32);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            const float4 t = kr[i], u = kr[i + 1];
            s0 = fmaf(v[4 * i], t.x, s0); s0 = fmaf(v[4 * i + 1], t.y, s0); s0 = fmaf(v[4 * i + 2], t.z, s0); s0 = fmaf(v[4 * i + 3], t.w, s0);
            s1 = fmaf(v[4 * i + 4], u.x, s1); s1 = fmaf(v[4 * i + 5], u.y, s1); s1 = fmaf(v[4 * i + 6], u.z, s1); s1 = fmaf(v[4 * i + 7], u.w, s1);
        }
        const float sj = (s0 + s1) * p.xscale_log2e, mn = fmaxf(mx, sj), corr = exp2f(mx - mn), pj = exp2f(sj - mn);
        mx = mn; l = fmaf(l, corr, pj);
        const float4* vr = reinterpret_cast<const float4*>(kb + (p.xk + j) * 32);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 t = vr[i];
            o[4 * i] = fmaf(pj, t.x, o[4 * i] * corr); o[4 * i + 1] = fmaf(pj, t.y, o[4 * i + 1] * corr);
            o[4 * i + 2] = fmaf(pj, t.z, o[4 * i + 2] * corr); o[4 * i + 3] = fmaf(pj, t.w, o[4 * i + 3] * corr);
        }
    }
    const float il = 1.f / l;
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = o[i] * il;
}

// GroupNorm statistics of the finished values (the consumer of this GEMM's output is a GroupNorm: ResBlock in/out_layers, the
// SpatialTransformer norm, the output head): per-(image, column) sum and sum of squares.  The lane holds 4 columns x 8 rows
// (r0 + 4 it): fp32 partial sums over the warp's 32 rows, two xor-shuffles across the row groups, then SHARED-memory fp32 atomics from the
// 8 lanes of row group 0 into the tile's column sums (a 128-row tile lies in one image: stats_hw is a multiple of 128).  At the end of
// the tile the epilogue warps flush the 2 * BN sums with fp64 global atomics (epi_stats_flush) -- one atomic per column and tile instead
// of one per column, warp and chunk (the first version: 393 K fp64 atomics per level-0 GEMM cost more than the statistics pass saved).
__device__ __forceinline__ void epi_stats(const TcKernelParams& p, const float4 (&t)[8], int lane, int m_warp0, int col, float* s_stat, int BN) {
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    const int r0 = lane >> 3;
#pragma unroll
    for (int it = 0; it < 8; it++) {
        if (m_warp0 + r0 + 4 * it < p.M) {
            s[0] += t[it].x; q[0] = fmaf(t[it].x, t[it].x, q[0]); s[1] += t[it].y; q[1] = fmaf(t[it].y, t[it].y, q[1]);
            s[2] += t[it].z; q[2] = fmaf(t[it].z, t[it].z, q[2]); s[3] += t[it].w; q[3] = fmaf(t[it].w, t[it].w, q[3]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        s[j] += __shfl_xor_sync(0xffffffffu, s[j], 8); s[j] += __shfl_xor_sync(0xffffffffu, s[j], 16);
        q[j] += __shfl_xor_sync(0xffffffffu, q[j], 8); q[j] += __shfl_xor_sync(0xffffffffu, q[j], 16);
    }
    if (r0 == 0) {
#pragma unroll
        for (int j = 0; j < 4; j++) { atomicAdd(s_stat + col + j, s[j]); atomicAdd(s_stat + BN + col + j, q[j]); }
    }
}
// all EPI_WARPS * 32 epilogue threads, once per tile: column sums of the tile -> fp64 global accumulators of image m_tile0 / stats_hw; the
// array is zeroed for the next tile (second named barrier)
__device__ __forceinline__ void epi_stats_flush(const TcKernelParams& p, float* s_stat, int BN, int m_tile0, int n0, int tid_epi) {
    asm volatile("bar.sync 2, %0;" ::"r"(EPI_WARPS * 32) : "memory");
    double* d = p.stats + ((size_t)(m_tile0 / p.stats_hw) * p.stats_ld + n0) * 2;
    for (int i = tid_epi; i < 2 * BN; i += EPI_WARPS * 32) {
        const int which = i >= BN ? 1 : 0, c = i - which * BN;
        if (n0 + c < p.N) atomicAdd(d + 2 * c + which, (double)s_stat[i]);
        s_stat[i] = 0.f;
    }
    asm volatile("bar.sync 2, %0;" ::"r"(EPI_WARPS * 32) : "memory");
}

// r: this lane's 32 accumulators (row m_warp0 + lane, columns nb..nb+31, nb + 32 <= N); stage: this warp's smem tile; pre / xp: operands of
// THIS chunk (epi_prefetch / xattn_prefetch), requested by the caller one whole chunk earlier.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const TcKernelParams& p, const uint32_t (&r)[32], float* stage, int lane, int m_warp0, int nb, const EpiPre& pre, XPre& xp, float* part,
                                               float* s_stat, int stat_col, int BN, int par) {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
    TSTAMP_EPI(8);
    const int cg = (lane & 7) * 4, r0 = lane >> 3, n = nb + cg;
    if (epi_is_xattn<EPI>(p) && !part) {           // no bias / residual operands on this path (checked on the host): `pre` is dead here
        if (EPI == EPI_ANY) xattn_prefetch(p, lane, m_warp0, nb, xp);       // the all-in-one kernel loads just in time (register budget)
        epilogue_xattn(p, v, stage, lane, m_warp0, nb, xp);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(epi_at(stage, lane, j)) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int mo = m_warp0 + r0 + 4 * it;
            const float4 o = *reinterpret_cast<const float4*>(epi_at(stage, r0 + 4 * it, cg));
            if (mo < p.M) {
                if (p.out) *reinterpret_cast<float4*>(p.out + (size_t)mo * p.out_ld + n) = o;
                else store_planes4(p.out_hi + (size_t)mo * p.out_bf_ld + n, p.out_lo ? p.out_lo + (size_t)mo * p.out_bf_ld + n : nullptr, p.f16, o.x, o.y, o.z, o.w);
            }
        }
        return;
    }
    if (epi_is_geglu<EPI>(p) && geglu_row_form(p) && !part) {
        const int m = m_warp0 + lane, no = nb >> 1;
        float o[16];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 bb = pre.r[i];
            o[2 * i] = (v[4 * i] + bb.x) * gelu_erf(v[4 * i + 1] + bb.y);
            o[2 * i + 1] = (v[4 * i + 2] + bb.z) * gelu_erf(v[4 * i + 3] + bb.w);
        }
        if (m < p.M) {
            if (p.out) {
                float4* d = reinterpret_cast<float4*>(p.out + (size_t)m * p.out_ld + no);
#pragma unroll
                for (int i = 0; i < 4; i++) d[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
            } else if (!p.out_lo) {
                uint4* d = reinterpret_cast<uint4*>(p.out_hi + (size_t)m * p.out_bf_ld + no);
                d[0] = make_uint4(pack16x2(o[0], o[1], p.f16), pack16x2(o[2], o[3], p.f16), pack16x2(o[4], o[5], p.f16), pack16x2(o[6], o[7], p.f16));
                d[1] = make_uint4(pack16x2(o[8], o[9], p.f16), pack16x2(o[10], o[11], p.f16), pack16x2(o[12], o[13], p.f16), pack16x2(o[14], o[15], p.f16));
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++)
                    store_planes4(p.out_hi + (size_t)m * p.out_bf_ld + no + 4 * i, p.out_lo + (size_t)m * p.out_bf_ld + no + 4 * i, p.f16, o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
            }
        }
        return;
    }
    // transpose through shared memory
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(epi_at(stage, lane, j)) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    __syncwarp();
    if (part) {                                    // split-K partial tile [M, N] fp32: raw accumulators
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int mo = m_warp0 + r0 + 4 * it;
            if (mo < p.M) *reinterpret_cast<float4*>(part + (size_t)mo * p.N + n) = *reinterpret_cast<const float4*>(epi_at(stage, r0 + 4 * it, cg));
        }
        return;
    }
    const bool slow_rv = EPI == EPI_ANY && p.rowvec && p.rows_per_batch < 16;
    if (epi_is_geglu<EPI>(p)) {
        float4 t[8];
#pragma unroll
        for (int it = 0; it < 8; it++) t[it] = *reinterpret_cast<const float4*>(epi_at(stage, r0 + 4 * it, cg));
        TSTAMP_EPI(9);
        const int no = n >> 1;
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int mo = m_warp0 + r0 + 4 * it;
            if (mo < p.M) {
                const float4 bb = pre.b[it >> 2];
                const float o0 = (t[it].x + bb.x) * gelu_erf(t[it].y + bb.y) + pre.r[it].x, o1 = (t[it].z + bb.z) * gelu_erf(t[it].w + bb.w) + pre.r[it].y;
                if (p.out) *reinterpret_cast<float2*>(p.out + (size_t)mo * p.out_ld + no) = make_float2(o0, o1);
                else if (!p.out_lo) *reinterpret_cast<uint32_t*>(p.out_hi + (size_t)mo * p.out_bf_ld + no) = pack16x2(o0, o1, p.f16);
                else {
                    unsigned short h0, l0, h1, l1; split16(o0, p.f16, h0, l0); split16(o1, p.f16, h1, l1);
                    *reinterpret_cast<uint32_t*>(p.out_hi + (size_t)mo * p.out_bf_ld + no) = (uint32_t)h0 | ((uint32_t)h1 << 16);
                    *reinterpret_cast<uint32_t*>(p.out_lo + (size_t)mo * p.out_bf_ld + no) = (uint32_t)l0 | ((uint32_t)l1 << 16);
                }
            }
        }
    } else if (EPI == EPI_ANY && (p.act != ACT_NONE || slow_rv)) {
        // Rare variants (SiLU of the time-embedding MLP, CLIP's QuickGELU, images smaller than 16 pixels): a ROLLED row loop that reads its
        // operands straight from shared / global memory.  Keeping these ~2000 instructions out of the unrolled loop below shrinks the code the
        // common layers have to fetch (the kernel starts with a cold instruction cache every launch; ncu: `no_instruction` stalls).
#pragma unroll 1
        for (int it = 0; it < 8; it++) {
            const int mo = m_warp0 + r0 + 4 * it;
            if (mo >= p.M) continue;
            float4 o = *reinterpret_cast<const float4*>(epi_at(stage, r0 + 4 * it, cg));
            const float4 bb = it < 4 ? pre.b[0] : pre.b[1];
            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
            if (slow_rv) {
                const float4 rv = __ldg(reinterpret_cast<const float4*>(p.rowvec + (size_t)(mo / p.rows_per_batch) * p.rowvec_ld + n));
                o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
            }
            if (p.act == ACT_SILU) { o.x = silu_f(o.x); o.y = silu_f(o.y); o.z = silu_f(o.z); o.w = silu_f(o.w); }
            else if (p.act == ACT_QUICKGELU) {
                o.x = o.x / (1.f + expf(-1.702f * o.x)); o.y = o.y / (1.f + expf(-1.702f * o.y));
                o.z = o.z / (1.f + expf(-1.702f * o.z)); o.w = o.w / (1.f + expf(-1.702f * o.w));
            }
            if (p.res) { const float4 q = *reinterpret_cast<const float4*>(p.res + (size_t)mo * p.res_ld + n); o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w; }
            if (p.out) *reinterpret_cast<float4*>(p.out + (size_t)mo * p.out_ld + n) = o;
            else store_planes4(p.out_hi + (size_t)mo * p.out_bf_ld + n, p.out_lo ? p.out_lo + (size_t)mo * p.out_bf_ld + n : nullptr, p.f16, o.x, o.y, o.z, o.w);
        }
    } else {
        // the common layers: bias (+ time-embedding row) and residual from the prefetched registers, fp32 or 16-bit-plane store
        float4 t[8];
#pragma unroll
        for (int it = 0; it < 8; it++) t[it] = *reinterpret_cast<const float4*>(epi_at(stage, r0 + 4 * it, cg));
        TSTAMP_EPI(9);
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const float4 bb = pre.b[it >> 2];
            t[it] = make_float4(t[it].x + bb.x + pre.r[it].x, t[it].y + bb.y + pre.r[it].y, t[it].z + bb.z + pre.r[it].z, t[it].w + bb.w + pre.r[it].w);
        }
        if (p.stats) epi_stats(p, t, lane, m_warp0, stat_col + cg, s_stat, BN);
        if (p.ups) {                               // upsample-folded conv: GEMM row (b, y, x) of parity (py, px) is output pixel (b, 2 y + py, 2 x + px)
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int mo = m_warp0 + r0 + 4 * it;
                if (mo >= p.M) continue;
                const int x = mo % p.W, yb = mo / p.W, y = yb % p.H, b = yb / p.H;
                const size_t row = ((size_t)(b * 2 * p.H + 2 * y + (par >> 1)) * (2 * p.W) + 2 * x + (par & 1));
                if (p.out) *reinterpret_cast<float4*>(p.out + row * p.out_ld + n) = t[it];
                else store_planes4(p.out_hi + row * p.out_bf_ld + n, p.out_lo ? p.out_lo + row * p.out_bf_ld + n : nullptr, p.f16, t[it].x, t[it].y, t[it].z, t[it].w);
            }
        } else if (p.out) {
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int mo = m_warp0 + r0 + 4 * it;
                if (mo < p.M) *reinterpret_cast<float4*>(p.out + (size_t)mo * p.out_ld + n) = t[it];
            }
        } else {
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int mo = m_warp0 + r0 + 4 * it;
                if (mo < p.M) store_planes4(p.out_hi + (size_t)mo * p.out_bf_ld + n, p.out_lo ? p.out_lo + (size_t)mo * p.out_bf_ld + n : nullptr, p.f16, t[it].x, t[it].y, t[it].z, t[it].w);
            }
        }
    }
    TSTAMP_EPI(10);
}

// epilogue of 4 consecutive accumulator columns n..n+3 of row m (bias, time-embedding row vector, activation, residual, store)
__device__ __forceinline__ void epi_store4(const TcKernelParams& p, int m, int n, float (&v)[4]) {
    if (p.bias) { float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + n)); v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w; }
    if (p.rowvec) { float4 t = __ldg(reinterpret_cast<const float4*>(p.rowvec + (size_t)(m / p.rows_per_batch) * p.rowvec_ld + n)); v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w; }
    if (p.act == ACT_GEGLU) {
        float o0 = v[0] * gelu_erf(v[1]), o1 = v[2] * gelu_erf(v[3]);
        const int no = n >> 1;
        if (p.res) { o0 += p.res[(size_t)m * p.res_ld + no]; o1 += p.res[(size_t)m * p.res_ld + no + 1]; }
        if (p.out) *reinterpret_cast<float2*>(p.out + (size_t)m * p.out_ld + no) = make_float2(o0, o1);
        else {
            unsigned short h0, l0, h1, l1; split16(o0, p.f16, h0, l0); split16(o1, p.f16, h1, l1);
            *reinterpret_cast<uint32_t*>(p.out_hi + (size_t)m * p.out_bf_ld + no) = (uint32_t)h0 | ((uint32_t)h1 << 16);
            if (p.out_lo) *reinterpret_cast<uint32_t*>(p.out_lo + (size_t)m * p.out_bf_ld + no) = (uint32_t)l0 | ((uint32_t)l1 << 16);
        }
        return;
    }
    if (p.act == ACT_SILU) { v[0] = silu_f(v[0]); v[1] = silu_f(v[1]); v[2] = silu_f(v[2]); v[3] = silu_f(v[3]); }
    else if (p.act == ACT_QUICKGELU) {
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = v[j] / (1.f + expf(-1.702f * v[j]));
    }
    if (p.res) { float4 t = *reinterpret_cast<const float4*>(p.res + (size_t)m * p.res_ld + n); v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w; }
    if (p.out) *reinterpret_cast<float4*>(p.out + (size_t)m * p.out_ld + n) = make_float4(v[0], v[1], v[2], v[3]);
    else store_planes4(p.out_hi + (size_t)m * p.out_bf_ld + n, p.out_lo ? p.out_lo + (size_t)m * p.out_bf_ld + n : nullptr, p.f16, v[0], v[1], v[2], v[3]);
}

// Fallback (RDM_TC_CLUSTER=0): out = epi( sum_s part[s] ), 4 consecutive columns of one row per thread
__global__ void splitk_reduce_kernel(const TcKernelParams p) {
    pdl_wait();
    pdl_launch_dependents();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int N4 = p.N >> 2;
    if (i >= (long long)p.M * N4) return;
    const int m = (int)(i / N4), n = (int)(i % N4) * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < p.splits; s++) {
        float4 t = __ldcs(reinterpret_cast<const float4*>(p.part + ((size_t)s * p.M + m) * p.N + n));
        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
    }
    float v[4] = {a.x, a.y, a.z, a.w};
    epi_store4(p, m, n, v);
}

// Persistent: grid = min(#tiles, #SMs); CTA c handles tiles c, c+grid, ...; tile -> (mt, nt) with nt fastest so
// co-scheduled CTAs share the A tile in L2.  TMEM holds TWO accumulators: the epilogue of tile i overlaps the MMAs of i+1.
template <int BN, int NSPLIT, int STAGES, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, const TcKernelParams p) {
    using S = TcSmem<BN, NSPLIT, STAGES>;
    constexpr int RED_LD = BN + 4;                 // row stride (floats) of the split-K staging tile: conflict-free float4 rows
    static_assert(BM * RED_LD * 4 <= STAGES * S::STAGE_BYTES, "split-K staging tile must fit the pipeline stages");
    extern __shared__ uint8_t smem_raw[];
#ifdef RDM_AB_GENERIC_SMEM
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
#else
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // 1 KB aligned; pointer arithmetic keeps the shared address space
#endif
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;          // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* epi_stage = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES + 256);
    float* stat_base = epi_stage + EPI_WARPS * EPI_WARP_FLOATS;          // [2 * BN]
    for (int i = threadIdx.x; i < 2 * BN; i += TC_THREADS) stat_base[i] = 0.f;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb_main = p.taps * p.kb_per_tap, nkb = nkb_main + (NSPLIT == 1 ? p.kb2 : 0);
    const int ntn = (p.N + BN - 1) / BN, ntm = (p.M + BM - 1) / BM, nitems_q = ntn * ntm * p.splits, nitems = nitems_q * (p.ups ? 4 : 1);
#ifdef RDM_AB_TIMING
    unsigned long long gt0 = 0;
    if (threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0)); for (int i = 0; i < 24; i++) g_ts[i] = 0; TSTAMP(0); }
    __syncthreads();
#endif
    pdl_launch_dependents();                       // the next kernel may start its own prologue as soon as every CTA of this grid runs

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA_hi); prefetch_tmap(&tmB_hi);
        if (NSPLIT >= 2) prefetch_tmap(&tmB_lo);
        if (NSPLIT == 3 || (NSPLIT == 1 && p.kb2)) prefetch_tmap(&tmA_lo);
        if (NSPLIT == 1 && p.kb2) prefetch_tmap(&tmB_lo);
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)S::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TSTAMP(1);

    // Register re-allocation between the warp roles (setmaxnreg acts on whole WARPGROUPS, hence the 4-warp issuer group): 12 warps start
    // with 168 registers/thread; warpgroup 0 (TMA issuer, MMA issuer, two idle warps) gives registers back, the two epilogue warpgroups
    // (operand prefetch + 32 accumulators + transpose tiles) take them: 4 x 32 x 56 + 8 x 32 x 224 = 64512 = 384 x 168.
    if (warp < EPI_WARP0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        if (lane == 0) {
            // PDL prologue: the WEIGHT tiles of the first stages do not depend on the previous kernel -> request them before waiting
            int pre = 0;
            {
                const int item0 = blockIdx.x;
                if (p.pdl && item0 < nitems) {
                    const int par = p.ups ? item0 / nitems_q : 0, item = item0 - par * nitems_q;
                    const int tile = item / p.splits, sp = item - tile * p.splits;
                    const int kb0 = sp * p.kb_per_split, kb1 = min(nkb, kb0 + p.kb_per_split), n0 = (tile % ntn) * BN + par * p.N;
                    for (int kb = kb0; kb < kb1 && pre < STAGES; kb++, pre++) {
                        uint8_t* st = smem + pre * S::STAGE_BYTES;
                        mbar_expect_tx(&full[pre], S::STAGE_BYTES);
                        if (NSPLIT == 1 && kb >= nkb_main) tma_load_2d(st + S::A_BYTES, &tmB_lo, &full[pre], (kb - nkb_main) * BK, n0);
                        else tma_load_2d(st + S::A_BYTES, &tmB_hi, &full[pre], kb * BK, n0);
                        if (NSPLIT >= 2) tma_load_2d(st + S::A_BYTES + S::B_BYTES, &tmB_lo, &full[pre], kb * BK, n0);
                    }
                }
            }
            pdl_wait();                                                  // activations / residuals of the previous kernels are now visible
            TSTAMP(2);
            int it = 0;                                                  // global k-block counter (ring position)
            for (int item0 = blockIdx.x; item0 < nitems; item0 += gridDim.x) {
                const int par = p.ups ? item0 / nitems_q : 0, item = item0 - par * nitems_q;
                const int tile = item / p.splits, sp = item - tile * p.splits;
                const int kb0 = sp * p.kb_per_split, kb1 = min(nkb, kb0 + p.kb_per_split);
                const int mt = tile / ntn, n0 = (tile % ntn) * BN + par * p.N;     // (ups: row block of this parity's weights)
                int x0 = 0, y0 = 0, b0 = 0;                              // tile origin in (x, y, b) of the NHWC plane
                if (p.plain) x0 = mt * BM;
                else if (p.W > BM) { const int tpr = p.W / BM; x0 = (mt % tpr) * BM; y0 = (mt / tpr) % p.H; b0 = mt / (tpr * p.H); }   // wide images: a tile is a 128-pixel piece of one row
                else if (p.bb > 1 || p.bh * p.bw == p.H * p.W) b0 = mt * p.bb;
                else { int tiles_per_img = (p.H * p.W) / BM; b0 = mt / tiles_per_img; y0 = (mt % tiles_per_img) * p.bh; }
                for (int kb = kb0; kb < kb1; kb++, it++) {
                    const int s = it % STAGES; const uint32_t ph = (it / STAGES) & 1;
                    uint8_t* st = smem + s * S::STAGE_BYTES;
                    const bool prefetched = it < pre;                    // weights already requested (and the barrier armed) before pdl_wait
                    if (!prefetched) { mbar_wait(&empty[s], ph ^ 1); mbar_expect_tx(&full[s], S::STAGE_BYTES); }
                    if (NSPLIT == 1 && kb >= nkb_main) {                // K extension: the second activation / weight pair (ResBlock skip_connection)
                        const int kc2 = (kb - nkb_main) * BK;
                        tma_load_4d(st, &tmA_lo, &full[s], kc2, x0, y0, b0);
                        if (!prefetched) tma_load_2d(st + S::A_BYTES, &tmB_lo, &full[s], kc2, n0);
                        continue;
                    }
                    const int tap = kb / p.kb_per_tap, kc = (kb - tap * p.kb_per_tap) * BK;
                    const int dy = p.ksize == 3 ? tap / 3 - 1 : p.ksize == 2 ? (tap >> 1) + (par >> 1) - 1 : 0;
                    const int dx = p.ksize == 3 ? tap % 3 - 1 : p.ksize == 2 ? (tap & 1) + (par & 1) - 1 : 0;
                    tma_load_4d(st, &tmA_hi, &full[s], kc, x0 + dx, y0 + dy, b0);
                    if (!prefetched) {
                        tma_load_2d(st + S::A_BYTES, &tmB_hi, &full[s], kb * BK, n0);
                        if (NSPLIT >= 2) tma_load_2d(st + S::A_BYTES + S::B_BYTES, &tmB_lo, &full[s], kb * BK, n0);
                    }
                    if (NSPLIT == 3) tma_load_4d(st + S::A_BYTES + 2 * S::B_BYTES, &tmA_lo, &full[s], kc, x0 + dx, y0 + dy, b0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(BN, p.f16);
            int it = 0, lt = 0;                                          // ring position, local tile counter
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, lt++) {
                const int sp = item % p.splits;
                const int kb0 = sp * p.kb_per_split, kb1 = min(nkb, kb0 + p.kb_per_split);
                const int buf = lt & 1;
                mbar_wait(&tmem_empty[buf], ((lt >> 1) & 1) ^ 1);        // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
                for (int kb = kb0; kb < kb1; kb++, it++) {
                    const int s = it % STAGES; const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
#ifdef RDM_AB_TIMING
                    if (it == 0) TSTAMP(3);
#endif
                    const uint32_t a_hi = smem_u32(smem + s * S::STAGE_BYTES), b_hi = a_hi + S::A_BYTES;
                    const uint32_t b_lo = b_hi + S::B_BYTES, a_lo = b_lo + S::B_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; k++) {
                        const uint64_t da = umma_desc_sw128(a_hi + k * 32), db = umma_desc_sw128(b_hi + k * 32);
                        umma_bf16(tmem_d, da, db, idesc, ((kb - kb0) | k) != 0);
                        if (NSPLIT >= 2) umma_bf16(tmem_d, da, umma_desc_sw128(b_lo + k * 32), idesc, 1);
                        if (NSPLIT == 3) umma_bf16(tmem_d, umma_desc_sw128(a_lo + k * 32), db, idesc, 1);
                    }
                    umma_commit(&empty[s]);                // frees the stage once the MMAs above have read it
                }
                umma_commit(&tmem_full[buf]);              // accumulator complete
                TSTAMP(4);
            }
        }
    }
    } else {
        // epilogue: warp w may touch TMEM lanes [32*(w%4), +32); each lane quarter is served by TWO warps that alternate column chunks
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        pdl_wait();                                   // residual / row-vector operands come from previous kernels
        const int q = warp & 3, half = (warp - EPI_WARP0) >> 2, ew = warp - EPI_WARP0;
        int lt = 0;
        if (p.cluster) {
            // cluster split-K: park this CTA's fp32 partial tile in (now idle) pipeline shared memory; the cluster-wide reduction follows below
            mbar_wait(&tmem_full[0], 0);
            tc_fence_after();
            float* red = reinterpret_cast<float*>(smem) + (size_t)(q * 32 + lane) * RED_LD;
#pragma unroll 1
            for (int c = half; c < BN / 32; c += 2) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(red + c * 32 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            }
        } else
        for (int item0 = blockIdx.x; item0 < nitems; item0 += gridDim.x, lt++) {
            const int par = p.ups ? item0 / nitems_q : 0, item = item0 - par * nitems_q;
            const int tile = item / p.splits, sp = item - tile * p.splits;
            float* part = p.splits > 1 ? p.part + (size_t)sp * p.M * p.N : nullptr;    // raw partial sums; the reduce kernel applies the epilogue
            const int buf = lt & 1;
            const int mt = tile / ntn, n0 = (tile % ntn) * BN;
            const int m_warp0 = mt * BM + q * 32;
            // Epilogue operands (bias / time-embedding row / residual, or the K/V rows of the fused cross-attention) are requested ONE WHOLE
            // CHUNK ahead: the first chunk's before the accumulator is waited for (overlaps the mainloop), chunk c + 1's before chunk c is
            // read from TMEM -- an L2 round trip (~700 cycles; ncu: long-scoreboard stalls on the residual adds and the stores) is then
            // covered by the TMEM load, the transpose and the stores of a full chunk instead of being paid per chunk.
            EpiPre pre, pre_nx; XPre xp, xp_nx;
            const bool xat = EPI == EPI_XATTN && !part;
            if (!part && m_warp0 < p.M && n0 + half * 32 + 32 <= p.N) {
                if (xat) xattn_prefetch(p, lane, m_warp0, n0 + half * 32, xp);
                else epi_prefetch<EPI>(p, lane, m_warp0, n0 + half * 32, pre);
            }
            mbar_wait(&tmem_full[buf], (lt >> 1) & 1);
            tc_fence_after();
#ifdef RDM_AB_TIMING
            if (warp == EPI_WARP0 && lane == 0 && lt == 0) TSTAMP(5);
#endif
#pragma unroll 1
            for (int c = half; c < BN / 32; c += 2) {
                const int nb = n0 + c * 32;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c * 32);
                if (m_warp0 >= p.M || nb >= p.N) continue;
                const int nbn = nb + 64;                           // this warp's next chunk of the tile
                const bool has_next = !part && c + 2 < BN / 32 && nbn + 32 <= p.N;
                if (has_next) {
                    if (xat) xattn_prefetch(p, lane, m_warp0, nbn, xp_nx);
                    else epi_prefetch<EPI>(p, lane, m_warp0, nbn, pre_nx);
                }
                uint32_t r[32];
                tmem_ld32(taddr, r);
                if (nb + 32 <= p.N) epilogue_chunk<EPI>(p, r, epi_stage + ew * EPI_WARP_FLOATS, lane, m_warp0, nb, pre, xp, part, stat_base, c * 32, BN, par);
                else if (EPI == EPI_ANY) epilogue_ragged(p, r, epi_stage + ew * EPI_WARP_FLOATS, lane, m_warp0 + lane, nb);     // (split-K needs N % 4 == 0 ... never a partial tile here)
                if (has_next) { if (xat) xp = xp_nx; else pre = pre_nx; }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[buf])) : "memory");
            if (p.stats && !part) epi_stats_flush(p, stat_base, BN, mt * BM, n0, (int)threadIdx.x - EPI_WARP0 * 32);
#ifdef RDM_AB_TIMING
            if (warp == EPI_WARP0 && lane == 0) TSTAMP(6);
#endif
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.cluster) {
        // Every CTA of the cluster (rank = split) now holds a 128 x BN fp32 partial in its shared memory.  CTA r reduces the r-th slice of
        // the tile over all ranks through distributed shared memory (fixed summation order: deterministic) and applies the real epilogue.
        cluster_sync_all();
        if (warp >= EPI_WARP0) {
            const int tile = blockIdx.x / p.splits, rank = blockIdx.x - tile * p.splits;
            const int mt = tile / ntn, n0 = (tile % ntn) * BN;
            constexpr int C4 = BN / 4, TOTAL4 = BM * C4;
            const int i0 = (int)((long long)TOTAL4 * rank / p.splits), i1 = (int)((long long)TOTAL4 * (rank + 1) / p.splits);
            const uint32_t red_base = smem_u32(smem);
            for (int i = i0 + (int)threadIdx.x - EPI_WARP0 * 32; i < i1; i += EPI_WARPS * 32) {
                const int row = i / C4, c4 = i - row * C4;
                const int m = mt * BM + row, n = n0 + c4 * 4;
                if (m >= p.M || n >= p.N) continue;
                const uint32_t off = red_base + (uint32_t)(row * RED_LD + c4 * 4) * 4u;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                for (int s = 0; s < p.splits; s++) {
                    float4 t = ld_dsmem_f4(off, (uint32_t)s);
                    v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
                }
                epi_store4(p, m, n, v);
            }
        }
        cluster_sync_all();                          // no CTA may exit (and free its shared memory) while a peer still reads it
    }
#ifdef RDM_AB_TIMING
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const long long t7 = clock64();
        printf("TCT M=%d N=%d K=%d BN=%d sp=%d cl=%d act=%d grid=%d gt0=%llu | setup %lld pdlwait %lld firstfull %lld lastmma %lld epi0 %lld epiend %lld end %lld\n", p.M, p.N, nkb * BK, BN, p.splits, p.cluster, p.act,
               (int)gridDim.x, gt0, g_ts[1] - g_ts[0], g_ts[2] - g_ts[0], g_ts[3] - g_ts[0], g_ts[4] - g_ts[0], g_ts[5] - g_ts[0], g_ts[6] - g_ts[0], t7 - g_ts[0]);
        printf("TCE chunk0: tmemld %lld transposed %lld stored %lld prefetched %lld (cycles after epi0)\n", g_ts[8] - g_ts[5], g_ts[9] - g_ts[5], g_ts[10] - g_ts[5], g_ts[11] - g_ts[5]);
    }
#endif
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)S::TMEM_COLS) : "memory");
    }
}

// ---- CTA-pair variant (cta_group::2) for the large layers ------------------------------------------------------------------------------------
// The mainloop of the one-CTA kernel is bound by SHARED-MEMORY bandwidth, not by L2 or HBM: a 128 x 192 x 64 k-block writes 40 KB of
// operands into shared memory (TMA) and the tensor core reads the same 40 KB back -- 80 KB through a 128 B/clk port = 640 cycles against
// 384 cycles of math (in-kernel stamps: 587 cycles per k-block; halving the A traffic with halo boxes or prefetching the weights into L2
// changed nothing, commit 9400ff7).  Here the two CTAs of a cluster work on ONE 256-row tile: each CTA holds its 128 rows of A and HALF
// of the B tile (rows [rank * BN / 2, +BN / 2) of the weight block), one `tcgen05.mma.cta_group::2` (M = 256) issued by the even CTA
// drives both tensor cores, each accumulating its 128 rows in its own TMEM.  Per SM and k-block: 28 KB written + 28 KB read instead of
// 40 + 40, and the 28 KB stages leave room for 6 of them.  Protocol (CUTLASS sm100 2-SM pipelines):
//   full[s]       lives in the even CTA: armed by its producer with the bytes of BOTH CTAs; both producers' TMA loads (.cta_group::2)
//                 count on it (peer bit of the barrier address cleared)
//   empty[s]      one per CTA (each producer waits for its own stage); the MMA completion arrives on both (commit ... multicast)
//   tmem_full[b]  one per CTA, same multicast commit; tmem_empty[b] lives in the even CTA and collects the epilogue warps of both
// Epilogue: unchanged (each CTA drains its own 128 TMEM lanes = its own 128 output rows).  No split-K, plain epilogue family only.
constexpr int TC2_STAGES_192 = 6;
template <int BN, int STAGES>
struct Tc2Smem {
    static constexpr int A_BYTES = BM * BK * 2, BH_BYTES = (BN / 2) * BK * 2, STAGE_BYTES = A_BYTES + BH_BYTES;
    static constexpr int EPI_BYTES = EPI_WARPS * 32 * 32 * 4;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 + 256 + EPI_BYTES;
    static constexpr int TMEM_COLS = 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
    static_assert(TOTAL <= 232448, "shared memory budget");
    static_assert((BN / 2) % 8 == 0 && BN % 32 == 0, "the B half tile is whole 8-row swizzle groups");
};
template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcKernelParams p) {
    using S = Tc2Smem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;          // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2] (used in the even CTA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* epi_stage = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int nkb = p.taps * p.kb_per_tap;
    const int ntn = (p.N + BN - 1) / BN, npair = ((p.M + 2 * BM - 1) / (2 * BM)) * ntn;        // work items of a PAIR: (256-row tile, N tile), N fastest
    const int pair0 = blockIdx.x >> 1, pairs = gridDim.x >> 1;
    pdl_launch_dependents();

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmB);
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 2 * EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) {                                  // the same warp of BOTH CTAs allocates (and frees) the pair's tensor memory
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)S::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                               // the peer's barriers exist before anything is signalled across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < EPI_WARP0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == 0 && lane == 0) {
            pdl_wait();
            int it = 0;
            for (int item = pair0; item < npair; item += pairs) {
                const int mt = 2 * (item / ntn) + (int)rank, n0 = (item % ntn) * BN + (int)rank * (BN / 2);
                int x0 = 0, y0 = 0, b0 = 0;
                if (p.plain) x0 = mt * BM;
                else if (p.W > BM) { const int tpr = p.W / BM; x0 = (mt % tpr) * BM; y0 = (mt / tpr) % p.H; b0 = mt / (tpr * p.H); }
                else if (p.bb > 1 || p.bh * p.bw == p.H * p.W) b0 = mt * p.bb;
                else { const int tiles_per_img = (p.H * p.W) / BM; b0 = mt / tiles_per_img; y0 = (mt % tiles_per_img) * p.bh; }
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % STAGES; const uint32_t ph = (it / STAGES) & 1;
                    uint8_t* st = smem + s * S::STAGE_BYTES;
                    mbar_wait(&empty[s], ph ^ 1);
                    if (rank == 0) mbar_expect_tx(&full[s], 2 * S::STAGE_BYTES);
                    const int tap = kb / p.kb_per_tap, kc = (kb - tap * p.kb_per_tap) * BK;
                    const int dy = p.ksize == 3 ? tap / 3 - 1 : 0, dx = p.ksize == 3 ? tap % 3 - 1 : 0;
                    tma_load_4d_2sm(st, &tmA, &full[s], kc, x0 + dx, y0 + dy, b0);           // rows beyond M: out-of-bounds boxes are zero-filled (and still counted)
                    tma_load_2d_2sm(st + S::A_BYTES, &tmB, &full[s], kb * BK, n0);
                }
            }
        } else if (warp == 1 && lane == 0 && rank == 0) {
            const uint32_t idesc = umma_idesc_bf16_2sm(BN, p.f16);
            int it = 0, lt = 0;
            for (int item = pair0; item < npair; item += pairs, lt++) {
                const int buf = lt & 1;
                mbar_wait(&tmem_empty[buf], ((lt >> 1) & 1) ^ 1);        // the epilogues of both CTAs have drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % STAGES; const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a = smem_u32(smem + s * S::STAGE_BYTES), b = a + S::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; k++) umma_bf16_2sm(tmem_d, umma_desc_sw128(a + k * 32), umma_desc_sw128(b + k * 32), idesc, (kb | k) != 0);
                    umma_commit_2sm(&empty[s]);
                }
                umma_commit_2sm(&tmem_full[buf]);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        pdl_wait();
        const int q = warp & 3, half = (warp - EPI_WARP0) >> 2, ew = warp - EPI_WARP0;
        int lt = 0;
        for (int item = pair0; item < npair; item += pairs, lt++) {
            const int buf = lt & 1;
            const int mt = 2 * (item / ntn) + (int)rank, n0 = (item % ntn) * BN;
            const int m_warp0 = mt * BM + q * 32;
            EpiPre pre, pre_nx; XPre xp;
            if (m_warp0 < p.M && n0 + half * 32 + 32 <= p.N) epi_prefetch<EPI>(p, lane, m_warp0, n0 + half * 32, pre);
            mbar_wait(&tmem_full[buf], (lt >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = half; c < BN / 32; c += 2) {
                const int nb = n0 + c * 32;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c * 32);
                if (m_warp0 >= p.M || nb + 32 > p.N) continue;
                const int nbn = nb + 64;
                const bool has_next = c + 2 < BN / 32 && nbn + 32 <= p.N;
                if (has_next) epi_prefetch<EPI>(p, lane, m_warp0, nbn, pre_nx);
                uint32_t r[32];
                tmem_ld32(taddr, r);
                epilogue_chunk<EPI>(p, r, epi_stage + ew * EPI_WARP_FLOATS, lane, m_warp0, nb, pre, xp, nullptr, nullptr, 0, BN, 0);
                if (has_next) pre = pre_nx;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[buf]), 0));      // on the even CTA's barrier
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                               // no CTA of the pair leaves (or frees tensor memory) while the other may still signal it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)S::TMEM_COLS) : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr; cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    }
    return fn;
}

// 4D bf16 tensor (c, x, y, b) with row stride `ld` elements; box (64, bw, bh, bb); 128B swizzle, zero OOB fill
int make_map_4d(CUtensorMap* tm, const void* base, int C, int W, int H, int B, int ld, int bw, int bh, int bb) {
    auto enc = get_encode();
    RDM_REQUIRE(enc, RDM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * W, (cuuint64_t)ld * 2 * W * H};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RDM_REQUIRE(r == CUDA_SUCCESS, RDM_ERR_CUDA, "cuTensorMapEncodeTiled(4d: C=%d W=%d H=%d B=%d ld=%d box %d,%d,%d) failed: %d", C, W, H, B, ld, bw, bh, bb, (int)r);
    return RDM_OK;
}
int make_map_2d(CUtensorMap* tm, const void* base, int K, int rows, int ld, int box_rows) {
    auto enc = get_encode();
    RDM_REQUIRE(enc, RDM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RDM_REQUIRE(r == CUDA_SUCCESS, RDM_ERR_CUDA, "cuTensorMapEncodeTiled(2d: K=%d rows=%d ld=%d box %d) failed: %d", K, rows, ld, box_rows, (int)r);
    return RDM_OK;
}

template <int BN, int NSPLIT, int STAGES, int EPI>
int launch_tc(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo, const TcKernelParams& p, cudaStream_t st, int* cluster_cap) {
    auto kern = gemm_tc_kernel<BN, NSPLIT, STAGES, EPI>;
    constexpr int smem = TcSmem<BN, NSPLIT, STAGES>::TOTAL;
    static bool configured[16] = {false};
    int dev = 0; cudaGetDevice(&dev);
    if (!configured[dev & 15]) {
        RDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured[dev & 15] = true;
    }
    if (cluster_cap) {                               // query only: how many clusters of p.splits CTAs can be co-resident
        static int cap[16][9];
        int& c = cap[dev & 15][p.splits];
        if (c == 0) {
            cudaLaunchConfig_t q{};
            q.gridDim = dim3(p.splits * 64); q.blockDim = dim3(TC_THREADS); q.dynamicSmemBytes = smem;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension; qa[0].val.clusterDim.x = p.splits; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            q.attrs = qa; q.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess) { cudaGetLastError(); n = 0; }
            c = n > 0 ? n : -1;
        }
        *cluster_cap = c > 0 ? c : 0;
        return RDM_OK;
    }
    const int nitems = ((p.N + BN - 1) / BN) * ((p.M + BM - 1) / BM) * p.splits * (p.ups ? 4 : 1);
    const int sms = rdm_num_sms(dev);
    const int use_pdl = g_rdm_use_pdl;
    TcKernelParams pl = p; pl.pdl = use_pdl && !p.pdl_off;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.cluster ? nitems : (nitems < sms ? nitems : sms)); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (use_pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; na++; }
    if (p.cluster) { attr[na].id = cudaLaunchAttributeClusterDimension; attr[na].val.clusterDim.x = p.splits; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1; na++; }
    cfg.attrs = attr; cfg.numAttrs = na;
    RDM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a_hi, a_lo, b_hi, b_lo, pl));
    RDM_COUNT_LAUNCH();
    RDM_CHECK_CUDA(cudaGetLastError());
    return RDM_OK;
}

template <int BN, int STAGES, int EPI>
int launch_tc2(const CUtensorMap& ta, const CUtensorMap& tb, const TcKernelParams& p, cudaStream_t st) {
    auto kern = gemm_tc2_kernel<BN, STAGES, EPI>;
    constexpr int smem = Tc2Smem<BN, STAGES>::TOTAL;
    static bool configured[16] = {false};
    int dev = 0; cudaGetDevice(&dev);
    if (!configured[dev & 15]) {
        RDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured[dev & 15] = true;
    }
    const int npair = ((p.M + 2 * BM - 1) / (2 * BM)) * ((p.N + BN - 1) / BN), max_pairs = rdm_num_sms(dev) / 2;
    const int pairs = npair < max_pairs ? npair : max_pairs;
    TcKernelParams pl = p; pl.pdl = 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (g_rdm_use_pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; na++; }
    attr[na].id = cudaLaunchAttributeClusterDimension; attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1; na++;
    cfg.attrs = attr; cfg.numAttrs = na;
    RDM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, pl));
    RDM_COUNT_LAUNCH();
    RDM_CHECK_CUDA(cudaGetLastError());
    return RDM_OK;
}

// dispatch on (tile width, MMAs per product, epilogue family); cluster_cap != nullptr: capacity query only (see launch_tc)
template <int EPI>
int dispatch_tc_epi(int BN, int nsplit, const CUtensorMap& ta_hi, const CUtensorMap& ta_lo, const CUtensorMap& tb_hi, const CUtensorMap& tb_lo,
                    const TcKernelParams& p, cudaStream_t st, int* cluster_cap) {
    if (nsplit == 2) {
        if (BN == 192) return launch_tc<192, 2, 3, EPI>(ta_hi, ta_lo, tb_hi, tb_lo, p, st, cluster_cap);
        if (BN == 128) return launch_tc<128, 2, 4, EPI>(ta_hi, ta_lo, tb_hi, tb_lo, p, st, cluster_cap);
        if (BN == 64) return launch_tc<64, 2, 6, EPI>(ta_hi, ta_lo, tb_hi, tb_lo, p, st, cluster_cap);
        return launch_tc<32, 2, 8, EPI>(ta_hi, ta_lo, tb_hi, tb_lo, p, st, cluster_cap);
    } else if (nsplit == 3) {
        if (BN == 192) return launch_tc<192, 3, 2, EPI>(ta_hi, ta_lo, tb_hi, tb_lo, p, st, cluster_cap);
        if (BN == 128) return launch_tc<128, 3, 3, EPI>(ta_hi, ta_lo, tb_hi, tb_lo, p, st, cluster_cap);
        if (BN == 64) return launch_tc<64, 3, 4, EPI>(ta_hi, ta_lo, tb_hi, tb_lo, p, st, cluster_cap);
        return launch_tc<32, 3, 4, EPI>(ta_hi, ta_lo, tb_hi, tb_lo, p, st, cluster_cap);
    }
    if (BN == 192) return launch_tc<192, 1, 4, EPI>(ta_hi, ta_lo, tb_hi, tb_lo, p, st, cluster_cap);
    if (BN == 128) return launch_tc<128, 1, 6, EPI>(ta_hi, ta_lo, tb_hi, tb_lo, p, st, cluster_cap);
    if (BN == 64) return launch_tc<64, 1, 8, EPI>(ta_hi, ta_lo, tb_hi, tb_lo, p, st, cluster_cap);
    return launch_tc<32, 1, 8, EPI>(ta_hi, ta_lo, tb_hi, tb_lo, p, st, cluster_cap);
}
int dispatch_tc(int BN, int nsplit, const CUtensorMap& ta_hi, const CUtensorMap& ta_lo, const CUtensorMap& tb_hi, const CUtensorMap& tb_lo,
                const TcKernelParams& p, cudaStream_t st, int* cluster_cap) {
    static const int one_family = getenv("RDM_TC_ONE_EPI") ? 1 : 0;              // A/B: always the all-in-one kernel
    int epi = EPI_ANY;
    if (!one_family) {
        if (p.act == ACT_XATTN) epi = EPI_XATTN;
        else if (p.act == ACT_GEGLU && (p.N & 31) == 0) epi = EPI_GEGLU;
        else if (p.act == ACT_NONE && (p.N & 31) == 0 && !(p.rowvec && p.rows_per_batch < 16)) epi = EPI_PLAIN;
    }
    if (epi == EPI_PLAIN) return dispatch_tc_epi<EPI_PLAIN>(BN, nsplit, ta_hi, ta_lo, tb_hi, tb_lo, p, st, cluster_cap);
    if (epi == EPI_GEGLU) return dispatch_tc_epi<EPI_GEGLU>(BN, nsplit, ta_hi, ta_lo, tb_hi, tb_lo, p, st, cluster_cap);
    if (epi == EPI_XATTN) return dispatch_tc_epi<EPI_XATTN>(BN, nsplit, ta_hi, ta_lo, tb_hi, tb_lo, p, st, cluster_cap);
    return dispatch_tc_epi<EPI_ANY>(BN, nsplit, ta_hi, ta_lo, tb_hi, tb_lo, p, st, cluster_cap);
}

}  // namespace

static int g_tc_ws_slot = 0;
void gemm_tc_set_workspace_slot(int slot) { g_tc_ws_slot = slot; }

bool gemm_tc_supported(const TcA& a) {
    if (a.C % BK != 0 || a.ld % 8 != 0) return false;
    if (a.ksize == 1) return true;
    // 3x3: a 128-row tile must be an (x, y, b) box of the image
    const int W = a.W, H = a.H;
    if (W > BM) return W % BM == 0;                    // first-stage decoder (128 / 256 pixel rows): tiles are row pieces
    if ((BM % W) != 0) return false;
    const int rows_y = BM / W;                        // image rows per tile if H is large enough
    if (H >= rows_y) return H % rows_y == 0;
    return (BM % (W * H)) == 0;
}

int gemm_tc(const TcA& a, const TcW& w, const GemmEpi& e, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int out_bf_ld, int nsplit, int f16, cudaStream_t st) {
    RDM_REQUIRE(gemm_tc_supported(a), RDM_ERR_UNSUPPORTED, "gemm_tc: shape not supported (C=%d W=%d H=%d ks=%d)", a.C, a.W, a.H, a.ksize);
    RDM_REQUIRE(nsplit >= 1 && nsplit <= 3 && a.hi && w.hi && (nsplit < 2 || w.lo) && (nsplit < 3 || a.lo), RDM_ERR_ARG, "gemm_tc: missing operand plane (nsplit %d)", nsplit);
    const int ntaps = a.ups ? 4 : a.ksize * a.ksize;
    RDM_REQUIRE(w.K == ntaps * a.C && w.ld % 8 == 0, RDM_ERR_ARG, "gemm_tc: weight K=%d vs %d", w.K, ntaps * a.C);
    const bool ext = a.hi2 != nullptr;
    RDM_REQUIRE(!a.ups || (a.ksize == 3 && a.W <= BM && !ext && e.act == ACT_NONE && !e.res && !e.rowvec && (w.N & 31) == 0), RDM_ERR_ARG,
                "gemm_tc: upsample-folded conv needs a plain epilogue and N %% 32 == 0 (N=%d W=%d)", w.N, a.W);
    RDM_REQUIRE(!ext || (nsplit == 1 && w.hi2 && a.C2 > 0 && a.C2 % BK == 0 && a.ld2 % 8 == 0 && w.ld2 % 8 == 0), RDM_ERR_ARG, "gemm_tc: K extension needs one-plane operands and C2 %% 64 == 0 (C2=%d nsplit=%d)", a.C2, nsplit);
    TcKernelParams p{};
    const int M = a.B * a.H * a.W;
    p.M = M; p.N = w.N; p.taps = ntaps; p.kb_per_tap = a.C / BK; p.ksize = a.ups ? 2 : a.ksize; p.ups = a.ups ? 1 : 0; p.H = a.H; p.W = a.W;
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    if (a.ksize == 1) {
        // plain [M, C] matrix: box = 128 rows
        p.plain = 1; p.bw = BM; p.bh = 1; p.bb = 1;
        // encode as (c, m, 1, 1): x = m
        RDM_TRY(make_map_4d(&ta_hi, a.hi, a.C, M, 1, 1, a.ld, BM, 1, 1));
        if (nsplit == 3) RDM_TRY(make_map_4d(&ta_lo, a.lo, a.C, M, 1, 1, a.ld, BM, 1, 1));
        else if (ext) RDM_TRY(make_map_4d(&ta_lo, a.hi2, a.C2, M, 1, 1, a.ld2, BM, 1, 1));
        else ta_lo = ta_hi;
    } else {
        const int W = a.W, H = a.H;
        if (W > BM) { p.bw = BM; p.bh = 1; p.bb = 1; }
        else { p.bw = W; p.bh = (BM / W) < H ? (BM / W) : H; p.bb = BM / (p.bw * p.bh); }
        RDM_TRY(make_map_4d(&ta_hi, a.hi, a.C, W, H, a.B, a.ld, p.bw, p.bh, p.bb));
        if (nsplit == 3) RDM_TRY(make_map_4d(&ta_lo, a.lo, a.C, W, H, a.B, a.ld, p.bw, p.bh, p.bb));
        else if (ext) RDM_TRY(make_map_4d(&ta_lo, a.hi2, a.C2, W, H, a.B, a.ld2, p.bw, p.bh, p.bb));      // same boxes, centre tap only
        else ta_lo = ta_hi;
    }
    p.bias = e.bias; p.rowvec = e.rowvec; p.rowvec_ld = e.rowvec_ld; p.rows_per_batch = e.rows_per_batch > 0 ? e.rows_per_batch : 1;
    p.res = e.res; p.res_ld = e.res_ld; p.act = e.act; p.out = e.out; p.out_ld = e.out_ld;
    p.xkv = e.xkv; p.xkv_ld = e.xkv_ld; p.xv_off = e.xv_off; p.xk = e.xk; p.xscale_log2e = e.xscale * 1.4426950408889634f;
    RDM_REQUIRE(e.act != ACT_XATTN || (e.xkv && e.xk >= 1 && e.xk <= 8 && p.rows_per_batch >= 16 && p.rows_per_batch % 16 == 0 && w.N % 32 == 0 && e.xkv_ld % 4 == 0 && !e.bias && !e.res && !e.rowvec), RDM_ERR_ARG,
                "gemm_tc: fused cross-attention needs 1..8 context rows, N %% 32 == 0 and no bias/residual (xk=%d N=%d)", e.xk, w.N);
    p.out_hi = out_hi; p.out_lo = out_lo; p.out_bf_ld = out_bf_ld; p.f16 = f16; p.pdl_off = w.dynamic;
    static const int no_geglu_rows = getenv("RDM_TC_GEGLU_TRANSPOSED") ? 1 : 0;       // A/B: the transposed GEGLU epilogue
    p.geglu_rows = e.act == ACT_GEGLU && !e.res && !e.rowvec && !no_geglu_rows && (w.N & 63) == 0 &&
                   (e.out ? e.out_ld % 4 == 0 && ((uintptr_t)e.out & 15) == 0 : out_bf_ld % 8 == 0 && ((uintptr_t)out_hi & 15) == 0 && (!out_lo || ((uintptr_t)out_lo & 15) == 0));
    RDM_REQUIRE((p.out != nullptr) != (p.out_hi != nullptr), RDM_ERR_ARG, "gemm_tc: exactly one of fp32 / bf16 outputs");
    // CTA pairs (gemm_tc2_kernel) for the large plain layers: 256-row tiles, no split-K.  RDM_TC_2SM=0 switches them off.
    {
        static const int use_2sm = getenv("RDM_TC_2SM") ? atoi(getenv("RDM_TC_2SM")) : 1;
        static const int min_m = getenv("RDM_TC_2SM_MIN_M") ? atoi(getenv("RDM_TC_2SM_MIN_M")) : 8192;
        const int bn2 = w.N % 192 == 0 ? 192 : w.N % 128 == 0 ? 128 : 0;
        // Measured per layer (B200, fp16): the pair kernel wins where the one-CTA kernel needs more than one wave of tiles AND several N
        // tiles (M = 32768, N = 384: 964 vs 598 TFLOP/s; M = 8192, N = 576: 713 vs 612) and loses a little on single-wave layers and on
        // short K (cluster launch + two cluster barriers around a 2 us mainloop: M = 8192, N = 384, K = 384: 150 vs 172).  The mainloop itself
        // gains nothing -- it already runs at the chip's burst tensor rate in both kernels (587 cycles per 128 x 192 x 64 k-block =
        // 1560 TFLOP/s over 148 SMs); RDM_TC_2SM=2 forces the pair kernel wherever it is legal.
        const int tiles1 = ((M + BM - 1) / BM) * ((w.N + 191) / 192);
        // Short-K layers with MANY waves of tiles (the GEGLU projection: M = 8192, N = 3072, K = 384 -- 1024 tiles of 6 k-blocks, 415 TFLOP/s)
        // looked bound by the operand traffic per tile ((128 + 192) x K x 2 bytes from L2 per 128 x 192 outputs; a pair shares the weight
        // tile: 1.43x fewer bytes).  Measured: 4.176 ms per forward with pairs for them, 4.156 without -- they are bound by the GEGLU
        // EPILOGUE (TMEM -> GELU -> fp16 stores of a tile take longer than its 6 k-blocks), which a pair does not shorten.  Opt-in:
        // RDM_TC_2SM_WAVES = waves of one-CTA tiles from which short K qualifies (0, the default: never).
        static const int short_k_waves = getenv("RDM_TC_2SM_WAVES") ? atoi(getenv("RDM_TC_2SM_WAVES")) : 0;
        const bool geglu = e.act == ACT_GEGLU && p.geglu_rows;
        const bool wins = use_2sm >= 2 || (tiles1 > 148 && w.N > 192 && (p.taps * p.kb_per_tap >= 16 || (short_k_waves > 0 && tiles1 >= short_k_waves * 148)));
        if (use_2sm && wins && !ext && !a.ups && nsplit == 1 && bn2 && M >= min_m && M % (2 * BM) == 0 && (e.act == ACT_NONE || geglu) && !e.stats && !(e.rowvec && p.rows_per_batch < 16)) {
            p.splits = 1; p.kb_per_split = p.taps * p.kb_per_tap; p.part = nullptr; p.cluster = 0; p.stats = nullptr; if (!geglu) p.geglu_rows = 0;
            RDM_TRY(make_map_2d(&tb_hi, w.hi, w.K, w.N, w.ld, bn2 / 2));
            if (e.stats_fused) *e.stats_fused = 0;
            if (getenv("RDM_TC_TRACE")) fprintf(stderr, "gemm_tc M=%d N=%d K=%d -> CTA pairs, BN=%d\n", M, w.N, w.K, bn2);
            if (geglu) return bn2 == 192 ? launch_tc2<192, 6, EPI_GEGLU>(ta_hi, tb_hi, p, st) : launch_tc2<128, 7, EPI_GEGLU>(ta_hi, tb_hi, p, st);
            return bn2 == 192 ? launch_tc2<192, 6, EPI_PLAIN>(ta_hi, tb_hi, p, st) : launch_tc2<128, 7, EPI_PLAIN>(ta_hi, tb_hi, p, st);
        }
    }
    // Tile width and split-K by a cost model fitted to measured layer times: one k-block of a work item costs ~(128 + BN) (MMA time
    // grows with BN, the 128-row A tile is re-loaded for every N tile); every item pays a fixed prologue/epilogue worth ~6 k-blocks.
    // Small-M layers (4x4 / 8x8 latents) split K so that all SMs stream a slice of the weights.
    // cost = waves * (kb_per_item + 6) * (128 + BN) (+ the reduction).
    p.kb2 = ext ? a.C2 / BK : 0;
    const int mtiles = (M + BM - 1) / BM, sms = 148, nkb_total = p.taps * p.kb_per_tap + p.kb2;
    static const int forced = getenv("RDM_TC_BN") ? atoi(getenv("RDM_TC_BN")) : 0;
    static const int no_split = getenv("RDM_TC_NOSPLIT") ? 1 : 0;
    static const int use_cluster = getenv("RDM_TC_CLUSTER") ? atoi(getenv("RDM_TC_CLUSTER")) : 0;
    static const int c_fix = getenv("RDM_TC_FIX") ? atoi(getenv("RDM_TC_FIX")) : 6, c_red = getenv("RDM_TC_RED") ? atoi(getenv("RDM_TC_RED")) : 5,
                     c_minkb = getenv("RDM_TC_MINKB") ? atoi(getenv("RDM_TC_MINKB")) : 4;       // cost-model constants (developer sweeps)
    // Split-K has two implementations: (a) persistent CTAs write fp32 partials, a dependent kernel reduces them; (b) the splits of a tile
    // form a thread-block cluster (CTA rank = split) and reduce through distributed shared memory -- no partials in HBM/L2 and no second
    // kernel, but splits <= 8 (portable cluster size), one CTA per item, and only `cap` clusters are co-resident (GPC granularity).
    // RDM_TC_CLUSTER = 0 (default since round 2): reduce kernel only; 1: both are candidates of the cost model below; 2: cluster variant only.
    // Measured on B200 (full architecture, B2 = 32, fp16 mode, graph replay): 4.46 ms per forward with 0, 4.64 ms with 1, 4.80 ms with 2 --
    // the cluster variant pays two cluster barriers, the DSMEM pass and the gang-scheduled launch on the critical path of every small GEMM
    // (in-kernel stamps: 5 us from the last MMA to the end of the kernel against ~2 us for the straight epilogue).
    CUtensorMap tdummy = ta_hi;
    int BN = 32, splits = 1, clustered = 0;
    const int cand[4] = {192, 128, 64, 32};
    long best = -1;
    for (int c : cand) {
        if (c > 32 && w.N <= c / 2) continue;
        if (forced && c != forced) continue;
        const int nt = (w.N + c - 1) / c;
        for (int cl = 0; cl <= 1; cl++) {
            if (cl == 1 && use_cluster == 0) continue;
            for (int sp = cl ? 2 : 1; sp <= (cl ? 8 : 16); sp++) {
                if (sp > 1 && (no_split || a.ups || e.act == ACT_XATTN || nkb_total / sp < c_minkb || (w.N & 31) || (!cl && (size_t)sp * M * w.N * 4 > ((size_t)48 << 20)))) break;
                if (sp > 1 && !cl && use_cluster == 2) break;
                const int kbps = (nkb_total + sp - 1) / sp;
                if ((sp - 1) * kbps >= nkb_total) continue;                  // an empty split
                long items = (long)mtiles * nt * sp * (a.ups ? 4 : 1), waves = (items + sms - 1) / sms;
                if (cl) {
                    TcKernelParams q = p; q.splits = sp; int cap = 0;
                    RDM_TRY(dispatch_tc(c, nsplit, tdummy, tdummy, tdummy, tdummy, q, st, &cap));
                    if (cap <= 0) continue;
                    waves = ((long)mtiles * nt + cap - 1) / cap;
                }
                // fixed cost of the reduction in k-block units: a dependent reduce kernel over partials in L2 vs an in-kernel DSMEM pass
                const long cost = waves * (kbps + c_fix) * (128 + c) + (sp > 1 ? (cl ? 2 : c_red) * (128 + c) : 0);
                if (best < 0 || cost < best) { best = cost; BN = c; splits = sp; clustered = cl; }
            }
        }
    }
    p.splits = splits; p.kb_per_split = (nkb_total + splits - 1) / splits; p.part = nullptr; p.cluster = clustered;
    // GroupNorm statistics in the epilogue: only on the straight path (no split-K: those tiles are finished by epi_store4), plain epilogue,
    // fp32 result, whole 32-column chunks, images of a multiple of 128 rows (a tile then lies in one image: the 32x32 and 16x16 levels)
    const bool stats_ok = e.stats && !a.ups && splits == 1 && e.act == ACT_NONE && p.out && (w.N & 31) == 0 && !(e.rowvec && p.rows_per_batch < 16) &&
                          e.stats_hw >= BM && e.stats_hw % BM == 0 && M % e.stats_hw == 0;
    p.stats = stats_ok ? e.stats : nullptr; p.stats_ld = e.stats_ld; p.stats_hw = e.stats_hw;
    if (e.stats_fused) *e.stats_fused = stats_ok ? 1 : 0;
    if (getenv("RDM_TC_TRACE")) fprintf(stderr, "gemm_tc M=%d N=%d K=%d -> BN=%d splits=%d cluster=%d\n", M, w.N, w.K, BN, splits, clustered);
    if (splits > 1 && !p.cluster) {
        static float* ws[16][8] = {{nullptr}}; static size_t ws_cap[16][8] = {{0}};      // [device][workspace slot]: concurrent chains must not share partials
        int dev = 0; cudaGetDevice(&dev);
        const int slot = g_tc_ws_slot & 7;
        const size_t need = (size_t)splits * M * w.N * sizeof(float);
        if (need > ws_cap[dev & 15][slot]) {                             // grown during the eager warm-up pass, never inside a graph capture
            if (ws[dev & 15][slot]) cudaFree(ws[dev & 15][slot]);
            ws[dev & 15][slot] = nullptr; ws_cap[dev & 15][slot] = 0;
            size_t cap = (size_t)48 << 20;
            RDM_CHECK_CUDA(cudaMalloc((void**)&ws[dev & 15][slot], cap));
            ws_cap[dev & 15][slot] = cap;
        }
        p.part = ws[dev & 15][slot];
    }
    const int w_rows = a.ups ? 4 * w.N : w.N;
    RDM_TRY(make_map_2d(&tb_hi, w.hi, w.K, w_rows, w.ld, BN));
    if (nsplit >= 2) RDM_TRY(make_map_2d(&tb_lo, w.lo, w.K, w_rows, w.ld, BN));
    else if (ext) RDM_TRY(make_map_2d(&tb_lo, w.hi2, a.C2, w.N, w.ld2, BN));
    else tb_lo = tb_hi;
    RDM_TRY(dispatch_tc(BN, nsplit, ta_hi, ta_lo, tb_hi, tb_lo, p, st, nullptr));
    if (splits > 1 && !p.cluster) {
        const long long n4 = (long long)M * (w.N >> 2);
        RDM_CHECK_CUDA(launch_pdl(splitk_reduce_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, st, p));
        RDM_COUNT_LAUNCH();
        RDM_CHECK_CUDA(cudaGetLastError());
    }
    return RDM_OK;
}
