// fp32 CUDA-core implicit GEMM: out[M,N] = epi(A[M,K] * W[N,K]^T).  The "strict" engine: every product and sum
// is fp32 (parity <= 1e-4 vs the fp32 oracle) and it handles every conv flavour of the U-Net by index arithmetic
// in the A-tile loader: 3x3/pad 1, 1x1, stride 2 (ldm Downsample), nearest-2x upsample folded into the gather
// (ldm Upsample), and plain Linear layers (ksize 1).  128x128x16 tiles, 256 threads, 8x8 micro-tiles,
// register-staged double buffering.  The tensor-core engine (gemm_tc.cu) replaces it for speed.
#include "kernels.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256, PAD = 4;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }

struct AParams {
    const float* x; int ld;
    int B, Hs, Ws, Cin, Ho, Wo, ksize, stride, ups;
    int M, K;
};
struct EParams {
    const float* bias; const float* rowvec; int rowvec_ld; int rows_per_batch;
    const float* res; int res_ld; int act; float* out; int out_ld;
};

// VEC: Cin % BK == 0 (a BK slab never straddles a tap) and all bases 16-byte aligned -> float4 loads.
template <bool VEC>
__global__ void __launch_bounds__(THREADS)
gemm_simt_kernel(AParams a, const float* __restrict__ W, int N, EParams e) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int lrow = tid >> 2, lk = (tid & 3) * 4;          // loader: rows lrow, lrow+64; k offset lk..lk+3
    const int pad = a.ksize == 3 ? 1 : 0;
    const int Hi = a.ups ? a.Hs * 2 : a.Hs, Wi = a.ups ? a.Ws * 2 : a.Ws;     // logical input grid of the conv

    int rb[2], roy[2], rox[2]; bool rvalid[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        int m = m0 + lrow + i * 64;
        rvalid[i] = m < a.M;
        int mm = rvalid[i] ? m : 0;
        rox[i] = mm % a.Wo; int t = mm / a.Wo; roy[i] = t % a.Ho; rb[i] = t / a.Ho;
    }
    const int nK = (a.K + BK - 1) / BK;
    float4 ra[2], rw[2];

    auto load_tiles = [&](int kt) {
        const int k0 = kt * BK + lk;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (VEC) {
                int tap = k0 / a.Cin, c = k0 - tap * a.Cin;
                int dy = a.ksize == 3 ? tap / 3 : 0, dx = a.ksize == 3 ? tap % 3 : 0;
                int iy = roy[i] * a.stride + dy - pad, ix = rox[i] * a.stride + dx - pad;
                if (rvalid[i] && k0 < a.K && iy >= 0 && iy < Hi && ix >= 0 && ix < Wi) {
                    int sy = a.ups ? iy >> 1 : iy, sx = a.ups ? ix >> 1 : ix;
                    const float4 q = *reinterpret_cast<const float4*>(a.x + ((size_t)(rb[i] * a.Hs + sy) * a.Ws + sx) * a.ld + c);
                    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    int k = k0 + j;
                    if (rvalid[i] && k < a.K) {
                        int tap = k / a.Cin, c = k - tap * a.Cin;
                        int dy = a.ksize == 3 ? tap / 3 : 0, dx = a.ksize == 3 ? tap % 3 : 0;
                        int iy = roy[i] * a.stride + dy - pad, ix = rox[i] * a.stride + dx - pad;
                        if (iy >= 0 && iy < Hi && ix >= 0 && ix < Wi) {
                            int sy = a.ups ? iy >> 1 : iy, sx = a.ups ? ix >> 1 : ix;
                            v[j] = a.x[((size_t)(rb[i] * a.Hs + sy) * a.Ws + sx) * a.ld + c];
                        }
                    }
                }
            }
            ra[i] = make_float4(v[0], v[1], v[2], v[3]);
            // weights
            int n = n0 + lrow + i * 64;
            float w[4] = {0.f, 0.f, 0.f, 0.f};
            if (n < N) {
                if (VEC) {
                    if (k0 < a.K) { const float4 q = *reinterpret_cast<const float4*>(W + (size_t)n * a.K + k0); w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w; }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) if (k0 + j < a.K) w[j] = W[(size_t)n * a.K + k0 + j];
                }
            }
            rw[i] = make_float4(w[0], w[1], w[2], w[3]);
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
            int r = lrow + i * 64;
            As[buf][lk + 0][r] = ra[i].x; As[buf][lk + 1][r] = ra[i].y; As[buf][lk + 2][r] = ra[i].z; As[buf][lk + 3][r] = ra[i].w;
            Bs[buf][lk + 0][r] = rw[i].x; Bs[buf][lk + 1][r] = rw[i].y; Bs[buf][lk + 2][r] = rw[i].z; Bs[buf][lk + 3][r] = rw[i].w;
        }
    };

    const int ty = tid >> 4, tx = tid & 15;     // rows ty*4(+64), cols tx*4(+64)
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nK; kt++) {
        const int buf = kt & 1;
        if (kt + 1 < nK) load_tiles(kt + 1);
#pragma unroll
        for (int k = 0; k < BK; k++) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nK) { store_tiles(buf ^ 1); __syncthreads(); }
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= a.M) continue;
        const int bidx = m / e.rows_per_batch;
#pragma unroll
        for (int jh = 0; jh < 2; jh++) {
            const int nb = n0 + jh * 64 + tx * 4;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int n = nb + j;
                float t = acc[i][jh * 4 + j];
                if (n < N) {
                    if (e.bias) t += e.bias[n];
                    if (e.rowvec) t += e.rowvec[(size_t)bidx * e.rowvec_ld + n];
                }
                v[j] = t;
            }
            if (e.act == ACT_GEGLU) {
                // columns (2j, 2j+1) = (value, gate) -> output column j  (weights interleaved at load time)
#pragma unroll
                for (int j = 0; j < 4; j += 2) {
                    int n = nb + j;
                    if (n + 1 < N) {
                        int no = n >> 1;
                        float t = v[j] * gelu_erf(v[j + 1]);
                        if (e.res) t += e.res[(size_t)m * e.res_ld + no];
                        e.out[(size_t)m * e.out_ld + no] = t;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    int n = nb + j;
                    if (n < N) {
                        float t = v[j];
                        if (e.act == ACT_SILU) t = silu_f(t);
                        else if (e.act == ACT_QUICKGELU) t = t / (1.f + expf(-1.702f * t));
                        if (e.res) t += e.res[(size_t)m * e.res_ld + n];
                        e.out[(size_t)m * e.out_ld + n] = t;
                    }
                }
            }
        }
    }
}

}  // namespace

int gemm_simt(const GemmA& a, const float* W, int N, const GemmEpi& e, cudaStream_t st) {
    RDM_REQUIRE(a.x && W && e.out, RDM_ERR_ARG, "gemm_simt: null operand");
    RDM_REQUIRE(a.ksize == 1 || a.ksize == 3, RDM_ERR_UNSUPPORTED, "gemm_simt: ksize %d", a.ksize);
    AParams ap{a.x, a.ld, a.B, a.Hs, a.Ws, a.Cin, a.Ho, a.Wo, a.ksize, a.stride, a.ups, a.M(), a.K()};
    EParams ep{e.bias, e.rowvec, e.rowvec_ld, e.rows_per_batch > 0 ? e.rows_per_batch : 1, e.res, e.res_ld, e.act, e.out, e.out_ld};
    dim3 grid((N + BN - 1) / BN, (ap.M + BM - 1) / BM);
    bool vec = (a.Cin % BK == 0) && (a.ld % 4 == 0) && (((uintptr_t)a.x & 15) == 0) && (((uintptr_t)W & 15) == 0);
    if (vec) gemm_simt_kernel<true><<<grid, THREADS, 0, st>>>(ap, W, N, ep);
    else gemm_simt_kernel<false><<<grid, THREADS, 0, st>>>(ap, W, N, ep);
    RDM_COUNT_LAUNCH();
    RDM_CHECK_CUDA(cudaGetLastError());
    return RDM_OK;
}
