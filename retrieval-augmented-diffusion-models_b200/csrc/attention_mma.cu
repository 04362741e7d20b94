// Tensor-core self-attention for heads of width 32 (CrossAttention.forward with context=None, rdm/modules/attention.py:42-74):
//   out = softmax(q k^T * d^-0.5) v  per (batch, head), q/k/v = column blocks of the fp16 QKV plane written by the QKV GEMM.
// The whole K and V of one (batch, head) fit in shared memory (256 keys x 32 x 2 B = 16 KB each), so a CTA stages them once
// and each warp owns 16 query rows: S = Q K^T and O += P V run on warp-level fp16 MMAs (m16n8k16, fp32 accumulate) with the
// online softmax in fp32 registers; the S accumulator fragment is re-used directly as the A fragment of the P V product, so
// no score tensor ever leaves the register file.  (0.2 % of the U-Net FLOPs: this op is latency/traffic bound, which is why it
// is a register-resident warp-MMA kernel and not a TMEM pipeline.)
#include "kernels.cuh"
#include "ptx.cuh"
#include <math_constants.h>

namespace {

constexpr int KPAD = 40;          // K row stride in halves (80 B): conflict-free B-fragment loads
constexpr int VPAD = 8;           // V^T row stride = Nk + 8 halves

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void store2(const Out4& o, size_t m, int c, float a, float b) {
    if (o.f) *reinterpret_cast<float2*>(o.f + m * o.ldf + c) = make_float2(a, b);
    if (o.hi) {
        unsigned short h0, l0, h1, l1;
        split16(a, o.f16, h0, l0); split16(b, o.f16, h1, l1);
        *reinterpret_cast<uint32_t*>(o.hi + m * o.ldb + c) = (uint32_t)h0 | ((uint32_t)h1 << 16);
        if (o.lo) *reinterpret_cast<uint32_t*>(o.lo + m * o.ldb + c) = (uint32_t)l0 | ((uint32_t)l1 << 16);
    }
}

// grid (Nq / QT, heads, B), block QT * 2 threads (one warp per 16 query rows).  Nq % QT == 0, Nk % 64 == 0.
template <int QT>
__global__ void __launch_bounds__(QT * 2)
attention_mma_kernel(const __half* __restrict__ q, const __half* __restrict__ k, const __half* __restrict__ v, int ld,
                     int Nq, int Nk, float scale_log2e, Out4 out) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(16) unsigned char smem_att[];
    __half* sK = reinterpret_cast<__half*>(smem_att);                 // [Nk][KPAD]
    __half* sVt = sK + (size_t)Nk * KPAD;                             // [32][Nk + VPAD]
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int vld = Nk + VPAD;
    {
        const __half* kb = k + (size_t)b * Nk * ld + h * 32;
        const __half* vb = v + (size_t)b * Nk * ld + h * 32;
        for (int i = tid; i < Nk * 4; i += QT * 2) {                  // 4 x 16-byte pieces per 64-byte row
            const int j = i >> 2, c = i & 3;
            *reinterpret_cast<uint4*>(sK + j * KPAD + c * 8) = *reinterpret_cast<const uint4*>(kb + (size_t)j * ld + c * 8);
            const uint4 vv = *reinterpret_cast<const uint4*>(vb + (size_t)j * ld + c * 8);
            const __half* vh = reinterpret_cast<const __half*>(&vv);
#pragma unroll
            for (int e = 0; e < 8; e++) sVt[(c * 8 + e) * vld + j] = vh[e];
        }
    }
    // Q fragments of this warp's 16 rows (A operand, row-major): straight from global memory
    const size_t r_lo = (size_t)b * Nq + q0 + warp * 16 + g, r_hi = r_lo + 8;
    uint32_t qa[2][4];
#pragma unroll
    for (int kk = 0; kk < 2; kk++) {
        const int c = h * 32 + kk * 16 + 2 * t;
        qa[kk][0] = *reinterpret_cast<const uint32_t*>(q + r_lo * ld + c);
        qa[kk][1] = *reinterpret_cast<const uint32_t*>(q + r_hi * ld + c);
        qa[kk][2] = *reinterpret_cast<const uint32_t*>(q + r_lo * ld + c + 8);
        qa[kk][3] = *reinterpret_cast<const uint32_t*>(q + r_hi * ld + c + 8);
    }
    __syncthreads();

    float o[4][4];
#pragma unroll
    for (int nb = 0; nb < 4; nb++) { o[nb][0] = 0.f; o[nb][1] = 0.f; o[nb][2] = 0.f; o[nb][3] = 0.f; }
    float m_lo = -CUDART_INF_F, m_hi = -CUDART_INF_F, l_lo = 0.f, l_hi = 0.f;
    for (int kc = 0; kc < Nk; kc += 64) {
        float s[8][4];
#pragma unroll
        for (int nb = 0; nb < 8; nb++) {
            s[nb][0] = 0.f; s[nb][1] = 0.f; s[nb][2] = 0.f; s[nb][3] = 0.f;
            const __half* kr = sK + (kc + nb * 8 + g) * KPAD + 2 * t;
#pragma unroll
            for (int kk = 0; kk < 2; kk++)
                mma_f16(s[nb], qa[kk], *reinterpret_cast<const uint32_t*>(kr + kk * 16), *reinterpret_cast<const uint32_t*>(kr + kk * 16 + 8));
        }
        float mx_lo = -CUDART_INF_F, mx_hi = -CUDART_INF_F;
#pragma unroll
        for (int nb = 0; nb < 8; nb++) {
            s[nb][0] *= scale_log2e; s[nb][1] *= scale_log2e; s[nb][2] *= scale_log2e; s[nb][3] *= scale_log2e;
            mx_lo = fmaxf(mx_lo, fmaxf(s[nb][0], s[nb][1])); mx_hi = fmaxf(mx_hi, fmaxf(s[nb][2], s[nb][3]));
        }
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1)); mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1)); mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
        const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
        const float c_lo = exp2f(m_lo - mn_lo), c_hi = exp2f(m_hi - mn_hi);      // exp2f(-inf) = 0 on the first chunk
        m_lo = mn_lo; m_hi = mn_hi;
        l_lo *= c_lo; l_hi *= c_hi;
#pragma unroll
        for (int nb = 0; nb < 4; nb++) { o[nb][0] *= c_lo; o[nb][1] *= c_lo; o[nb][2] *= c_hi; o[nb][3] *= c_hi; }
#pragma unroll
        for (int nb = 0; nb < 8; nb++) {
            s[nb][0] = exp2f(s[nb][0] - mn_lo); s[nb][1] = exp2f(s[nb][1] - mn_lo);
            s[nb][2] = exp2f(s[nb][2] - mn_hi); s[nb][3] = exp2f(s[nb][3] - mn_hi);
            l_lo += s[nb][0] + s[nb][1]; l_hi += s[nb][2] + s[nb][3];
        }
#pragma unroll
        for (int k2 = 0; k2 < 4; k2++) {               // 16 keys per MMA k-step: S fragments 2*k2, 2*k2+1 become the A fragment of P
            uint32_t pa[4];
            pa[0] = pack_h2(s[2 * k2][0], s[2 * k2][1]); pa[1] = pack_h2(s[2 * k2][2], s[2 * k2][3]);
            pa[2] = pack_h2(s[2 * k2 + 1][0], s[2 * k2 + 1][1]); pa[3] = pack_h2(s[2 * k2 + 1][2], s[2 * k2 + 1][3]);
#pragma unroll
            for (int nb = 0; nb < 4; nb++) {
                const __half* vr = sVt + (nb * 8 + g) * vld + kc + k2 * 16 + 2 * t;
                mma_f16(o[nb], pa, *reinterpret_cast<const uint32_t*>(vr), *reinterpret_cast<const uint32_t*>(vr + 8));
            }
        }
    }
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    const float i_lo = 1.f / l_lo, i_hi = 1.f / l_hi;
#pragma unroll
    for (int nb = 0; nb < 4; nb++) {
        const int c = h * 32 + nb * 8 + 2 * t;
        store2(out, r_lo, c, o[nb][0] * i_lo, o[nb][1] * i_lo);
        store2(out, r_hi, c, o[nb][2] * i_hi, o[nb][3] * i_hi);
    }
}

}  // namespace

bool k_attention_mma_supported(int Nq, int Nk) { return Nq % 64 == 0 && Nk % 64 == 0 && Nk <= 1024; }

int k_attention_mma(const __half* q, const __half* k, const __half* v, int ld, int B, int Nq, int Nk, int heads, float scale, Out4 out, cudaStream_t st) {
    RDM_REQUIRE(k_attention_mma_supported(Nq, Nk) && ld % 8 == 0, RDM_ERR_UNSUPPORTED, "attention_mma: Nq=%d Nk=%d ld=%d", Nq, Nk, ld);
    const size_t smem = (size_t)Nk * KPAD * 2 + (size_t)32 * (Nk + VPAD) * 2;
    const float sl = scale * 1.4426950408889634f;
    if (Nq % 128 == 0) {
        static bool cfgd = false;
        if (!cfgd) { RDM_CHECK_CUDA(cudaFuncSetAttribute(attention_mma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); cfgd = true; }
        RDM_CHECK_CUDA(launch_pdl(attention_mma_kernel<128>, dim3(Nq / 128, heads, B), dim3(256), smem, st, q, k, v, ld, Nq, Nk, sl, out));
    } else {
        static bool cfgd = false;
        if (!cfgd) { RDM_CHECK_CUDA(cudaFuncSetAttribute(attention_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); cfgd = true; }
        RDM_CHECK_CUDA(launch_pdl(attention_mma_kernel<64>, dim3(Nq / 64, heads, B), dim3(128), smem, st, q, k, v, ld, Nq, Nk, sl, out));
    }
    RDM_COUNT_LAUNCH();
    RDM_CHECK_CUDA(cudaGetLastError());
    return RDM_OK;
}
