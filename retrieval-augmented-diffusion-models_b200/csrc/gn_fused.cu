// GroupNorm (+ SiLU) in ONE launch for the tensor-core engine modes (ldm GroupNorm32 + SiLU of ResBlock in_layers / out_layers, the
// SpatialTransformer norm and the output head; SURVEY.md Appendix A, rdm/modules/attention.py:183).
//
// The strict fp32 mode keeps the two-kernel form (gn_stats_kernel + gn_apply_kernel, kernels.cu), which also runs under the host emulation.
// Here the statistics never leave the chip:
//   CLUSTER form  the CTAs that share one image form a thread-block cluster (<= 8 CTAs).  Phase 1: every CTA accumulates the fp64 sums of
//                 its rows per group in shared memory.  Cluster barrier.  Every CTA then adds up the partial sums of all ranks through
//                 distributed shared memory IN RANK ORDER (bit-reproducible, unlike global atomics) and normalises its own rows (second
//                 read of x: L1 / L2 hits).  No statistics buffer, no memset, one launch instead of two.
//   PRE form      the producing tcgen05 GEMM already left per-(image, channel) sums in `chan` (gemm_tc.cu epilogue): the group sums are
//                 folded from them in the prologue and the rows are read ONCE.
// Arithmetic of the apply phase is the one of gn_apply_kernel (same rounding sequence given the same sums).
#include "kernels.cuh"
#include "ptx.cuh"

namespace {

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }
__device__ __forceinline__ float silu_fast(float x) { return x * __frcp_rn(1.f + __expf(-x)); }
__device__ __forceinline__ void store4(const Out4& o, size_t m, int c, float a, float b, float cc, float d) {
    if (o.f) *reinterpret_cast<float4*>(o.f + m * o.ldf + c) = make_float4(a, b, cc, d);
    if (o.hi) store_planes4(o.hi + m * o.ldb + c, o.lo ? o.lo + m * o.ldb + c : nullptr, o.f16, a, b, cc, d);
}
__device__ __forceinline__ double ld_dsmem_f64(uint32_t cta_addr, uint32_t rank) {
    uint32_t ra; double v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(cta_addr), "r"(rank));
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
    return v;
}

constexpr int GN_MAX_GROUPS = 64;

// grid (chunks, B); CLUSTER form: cluster (chunks, 1, 1).  Thread t owns the float4 column v = t % V and walks rows slot, slot + nslots, ...
template <bool PRE>
__global__ void __launch_bounds__(1024) gn_fused_kernel(const float* __restrict__ x, int ld, int C, int HW, int groups, const double* __restrict__ chan, int chan_ld,
                                                        float eps, const float* __restrict__ gamma, const float* __restrict__ beta, int silu, Out4 y, Out4 raw, int rows_per_cta) {
    __shared__ double s_acc[2 * GN_MAX_GROUPS];
    __shared__ float s_mean[GN_MAX_GROUPS], s_rstd[GN_MAX_GROUPS];
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.y, V = C / 4, cpg = C / groups;
    const int nslots = blockDim.x / V, v = threadIdx.x % V, slot = threadIdx.x / V;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(HW, r0 + rows_per_cta);
    const size_t m0 = (size_t)b * HW;
    const double cnt = (double)HW * cpg;
    if (PRE) {
        if (threadIdx.x < groups) {
            const double* cs = chan + ((size_t)b * chan_ld + (size_t)threadIdx.x * cpg) * 2;
            double s = 0.0, ss = 0.0;
            for (int c = 0; c < cpg; c++) { s += cs[2 * c]; ss += cs[2 * c + 1]; }
            const double mean = s / cnt, var = ss / cnt - mean * mean;
            s_mean[threadIdx.x] = (float)mean;
            s_rstd[threadIdx.x] = (float)(1.0 / sqrt((var > 0 ? var : 0.0) + (double)eps));
        }
    } else {
        for (int i = threadIdx.x; i < 2 * groups; i += blockDim.x) s_acc[i] = 0.0;
        __syncthreads();
        if (slot < nslots) {
            const float* base = x + m0 * ld + v * 4;
            double s[4] = {0.0, 0.0, 0.0, 0.0}, ss[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
            for (int r = r0 + slot; r < r1; r += nslots) {
                const float4 q = *reinterpret_cast<const float4*>(base + (long long)r * ld);
                const double q0 = q.x, q1 = q.y, q2 = q.z, q3 = q.w;
                s[0] += q0; ss[0] = fma(q0, q0, ss[0]); s[1] += q1; ss[1] = fma(q1, q1, ss[1]);
                s[2] += q2; ss[2] = fma(q2, q2, ss[2]); s[3] += q3; ss[3] = fma(q3, q3, ss[3]);
            }
            int g = (v * 4) / cpg; double gs = 0.0, gss = 0.0;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int gt = (v * 4 + t) / cpg;
                if (gt != g) { atomicAdd(&s_acc[g], gs); atomicAdd(&s_acc[groups + g], gss); g = gt; gs = 0.0; gss = 0.0; }
                gs += s[t]; gss += ss[t];
            }
            atomicAdd(&s_acc[g], gs); atomicAdd(&s_acc[groups + g], gss);
        }
        __syncthreads();
        cluster_sync_all();                                 // every rank's partial sums are complete and visible
        if (threadIdx.x < groups) {
            const uint32_t a0 = smem_u32(&s_acc[threadIdx.x]), a1 = smem_u32(&s_acc[groups + threadIdx.x]);
            double s = 0.0, ss = 0.0;
            for (uint32_t r = 0; r < gridDim.x; r++) { s += ld_dsmem_f64(a0, r); ss += ld_dsmem_f64(a1, r); }       // cluster = the whole x extent of the grid
            const double mean = s / cnt, var = ss / cnt - mean * mean;
            s_mean[threadIdx.x] = (float)mean;
            s_rstd[threadIdx.x] = (float)(1.0 / sqrt((var > 0 ? var : 0.0) + (double)eps));
        }
        __syncwarp();
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");      // this CTA is done reading its peers; waited for before exit
    }
    __syncthreads();
    if (slot < nslots) {
        float sc[4], sh[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int c = v * 4 + t, g = c / cpg;
            sc[t] = s_rstd[g] * gamma[c];
            sh[t] = beta[c] - s_mean[g] * sc[t];
        }
        const bool want_raw = raw.any();
        const bool fast = y.f == nullptr && y.lo == nullptr;   // single 16-bit plane output: MUFU exp + reciprocal are exact enough (as gn_apply_kernel)
#pragma unroll 2
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
    __syncwarp();
    if (!PRE) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // no CTA may exit while a peer still reads its shared memory
}

}  // namespace

// chan != nullptr: PRE form (per-(image, channel) {sum, sumsq} pairs, image stride chan_ld channels).  Returns RDM_ERR_UNSUPPORTED (without
// setting the error text) when the shape does not suit the cluster form -- the caller then uses the two-kernel path.
bool k_gn_fused_supported(int C, int HW, int groups, bool pre) {
    if (C % 4 || C % groups || groups > GN_MAX_GROUPS || C / 4 > 1024) return false;
    if (pre) return true;
    const int V = C / 4, nslots = 1024 / V;
    return (HW + 7) / 8 <= 24 * (nslots < 1 ? 1 : nslots);       // <= 24 rows per thread with the largest block and an 8-CTA cluster
}
int k_gn_fused(View x, int B, int HW, int groups, const double* chan, int chan_ld, float eps, const float* gamma, const float* beta, int silu, Out4 y, Out4 raw, cudaStream_t st) {
    RDM_REQUIRE(k_gn_fused_supported(x.C, HW, groups, chan != nullptr) && x.ld % 4 == 0 && y.ldf % 4 == 0 && y.ldb % 4 == 0, RDM_ERR_UNSUPPORTED, "gn_fused: C=%d HW=%d groups=%d", x.C, HW, groups);
    const int V = x.C / 4;
    auto threads_for = [&](int cap) { int t = V >= cap ? ((V + 31) / 32) * 32 : (cap / V) * V; return t < groups ? ((groups + 31) / 32) * 32 : t; };
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[2]; int na = 0;
    if (g_rdm_use_pdl && g_rdm_use_pdl_glue) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; na++; }
    cfg.stream = st; cfg.attrs = attr;
    if (chan) {
        const int threads = threads_for(256);
        int nslots = threads / V; if (nslots < 1) nslots = 1;
        int rows_per_cta = nslots * 4, chunks = (HW + rows_per_cta - 1) / rows_per_cta;
        while (chunks * B > 148 * 16 && rows_per_cta < HW) { rows_per_cta *= 2; chunks = (HW + rows_per_cta - 1) / rows_per_cta; }
        cfg.gridDim = dim3(chunks, B); cfg.blockDim = dim3(threads); cfg.numAttrs = na;
        RDM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gn_fused_kernel<true>, (const float*)x.p, x.ld, x.C, HW, groups, chan, chan_ld, eps, gamma, beta, silu, y, raw, rows_per_cta));
    } else {
        // cluster size: as many CTAs per image as keeps >= 2 rows per thread slot (<= 8, power of two); block size: 512 threads unless the
        // rows per thread would exceed 12, then up to 1024
        int threads = threads_for(512);
        int nslots = threads / V; if (nslots < 1) nslots = 1;
        int cs = 8;
        while (cs > 1 && (HW + cs - 1) / cs < 2 * nslots) cs >>= 1;
        if (((HW + cs - 1) / cs + nslots - 1) / nslots > 12) { threads = threads_for(1024); nslots = threads / V; if (nslots < 1) nslots = 1; }
        const int rows_per_cta = (HW + cs - 1) / cs;
        attr[na].id = cudaLaunchAttributeClusterDimension; attr[na].val.clusterDim.x = cs; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1; na++;
        cfg.gridDim = dim3(cs, B); cfg.blockDim = dim3(threads); cfg.numAttrs = na;
        RDM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gn_fused_kernel<false>, (const float*)x.p, x.ld, x.C, HW, groups, (const double*)nullptr, 0, eps, gamma, beta, silu, y, raw, rows_per_cta));
    }
    RDM_COUNT_LAUNCH();
    RDM_CHECK_CUDA(cudaGetLastError());
    return RDM_OK;
}
