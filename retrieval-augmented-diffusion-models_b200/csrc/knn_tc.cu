// Tensor-core scans (sample + main) of the exact kNN for fp16 databases (d = 512), 16 / 32 / 64 query columns per pass.
//
// With >= 8 queries the CUDA-core scan (knn.cu) is FMA-bound; here the 128-row database tile is the A operand of tcgen05.mma
// straight from TMA (the stored fp16 rows are used as they are -- no conversion pass), the queries are the B operand, resident in
// shared memory as fp16 hi + lo rows (q = hi + lo + O(2^-22): products with the exact fp16 database values are exact in the
// fp32 accumulator), so one row tile costs 32 tiny MMAs (N = 32 columns) and the kernel is HBM-bound again:
//   warp 0  : TMA producer, ring of 16 KB stages (128 rows x 64 halves, 128B swizzle), 8 stages per row tile
//   warp 1  : MMA issuer; accumulators [128 rows x 32] double-buffered in TMEM
//   warps 2-5: epilogue -- tcgen05.ld, score = (acc_hi + acc_lo) * inv_norm[row], ONE compare against the threshold of the
//              sample pass, rare survivors appended to the global candidate buffer (same contract as knn_scan_kernel<MAIN>)
// Exactness is unchanged: the survivors are re-ranked in fp64 by knn_select_kernel.
#include "common.cuh"
#include "ptx.cuh"
#include "knn_tc.cuh"
#include <cudaTypedefs.h>
#include <math_constants.h>

namespace {

typedef unsigned long long u64;
constexpr int TM = 128, TK = 64, D = 512, KB = D / TK, EPI_WARPS = 8, EPI_PER_Q = EPI_WARPS / 4, THREADS = 64 + EPI_WARPS * 32;      // warp 0 TMA, warp 1 MMA, two epilogue warps per TMEM lane quarter (4 -> 8 warps: 20 M rows x 64 queries 54 % -> 72 % of HBM peak; 16 warps measured no better)
constexpr int A_BYTES = TM * TK * 2;
constexpr int CAND_CAP = 2048;
template <int NQ> struct Cfg {
    static constexpr int NCOL = 2 * NQ, B_KB_BYTES = NCOL * TK * 2, B_BYTES = KB * B_KB_BYTES;
    static constexpr int STAGES = NQ <= 16 ? 8 : NQ <= 32 ? 6 : 4;      // (6 stages fit for 64 queries too, but measured no faster: that case is bound by the epilogue / MMA, not by bytes in flight)
    static constexpr int SMEM_TOTAL = STAGES * A_BYTES + B_BYTES + 1024 + 1024;
    static_assert(SMEM_TOTAL <= 232448, "kNN tensor-core scan: shared-memory budget");
    static constexpr int TMEM_COLS = 2 * NCOL <= 64 ? 64 : 2 * NCOL <= 128 ? 128 : 256;
};

__device__ __forceinline__ uint32_t order_f32(float f) { uint32_t b = __float_as_uint(f); return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u); }
__device__ __forceinline__ float unorder_f32(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u); }
__device__ __forceinline__ u64 make_key(float s, uint32_t idx) { return ((u64)order_f32(s) << 32) | (u64)(0xffffffffu - idx); }

// q fp32 [nq, 512] -> fp16 rows [0, NQ) = hi, [NQ, 2 NQ) = lo (zero rows beyond nq)
__global__ void split_queries_kernel(const float* __restrict__ q, int nq, int NQ, __half* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NQ * D) return;
    int qi = i / D, c = i % D;
    float v = qi < nq ? q[(size_t)qi * D + c] : 0.f;
    __half h = __float2half_rn(v);
    out[(size_t)qi * D + c] = h;
    out[(size_t)(NQ + qi) * D + c] = __float2half_rn(v - __half2float(h));
}

// SAMPLE: visits every `tile_stride`-th row tile and writes the key of EVERY (row, query) to maxima[q][sample_row] (per_q keys per
// query; the threshold kernel takes the 32nd largest).  MAIN: all tiles, survivors of the threshold go to the candidate buffer.
template <int NQ, bool SAMPLE>
__global__ void __launch_bounds__(THREADS, 1)
knn_scan_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const float* __restrict__ inv, long long n,
                   int nq_valid, const u64* __restrict__ thr_key, u64* __restrict__ cand, unsigned* __restrict__ cand_cnt,
                   int tile_stride, u64* __restrict__ maxima, long long per_q) {
    constexpr int NCOL = Cfg<NQ>::NCOL, B_KB_BYTES = Cfg<NQ>::B_KB_BYTES, B_BYTES = Cfg<NQ>::B_BYTES, STAGES = Cfg<NQ>::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + B_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* b_full = tmem_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);
    float* s_thr = reinterpret_cast<float*>(tmem_slot + 1);           // [NQ]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ntiles_all = (n + TM - 1) / TM;
    const long long ntiles = SAMPLE ? (ntiles_all + tile_stride - 1) / tile_stride : ntiles_all;       // work items of this launch
    const long long tmul = SAMPLE ? tile_stride : 1;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmB);
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], EPI_WARPS); }
        mbar_init(b_full, 1);
        fence_barrier_init();
    }
    if (!SAMPLE && threadIdx.x < NQ) { u64 k = thr_key[threadIdx.x]; s_thr[threadIdx.x] = k == 0ull ? -CUDART_INF_F : unorder_f32((uint32_t)(k >> 32)); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg<NQ>::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(b_full, B_BYTES);                          // the queries: loaded once, resident for the whole kernel
            for (int kb = 0; kb < KB; kb++) tma_load_2d(sB + kb * B_KB_BYTES, &tmB, b_full, kb * TK, 0);
            long long it = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int kb = 0; kb < KB; kb++, it++) {
                    const int s = (int)(it % STAGES); const uint32_t ph = (uint32_t)((it / STAGES) & 1);
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], A_BYTES);
                    tma_load_2d(smem + s * A_BYTES, &tmA, &full[s], kb * TK, (int)(tile * tmul * TM));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(NCOL, /*f16=*/1);
            mbar_wait(b_full, 0);
            long long it = 0; int lt = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, lt++) {
                const int buf = lt & 1;
                mbar_wait(&tmem_empty[buf], ((lt >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * NCOL);
                for (int kb = 0; kb < KB; kb++, it++) {
                    const int s = (int)(it % STAGES); const uint32_t ph = (uint32_t)((it / STAGES) & 1);
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a = smem_u32(smem + s * A_BYTES), b = smem_u32(sB + kb * B_KB_BYTES);
#pragma unroll
                    for (int k = 0; k < TK / 16; k++) umma_bf16(tmem_d, umma_desc_sw128(a + k * 32), umma_desc_sw128(b + k * 32), idesc, (kb | k) != 0);
                    umma_commit(&empty[s]);
                }
                umma_commit(&tmem_full[buf]);
            }
        }
    } else {
        const int q4 = warp & 3, half = (warp - 2) >> 2;                 // the EPI_PER_Q warps of a lane quarter take the 16-query groups round-robin
        int lt = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, lt++) {
            const int buf = lt & 1;
            const long long row = tile * tmul * TM + q4 * 32 + lane;
            const long long srow = tile * TM + q4 * 32 + lane;                  // position in the sample (SAMPLE mode)
            mbar_wait(&tmem_full[buf], (lt >> 1) & 1);
            tc_fence_after();
            const float iv = row < n ? __ldg(inv + row) : 0.f;
            // columns [0, NQ) = hi products, [NQ, 2 NQ) = lo products; processed 16 queries at a time (two 32-bit x16 halves)
            if (half * 16 >= NQ) {                                           // fewer groups than warps (16 queries): nothing to read, release at once
                tc_fence_before();
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[buf])) : "memory");
            }
#pragma unroll 1
            for (int q0 = half * 16; q0 < NQ; q0 += 16 * EPI_PER_Q) {
                uint32_t r[32];
                {
                    uint32_t rh[32];
                    // hi block: columns q0..q0+15 ; lo block: NQ+q0 .. NQ+q0+15 (two x16 loads packed into one x32 array)
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                 : "=r"(rh[0]), "=r"(rh[1]), "=r"(rh[2]), "=r"(rh[3]), "=r"(rh[4]), "=r"(rh[5]), "=r"(rh[6]), "=r"(rh[7]),
                                   "=r"(rh[8]), "=r"(rh[9]), "=r"(rh[10]), "=r"(rh[11]), "=r"(rh[12]), "=r"(rh[13]), "=r"(rh[14]), "=r"(rh[15])
                                 : "r"(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * NCOL + q0)));
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                 : "=r"(rh[16]), "=r"(rh[17]), "=r"(rh[18]), "=r"(rh[19]), "=r"(rh[20]), "=r"(rh[21]), "=r"(rh[22]), "=r"(rh[23]),
                                   "=r"(rh[24]), "=r"(rh[25]), "=r"(rh[26]), "=r"(rh[27]), "=r"(rh[28]), "=r"(rh[29]), "=r"(rh[30]), "=r"(rh[31])
                                 : "r"(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * NCOL + NQ + q0)));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; j++) r[j] = rh[j];
                }
                if (q0 + 16 * EPI_PER_Q >= NQ) {                                // this warp's last read of the accumulator: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[buf])) : "memory");
                }
                if (row < n) {
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int qi = q0 + j;
                        float s = (__uint_as_float(r[j]) + __uint_as_float(r[16 + j])) * iv;
                        if (!(s == s)) s = -CUDART_INF_F;
                        if (SAMPLE) {
                            if (qi < nq_valid) maxima[(size_t)qi * per_q + srow] = make_key(s, (uint32_t)row);
                        } else if (qi < nq_valid && s >= s_thr[qi]) {
                            const u64 key = make_key(s, (uint32_t)row);
                            if (key >= thr_key[qi]) {
                                unsigned pos = atomicAdd(&cand_cnt[qi], 1u);
                                if (pos < (unsigned)CAND_CAP) cand[(size_t)qi * CAND_CAP + pos] = key;
                            }
                        }
                    }
                } else if (SAMPLE && srow < per_q) {
#pragma unroll 1
                    for (int j = 0; j < 16; j++) if (q0 + j < nq_valid) maxima[(size_t)(q0 + j) * per_q + srow] = 0ull;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg<NQ>::TMEM_COLS) : "memory");
    }
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr; cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    }
    return fn;
}
int make_map(CUtensorMap* tm, const void* base, long long rows, int box_rows) {
    auto enc = get_encode();
    RDM_REQUIRE(enc, RDM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)D * 2};
    cuuint32_t box[2] = {(cuuint32_t)TK, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RDM_REQUIRE(r == CUDA_SUCCESS, RDM_ERR_CUDA, "cuTensorMapEncodeTiled(knn rows=%lld) failed: %d", rows, (int)r);
    return RDM_OK;
}


template <int NQ, bool SAMPLE>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const float* inv, long long n, int device, int nq_valid, const u64* thr_key, u64* cand,
           unsigned* cand_cnt, int tile_stride, u64* maxima, long long per_q, cudaStream_t st) {
    auto kern = knn_scan_tc_kernel<NQ, SAMPLE>;
    static bool configured[16] = {false};
    if (!configured[device & 15]) {
        RDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<NQ>::SMEM_TOTAL));
        configured[device & 15] = true;
    }
    const long long ntiles_all = (n + TM - 1) / TM, items = SAMPLE ? (ntiles_all + tile_stride - 1) / tile_stride : ntiles_all;
    const int sms = rdm_num_sms(device);
    kern<<<(int)(items < sms ? items : sms), THREADS, Cfg<NQ>::SMEM_TOTAL, st>>>(ta, tb, inv, n, nq_valid, thr_key, cand, cand_cnt, tile_stride, maxima, per_q);
    RDM_COUNT_LAUNCH();
    RDM_CHECK_CUDA(cudaGetLastError());
    return RDM_OK;
}

}  // namespace

int knn_tc_queries_bytes() { return 2 * 64 * D * (int)sizeof(__half); }
int knn_tc_pass_queries(int nq) { return nq <= 16 ? 16 : nq <= 32 ? 32 : 64; }
long long knn_tc_sample_rows(long long n, int tile_stride) { long long t = (n + TM - 1) / TM; return ((t + tile_stride - 1) / tile_stride) * TM; }

int knn_scan_tc(const void* db_f16, const float* inv, long long n, int device, const float* q, int nq_valid, void* qsplit_ws, int sample, int tile_stride,
                unsigned long long* maxima, long long per_q, const unsigned long long* thr_key, unsigned long long* cand, unsigned* cand_cnt, cudaStream_t st) {
    RDM_REQUIRE(nq_valid >= 1 && nq_valid <= 64, RDM_ERR_ARG, "knn_scan_tc: %d queries", nq_valid);
    const int NQ = knn_tc_pass_queries(nq_valid);
    if (sample) {                                     // the sample pass runs first: it also prepares the fp16 hi/lo query rows
        split_queries_kernel<<<(NQ * D + 255) / 256, 256, 0, st>>>(q, nq_valid, NQ, (__half*)qsplit_ws);
        RDM_COUNT_LAUNCH();
    }
    CUtensorMap ta, tb;
    RDM_TRY(make_map(&ta, db_f16, n, TM));
    RDM_TRY(make_map(&tb, qsplit_ws, 2 * NQ, 2 * NQ));
#define KNN_TC_GO(NQV) (sample ? launch<NQV, true>(ta, tb, inv, n, device, nq_valid, thr_key, cand, cand_cnt, tile_stride, maxima, per_q, st) \
                               : launch<NQV, false>(ta, tb, inv, n, device, nq_valid, thr_key, cand, cand_cnt, tile_stride, maxima, per_q, st))
    if (NQ == 16) return KNN_TC_GO(16);
    if (NQ == 32) return KNN_TC_GO(32);
    return KNN_TC_GO(64);
#undef KNN_TC_GO
}
