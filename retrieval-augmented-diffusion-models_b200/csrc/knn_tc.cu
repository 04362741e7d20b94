// Tensor-core main scan of the exact kNN for fp16 databases (d = 512) and 8..16 queries per pass.
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
constexpr int TM = 128, TK = 64, D = 512, KB = D / TK, NQ = 16, NCOL = 2 * NQ, STAGES = 8, THREADS = 192;
constexpr int A_BYTES = TM * TK * 2, B_KB_BYTES = NCOL * TK * 2, B_BYTES = KB * B_KB_BYTES;
constexpr int SMEM_TOTAL = STAGES * A_BYTES + B_BYTES + 1024 + 512;
constexpr int CAND_CAP = 2048;

__device__ __forceinline__ uint32_t order_f32(float f) { uint32_t b = __float_as_uint(f); return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u); }
__device__ __forceinline__ float unorder_f32(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u); }
__device__ __forceinline__ u64 make_key(float s, uint32_t idx) { return ((u64)order_f32(s) << 32) | (u64)(0xffffffffu - idx); }

// q fp32 [nq, 512] -> fp16 rows [0, NQ) = hi, [NQ, 2 NQ) = lo (zero rows beyond nq)
__global__ void split_queries_kernel(const float* __restrict__ q, int nq, __half* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NQ * D) return;
    int qi = i / D, c = i % D;
    float v = qi < nq ? q[(size_t)qi * D + c] : 0.f;
    __half h = __float2half_rn(v);
    out[(size_t)qi * D + c] = h;
    out[(size_t)(NQ + qi) * D + c] = __float2half_rn(v - __half2float(h));
}

__global__ void __launch_bounds__(THREADS, 1)
knn_scan_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const float* __restrict__ inv, long long n,
                   int nq_valid, const u64* __restrict__ thr_key, u64* __restrict__ cand, unsigned* __restrict__ cand_cnt) {
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
    const long long ntiles = (n + TM - 1) / TM;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmB);
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
        mbar_init(b_full, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < NQ) { u64 k = thr_key[threadIdx.x]; s_thr[threadIdx.x] = k == 0ull ? -CUDART_INF_F : unorder_f32((uint32_t)(k >> 32)); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64u) : "memory");
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
                    tma_load_2d(smem + s * A_BYTES, &tmA, &full[s], kb * TK, (int)(tile * TM));
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
        const int q4 = warp & 3;
        int lt = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, lt++) {
            const int buf = lt & 1;
            const long long row = tile * TM + q4 * 32 + lane;
            mbar_wait(&tmem_full[buf], (lt >> 1) & 1);
            tc_fence_after();
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * NCOL), r);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[buf])) : "memory");
            if (row < n) {
                const float iv = __ldg(inv + row);
#pragma unroll
                for (int qi = 0; qi < NQ; qi++) {
                    float s = (__uint_as_float(r[qi]) + __uint_as_float(r[NQ + qi])) * iv;
                    if (!(s == s)) s = -CUDART_INF_F;
                    if (qi < nq_valid && s >= s_thr[qi]) {
                        const u64 key = make_key(s, (uint32_t)row);
                        if (key >= thr_key[qi]) {
                            unsigned pos = atomicAdd(&cand_cnt[qi], 1u);
                            if (pos < (unsigned)CAND_CAP) cand[(size_t)qi * CAND_CAP + pos] = key;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
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

}  // namespace

int knn_tc_queries_bytes() { return 2 * NQ * D * (int)sizeof(__half); }

int knn_scan_tc(const void* db_f16, const float* inv, long long n, int device, const float* q, int nq_valid, void* qsplit_ws,
                const unsigned long long* thr_key, unsigned long long* cand, unsigned* cand_cnt, cudaStream_t st) {
    RDM_REQUIRE(nq_valid >= 1 && nq_valid <= NQ, RDM_ERR_ARG, "knn_scan_tc: %d queries", nq_valid);
    split_queries_kernel<<<(NQ * D + 255) / 256, 256, 0, st>>>(q, nq_valid, (__half*)qsplit_ws);
    RDM_COUNT_LAUNCH();
    CUtensorMap ta, tb;
    RDM_TRY(make_map(&ta, db_f16, n, TM));
    RDM_TRY(make_map(&tb, qsplit_ws, NCOL, NCOL));
    static bool configured[16] = {false};
    if (!configured[device & 15]) {
        RDM_CHECK_CUDA(cudaFuncSetAttribute(knn_scan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
        configured[device & 15] = true;
    }
    const long long ntiles = (n + TM - 1) / TM;
    const int sms = rdm_num_sms(device);
    knn_scan_tc_kernel<<<(int)(ntiles < sms ? ntiles : sms), THREADS, SMEM_TOTAL, st>>>(ta, tb, inv, n, nq_valid, thr_key, cand, cand_cnt);
    RDM_COUNT_LAUNCH();
    RDM_CHECK_CUDA(cudaGetLastError());
    return RDM_OK;
}
