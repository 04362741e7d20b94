// Tensor-core scans (sample + main) of the exact kNN for fp16 databases (d = 512), 16 / 32 / 64 / 128 query columns per pass, and (fused scan only)
// for fp32 databases on kind::tf32 MMAs, 16 / 32 / 64 query columns per pass.
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
constexpr int MAX_FUSED_Q = 128;                    // queries per fused pass (= MAX_TCQ of knn.cu)
// HILO: the queries are resident as fp16 hi + lo rows (q = hi + lo + O(2^-22): scan scores within ~1e-6 of the exact ones, score slack 3e-5);
// !HILO (round 2, passes of more than 16 queries): hi rows only -- half the MMA work, half the TMEM columns and half the epilogue, and up to
// 128 queries per pass.  The scan score then differs from the exact one by at most u * |q_hat| * |d_i| * inv_i = u = 2^-11 = 4.9e-4
// (fp16 unit roundoff, Cauchy-Schwarz; plus ~1e-6 of accumulation), so the decisions of such a pass are relaxed by HI_SLACK >= 2 * that
// bound: thresholds and the select cut admit a few more survivors, the exact fp64 re-rank of every survivor keeps the result bit-exact.
constexpr float HILO_SLACK = 3e-5f;                  // = SCORE_SLACK of knn.cu
constexpr float HI_SLACK = 1.05e-3f;
// F32 (round 2): fp32 databases on the tensor cores -- rows and queries stay fp32 in shared memory and are read as tf32 by kind::tf32 MMAs
// (32 floats = 128 bytes per k-block, 16 k-blocks per 512-d row tile, hi-only queries).  Both operands lose their low 13 mantissa bits:
// |scan score - exact score| <= 2 * 2^-10, decisions relaxed by F32_SLACK >= 2 * that bound.  Replaces the CUDA-core scan in passes of 16
// queries (0.42 of the HBM peak at 16 queries, 0.10 at 64) for fp32 databases large enough for the fused scan.
constexpr float F32_SLACK = 4.2e-3f;
template <int NQ, bool HILO, bool F32 = false> struct Cfg {
    static constexpr int KBN = F32 ? 16 : KB, TKN = F32 ? 32 : TK;             // k-blocks per row tile, elements per k-block (always 128 bytes)
    static constexpr int NCOL = (HILO ? 2 : 1) * NQ, B_KB_BYTES = NCOL * 128, B_BYTES = KBN * B_KB_BYTES;
    static_assert(!(F32 && HILO), "fp32 databases: hi-only (fp32 query rows read as tf32)");
    static constexpr int STAGES = F32 ? 4 : NCOL <= 32 ? 8 : NCOL <= 64 ? 6 : 4;      // (6 stages fit for 128 columns too, but measured no faster: that case is bound by the epilogue / MMA, not by bytes in flight)
    static constexpr int SMEM_TOTAL = STAGES * A_BYTES + B_BYTES + 1024 + 1024;
    static_assert(SMEM_TOTAL <= 232448, "kNN tensor-core scan: shared-memory budget");
    static constexpr int TMEM_COLS = 2 * NCOL <= 64 ? 64 : 2 * NCOL <= 128 ? 128 : 256;
    static_assert(2 * NCOL <= 256, "two accumulator buffers");
};

__device__ __forceinline__ uint32_t order_f32(float f) { uint32_t b = __float_as_uint(f); return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u); }
__device__ __forceinline__ float unorder_f32(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u); }
__device__ __forceinline__ u64 make_key(float s, uint32_t idx) { return ((u64)order_f32(s) << 32) | (u64)(0xffffffffu - idx); }

// q fp32 [nq, 512] -> fp16 rows [0, NQ) = hi, [NQ, 2 NQ) = lo (zero rows beyond nq)
__global__ void split_queries_kernel(const float* __restrict__ q, int nq, int NQ, __half* __restrict__ out, unsigned* __restrict__ grid_bar) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && grid_bar) *grid_bar = 0u;              // arrival counter of the fused scan's grid barrier (that kernel follows in stream order)
    if (i >= NQ * D) return;
    int qi = i / D, c = i % D;
    float v = qi < nq ? q[(size_t)qi * D + c] : 0.f;
    __half h = __float2half_rn(v);
    out[(size_t)qi * D + c] = h;
    out[(size_t)(NQ + qi) * D + c] = __float2half_rn(v - __half2float(h));
}

// fp32 databases: q fp32 [nq, 512] -> NQ fp32 rows (zero rows beyond nq) for the kind::tf32 scan
__global__ void pad_queries_f32_kernel(const float* __restrict__ q, int nq, int NQ, float* __restrict__ out, unsigned* __restrict__ grid_bar) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && grid_bar) *grid_bar = 0u;
    if (i >= NQ * D) return;
    out[i] = i / D < nq ? q[i] : 0.f;
}

// SAMPLE: visits every `tile_stride`-th row tile and writes the key of EVERY (row, query) to maxima[q][sample_row] (per_q keys per
// query; the threshold kernel takes the 32nd largest).  MAIN: all tiles, survivors of the threshold go to the candidate buffer.
template <int NQ, bool SAMPLE>
__global__ void __launch_bounds__(THREADS, 1)
knn_scan_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const float* __restrict__ inv, long long n,
                   int nq_valid, const u64* __restrict__ thr_key, u64* __restrict__ cand, unsigned* __restrict__ cand_cnt,
                   int tile_stride, u64* __restrict__ maxima, long long per_q) {
    constexpr int NCOL = Cfg<NQ, true>::NCOL, B_KB_BYTES = Cfg<NQ, true>::B_KB_BYTES, B_BYTES = Cfg<NQ, true>::B_BYTES, STAGES = Cfg<NQ, true>::STAGES;
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
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg<NQ, true>::TMEM_COLS) : "memory");
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
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg<NQ, true>::TMEM_COLS) : "memory");
    }
}


// ---- ONE launch per search: sample phase -> grid barrier -> thresholds -> main scan ---------------------------------------------------------
// The three-kernel sequence above (sample scan, threshold kernel, main scan) costs two extra launches, a second prologue and a 19 us
// selection kernel around a 187 us scan of 1.28 M rows.  Here the persistent CTAs (one per SM, cooperative launch: all co-resident) first
// run their own first `t_sample` tiles and keep, per TMEM lane quarter and query, the MAXIMUM score (an ordered u32) of the 32 * t_sample
// rows that quarter saw: 4 * grid group maxima per query, each the score of a distinct real row.  After a grid-wide barrier every CTA
// derives the thresholds redundantly: the k_eff-th largest group maximum per query (a warp-wide bisection over <= 19 values per lane with
// redux.sync counts) is a VALID lower bound of the k_eff-th best score of the whole database (k_eff distinct rows reach it), relaxed by the
// score slack like before.  The main phase then scans every tile (the sample tiles again: they are L2 hits) and appends the survivors.
// Expected survivors per query ~ k_eff * n / (32 * t_sample * 4 * grid); the host sizes t_sample for ~512.
constexpr int FUSED_MAXV = 20;                       // group maxima per lane in the threshold bisection: 4 * grid <= 32 * FUSED_MAXV

template <int NQ, bool HILO, bool F32>
__global__ void __launch_bounds__(THREADS, 1)
knn_scan_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const float* __restrict__ inv, long long n,
                      int nq_valid, int k_eff, int t_sample, u64* __restrict__ cand, unsigned* __restrict__ cand_cnt, unsigned* __restrict__ overflow,
                      uint32_t* __restrict__ gmax, unsigned* __restrict__ grid_bar, float slack) {
    constexpr int NCOL = Cfg<NQ, HILO, F32>::NCOL, B_KB_BYTES = Cfg<NQ, HILO, F32>::B_KB_BYTES, B_BYTES = Cfg<NQ, HILO, F32>::B_BYTES, STAGES = Cfg<NQ, HILO, F32>::STAGES;
    constexpr int NGRP = (NQ / 16 + EPI_PER_Q - 1) / EPI_PER_Q;          // 16-query groups per epilogue warp
    constexpr int KBN = Cfg<NQ, HILO, F32>::KBN, TKN = Cfg<NQ, HILO, F32>::TKN;
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
    pdl_launch_dependents();                            // the select kernel (launched with programmatic serialisation) may be staged behind this grid
    const long long ntiles = (n + TM - 1) / TM;
    const long long my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;      // >= 1: grid <= ntiles
    const long long ts = t_sample < my_tiles ? t_sample : my_tiles;                    // sample tiles of this CTA (its own first tiles)
    const long long nseq = ts + my_tiles;                                              // tile sequence: sample tiles, then all tiles
    auto tile_at = [&](long long p) { return (long long)blockIdx.x + (p < ts ? p : p - ts) * gridDim.x; };
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmB);
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], EPI_WARPS); }
        mbar_init(b_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg<NQ, HILO, F32>::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(b_full, B_BYTES);                          // the queries: loaded once, resident for the whole kernel
            for (int kb = 0; kb < KBN; kb++) tma_load_2d(sB + kb * B_KB_BYTES, &tmB, b_full, kb * TKN, 0);
            long long it = 0;
            for (long long p = 0; p < nseq; p++) {
                const long long tile = tile_at(p);
                for (int kb = 0; kb < KBN; kb++, it++) {
                    const int s = (int)(it % STAGES); const uint32_t ph = (uint32_t)((it / STAGES) & 1);
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], A_BYTES);
                    tma_load_2d(smem + s * A_BYTES, &tmA, &full[s], kb * TKN, (int)(tile * TM));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = F32 ? umma_idesc_tf32(NCOL) : umma_idesc_bf16(NCOL, /*f16=*/1);
            mbar_wait(b_full, 0);
            long long it = 0;
            for (long long p = 0; p < nseq; p++) {
                const int buf = (int)(p & 1);
                mbar_wait(&tmem_empty[buf], (uint32_t)(((p >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * NCOL);
                for (int kb = 0; kb < KBN; kb++, it++) {
                    const int s = (int)(it % STAGES); const uint32_t ph = (uint32_t)((it / STAGES) & 1);
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a = smem_u32(smem + s * A_BYTES), b = smem_u32(sB + kb * B_KB_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; k++) {                      // four 32-byte K slices per 128-byte k-block (16 halves or 8 floats each)
                        if (F32) umma_tf32(tmem_d, umma_desc_sw128(a + k * 32), umma_desc_sw128(b + k * 32), idesc, (kb | k) != 0);
                        else umma_bf16(tmem_d, umma_desc_sw128(a + k * 32), umma_desc_sw128(b + k * 32), idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&tmem_full[buf]);
            }
        }
    } else {
        const int q4 = warp & 3, half = (warp - 2) >> 2, ew = warp - 2;   // the EPI_PER_Q warps of a lane quarter take the 16-query groups round-robin
        uint32_t mx[NGRP][16];                                            // sample phase: running maximum (ordered score) per query of this lane's rows
#pragma unroll
        for (int g = 0; g < NGRP; g++)
#pragma unroll
            for (int j = 0; j < 16; j++) mx[g][j] = 0u;
        for (long long p = 0; p < nseq; p++) {
            if (p == ts) {
                // ---- end of the sample phase: publish the group maxima, grid barrier, thresholds ------------------------------------------
                const unsigned G4 = gridDim.x * 4u;
#pragma unroll
                for (int g = 0; g < NGRP; g++) {
                    const int q0 = (half + g * EPI_PER_Q) * 16;
                    if (q0 < NQ) {
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const uint32_t w = __reduce_max_sync(0xffffffffu, mx[g][j]);
                            if (lane == 0 && q0 + j < nq_valid) gmax[(size_t)(q0 + j) * G4 + blockIdx.x * 4u + q4] = w;
                        }
                    }
                }
                __threadfence();
                asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
                if (ew == 0 && lane == 0) {
                    if (blockIdx.x == 0) {                                  // survivor counters of this search: zero before any CTA leaves the barrier
                        for (int i = 0; i < NQ; i++) cand_cnt[i] = 0u;
                        for (int i = 0; i < NQ; i++) overflow[i] = 0u;
                        __threadfence();
                    }
                    atomicAdd(grid_bar, 1u);
                    while (*(volatile unsigned*)grid_bar < gridDim.x) __nanosleep(32);
                    __threadfence();
                }
                asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
                // thresholds: CTA c answers queries c, c + grid, ... (one warp each): the k_eff-th largest of the 4 * grid group maxima by a
                // warp-wide bisection over the ordered scores; a second grid barrier publishes them to every CTA
                float* thr_out = reinterpret_cast<float*>(gmax + (size_t)MAX_FUSED_Q * G4);
                for (int qi = blockIdx.x + ew * gridDim.x; qi < nq_valid; qi += EPI_WARPS * gridDim.x) {
                    uint32_t v[FUSED_MAXV];
#pragma unroll
                    for (int i = 0; i < FUSED_MAXV; i++) { const unsigned e = lane + 32u * i; v[i] = e < G4 ? __ldcg(gmax + (size_t)qi * G4 + e) : 0u; }
                    uint32_t res = 0u;
#pragma unroll 1
                    for (int bit = 31; bit >= 0; bit--) {
                        const uint32_t c = res | (1u << bit);
                        unsigned cnt = 0;
#pragma unroll
                        for (int i = 0; i < FUSED_MAXV; i++) cnt += v[i] >= c ? 1u : 0u;
                        cnt = __reduce_add_sync(0xffffffffu, cnt);
                        if (cnt >= (unsigned)k_eff) res = c;
                    }
                    if (lane == 0) thr_out[qi] = res == 0u ? -CUDART_INF_F : unorder_f32(res) - slack;      // fewer than k_eff sampled rows: keep everything
                }
                __threadfence();
                asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
                if (ew == 0 && lane == 0) {
                    atomicAdd(grid_bar, 1u);
                    while (*(volatile unsigned*)grid_bar < 2u * gridDim.x) __nanosleep(32);
                    __threadfence();
                }
                asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
                for (int qi = ew * 32 + lane; qi < nq_valid; qi += EPI_WARPS * 32) s_thr[qi] = __ldcg(thr_out + qi);
                asm volatile("bar.sync 1, %0;" ::"r"(EPI_WARPS * 32) : "memory");
                // main phase: the thresholds of this warp's query groups live in the registers the sample maxima occupied (as ordered-u32
                // bit patterns of the float: +inf for padding queries, so they never pass)
#pragma unroll
                for (int g = 0; g < NGRP; g++) {
                    const int q0 = (half + g * EPI_PER_Q) * 16;
#pragma unroll
                    for (int j = 0; j < 16; j++) mx[g][j] = __float_as_uint(q0 + j < nq_valid ? s_thr[q0 + j] : CUDART_INF_F);
                }
            }
            const bool sample = p < ts;
            const int buf = (int)(p & 1);
            const long long row = tile_at(p) * TM + q4 * 32 + lane;
            mbar_wait(&tmem_full[buf], (uint32_t)((p >> 1) & 1));
            tc_fence_after();
            const float iv = row < n ? __ldg(inv + row) : 0.f;
            if (half * 16 >= NQ) {                                           // fewer groups than warps (16 queries): nothing to read, release at once
                tc_fence_before();
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[buf])) : "memory");
            }
#pragma unroll
            for (int g = 0; g < NGRP; g++) {
                const int q0 = (half + g * EPI_PER_Q) * 16;
                if (q0 >= NQ) continue;
                uint32_t r[32];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                             : "r"(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * NCOL + q0)));
                if (HILO) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                 : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                                 : "r"(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * NCOL + NQ + q0)));
                } else {
#pragma unroll
                    for (int j = 16; j < 32; j++) r[j] = 0u;             // +0.0f: the lo products of a hi-only pass
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (q0 + 16 * EPI_PER_Q >= NQ) {                                // this warp's last read of the accumulator: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[buf])) : "memory");
                }
                if (sample) {
                    if (row < n) {
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            float s = (__uint_as_float(r[j]) + __uint_as_float(r[16 + j])) * iv;
                            if (!(s == s)) s = -CUDART_INF_F;
                            const uint32_t o = order_f32(s);
                            mx[g][j] = o > mx[g][j] ? o : mx[g][j];
                        }
                    }
                } else {
                    // ONE compare per (row, query) folded into a 16-bit survivor mask -- no branch per element (ncu: the per-element branch and
                    // the shared-memory threshold load were the top stall sites at 64 queries); survivors (~4e-4 of the pairs) are handled after.
                    // A NaN score (zero row: 0 * inf) takes the slow path too, where it becomes -inf as before.
                    unsigned pass = 0u;
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const float s = (__uint_as_float(r[j]) + __uint_as_float(r[16 + j])) * iv;
                        pass |= !(s < __uint_as_float(mx[g][j])) ? (1u << j) : 0u;
                    }
                    if (row >= n) pass = 0u;
                    while (pass) {
                        const int j = __ffs(pass) - 1; pass &= pass - 1u;
                        uint32_t hi_v = r[0], lo_v = r[16];
#pragma unroll
                        for (int jj = 1; jj < 16; jj++) if (jj == j) { hi_v = r[jj]; lo_v = r[16 + jj]; }
                        float s = (__uint_as_float(hi_v) + __uint_as_float(lo_v)) * iv;
                        if (!(s == s)) s = -CUDART_INF_F;
                        const int qi = q0 + j;
                        if (qi < nq_valid && s >= s_thr[qi]) {
                            const unsigned pos = atomicAdd(&cand_cnt[qi], 1u);
                            if (pos < (unsigned)CAND_CAP) cand[(size_t)qi * CAND_CAP + pos] = make_key(s, (uint32_t)row);
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
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg<NQ, HILO, F32>::TMEM_COLS) : "memory");
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
int make_map(CUtensorMap* tm, const void* base, long long rows, int box_rows, bool f32 = false) {
    auto enc = get_encode();
    RDM_REQUIRE(enc, RDM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const int esz = f32 ? 4 : 2;
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)D * esz};
    cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows}, estr[2] = {1, 1};
    CUresult r = enc(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RDM_REQUIRE(r == CUDA_SUCCESS, RDM_ERR_CUDA, "cuTensorMapEncodeTiled(kNN, %lld rows, box %d, %s) failed: %d", rows, box_rows, f32 ? "fp32" : "fp16", (int)r);
    return RDM_OK;
}


template <int NQ, bool SAMPLE>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const float* inv, long long n, int device, int nq_valid, const u64* thr_key, u64* cand,
           unsigned* cand_cnt, int tile_stride, u64* maxima, long long per_q, cudaStream_t st) {
    auto kern = knn_scan_tc_kernel<NQ, SAMPLE>;
    static bool configured[16] = {false};
    if (!configured[device & 15]) {
        RDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<NQ, true>::SMEM_TOTAL));
        configured[device & 15] = true;
    }
    const long long ntiles_all = (n + TM - 1) / TM, items = SAMPLE ? (ntiles_all + tile_stride - 1) / tile_stride : ntiles_all;
    const int sms = rdm_num_sms(device);
    kern<<<(int)(items < sms ? items : sms), THREADS, Cfg<NQ, true>::SMEM_TOTAL, st>>>(ta, tb, inv, n, nq_valid, thr_key, cand, cand_cnt, tile_stride, maxima, per_q);
    RDM_COUNT_LAUNCH();
    RDM_CHECK_CUDA(cudaGetLastError());
    return RDM_OK;
}


template <int NQ, bool HILO, bool F32 = false>
int launch_fused(const CUtensorMap& ta, const CUtensorMap& tb, const float* inv, long long n, int device, int nq_valid, int k_eff, int t_sample,
                 u64* cand, unsigned* cand_cnt, unsigned* overflow, uint32_t* gmax, unsigned* grid_bar, int grid, float slack, cudaStream_t st) {
    auto kern = knn_scan_fused_kernel<NQ, HILO, F32>;
    static int ok[16] = {0};                          // 0 unknown, 1 usable, -1 not (no cooperative launch / grid does not fit)
    if (ok[device & 15] == 0) {
        RDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<NQ, HILO, F32>::SMEM_TOTAL));
        int coop = 0, per_sm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, Cfg<NQ, HILO, F32>::SMEM_TOTAL);
        ok[device & 15] = (coop && per_sm >= 1) ? 1 : -1;
    }
    if (ok[device & 15] < 0) return 1;                // caller falls back to the three-kernel path
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = Cfg<NQ, HILO, F32>::SMEM_TOTAL; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;      // every CTA resident at once: the in-kernel grid barrier cannot deadlock
    cfg.attrs = attr; cfg.numAttrs = 1;
    RDM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, inv, n, nq_valid, k_eff, t_sample, cand, cand_cnt, overflow, gmax, grid_bar, slack));
    RDM_COUNT_LAUNCH();
    return RDM_OK;
}
}  // namespace

int knn_tc_queries_bytes() { return 2 * MAX_FUSED_Q * D * (int)sizeof(__half); }      // hi rows [0, NQ), lo rows [NQ, 2 NQ), NQ <= 128
int knn_tc_pass_queries(int nq) { return nq <= 16 ? 16 : nq <= 32 ? 32 : nq <= 64 ? 64 : 128; }
long long knn_tc_sample_rows(long long n, int tile_stride) { long long t = (n + TM - 1) / TM; return ((t + tile_stride - 1) / tile_stride) * TM; }

int knn_scan_tc(const void* db_f16, const float* inv, long long n, int device, const float* q, int nq_valid, void* qsplit_ws, int sample, int tile_stride,
                unsigned long long* maxima, long long per_q, const unsigned long long* thr_key, unsigned long long* cand, unsigned* cand_cnt, cudaStream_t st) {
    RDM_REQUIRE(nq_valid >= 1 && nq_valid <= 64, RDM_ERR_ARG, "knn_scan_tc: %d queries", nq_valid);
    const int NQ = knn_tc_pass_queries(nq_valid);
    if (sample) {                                     // the sample pass runs first: it also prepares the fp16 hi/lo query rows
        split_queries_kernel<<<(NQ * D + 255) / 256, 256, 0, st>>>(q, nq_valid, NQ, (__half*)qsplit_ws, nullptr);
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

// [128 queries][4 * grid] group maxima | [128] thresholds | grid-barrier counter
size_t knn_tc_fused_ws_bytes(int device) { return (size_t)MAX_FUSED_Q * 4 * rdm_num_sms(device) * sizeof(uint32_t) + MAX_FUSED_Q * sizeof(float) + 256; }
unsigned* knn_tc_fused_grid_bar(void* fused_ws, int device) {
    return reinterpret_cast<unsigned*>(reinterpret_cast<uint8_t*>(fused_ws) + (size_t)MAX_FUSED_Q * 4 * rdm_num_sms(device) * sizeof(uint32_t) + MAX_FUSED_Q * sizeof(float));
}

// Returns RDM_OK, an error code, or 1 when the fused path is not usable here (database too small for a sample, no cooperative launch).
int knn_scan_tc_fused(const void* db_f16, const float* inv, long long n, int device, const float* q, int nq_valid, int k, void* qsplit_ws,
                      unsigned long long* cand, unsigned* cand_cnt, unsigned* overflow, void* fused_ws, int presplit, float* slack_used, cudaStream_t st) {
    RDM_REQUIRE(nq_valid >= 1 && nq_valid <= MAX_FUSED_Q, RDM_ERR_ARG, "knn_scan_tc_fused: %d queries", nq_valid);
    const int sms = rdm_num_sms(device);
    const long long ntiles = (n + TM - 1) / TM;
    if (ntiles < 4LL * sms || 4 * sms > 32 * FUSED_MAXV) return 1;          // a sample tile per CTA next to >= 3 more; the group maxima must fit the bisection registers
    const int grid = sms;
    const int NQ = knn_tc_pass_queries(nq_valid);
    const int k_eff = k < 8 ? 8 : k;
    // survivors per query ~ k_eff * n / (32 * t_sample * 4 * grid): aim at <= ~700, between 1 tile and 1/8 of a CTA's share
    long long t_sample = (long long)((double)k_eff * (double)n / (32.0 * 4.0 * grid * 700.0)) + 1;
    const long long cap = ntiles / grid / 8 > 1 ? ntiles / grid / 8 : 1;
    if (t_sample > cap) t_sample = cap;
    uint32_t* gmax = reinterpret_cast<uint32_t*>(fused_ws);
    unsigned* grid_bar = knn_tc_fused_grid_bar(fused_ws, device);
    if (!presplit) {                                   // (rdm_knn_search_raw: the normalisation kernel already wrote the fp16 hi / lo rows and reset the barrier)
        split_queries_kernel<<<(NQ * D + 255) / 256, 256, 0, st>>>(q, nq_valid, NQ, (__half*)qsplit_ws, grid_bar);
        RDM_COUNT_LAUNCH();
    }
    CUtensorMap ta, tb;
    RDM_TRY(make_map(&ta, db_f16, n, TM));
    // passes of up to 16 queries keep the hi + lo query rows (HBM-bound already, tight slack); wider passes use the hi rows only
    static const int force_hilo = getenv("RDM_KNN_HILO") ? 1 : 0;       // A/B: hi + lo rows for every pass (then <= 64 queries per pass)
    const bool hilo = NQ == 16 || (force_hilo && NQ <= 64);
    if (force_hilo && NQ > 64) return 1;
    *slack_used = hilo ? HILO_SLACK : HI_SLACK;
    RDM_TRY(make_map(&tb, qsplit_ws, hilo ? 2 * NQ : NQ, hilo ? 2 * NQ : NQ));
    if (hilo) {
        if (NQ == 16) return launch_fused<16, true>(ta, tb, inv, n, device, nq_valid, k_eff, (int)t_sample, cand, cand_cnt, overflow, gmax, grid_bar, grid, HILO_SLACK, st);
        if (NQ == 32) return launch_fused<32, true>(ta, tb, inv, n, device, nq_valid, k_eff, (int)t_sample, cand, cand_cnt, overflow, gmax, grid_bar, grid, HILO_SLACK, st);
        return launch_fused<64, true>(ta, tb, inv, n, device, nq_valid, k_eff, (int)t_sample, cand, cand_cnt, overflow, gmax, grid_bar, grid, HILO_SLACK, st);
    }
    if (NQ == 32) return launch_fused<32, false>(ta, tb, inv, n, device, nq_valid, k_eff, (int)t_sample, cand, cand_cnt, overflow, gmax, grid_bar, grid, HI_SLACK, st);
    if (NQ == 64) return launch_fused<64, false>(ta, tb, inv, n, device, nq_valid, k_eff, (int)t_sample, cand, cand_cnt, overflow, gmax, grid_bar, grid, HI_SLACK, st);
    return launch_fused<128, false>(ta, tb, inv, n, device, nq_valid, k_eff, (int)t_sample, cand, cand_cnt, overflow, gmax, grid_bar, grid, HI_SLACK, st);
}

// fp32 database (d = 512): the same fused scan on kind::tf32 MMAs, passes of up to 64 queries.  Same return convention as knn_scan_tc_fused.
int knn_scan_tc_fused_f32(const void* db_f32, const float* inv, long long n, int device, const float* q, int nq_valid, int k, void* qpad_ws,
                          unsigned long long* cand, unsigned* cand_cnt, unsigned* overflow, void* fused_ws, float* slack_used, cudaStream_t st) {
    RDM_REQUIRE(nq_valid >= 1 && nq_valid <= 64, RDM_ERR_ARG, "knn_scan_tc_fused_f32: %d queries", nq_valid);
    static const int off = getenv("RDM_KNN_NO_TF32") ? 1 : 0;
    const int sms = rdm_num_sms(device);
    const long long ntiles = (n + TM - 1) / TM;
    if (off || ntiles < 4LL * sms || 4 * sms > 32 * FUSED_MAXV) return 1;
    const int grid = sms;
    const int NQ = nq_valid <= 16 ? 16 : nq_valid <= 32 ? 32 : 64;
    const int k_eff = k < 8 ? 8 : k;
    long long t_sample = (long long)((double)k_eff * (double)n / (32.0 * 4.0 * grid * 700.0)) + 1;
    const long long cap = ntiles / grid / 8 > 1 ? ntiles / grid / 8 : 1;
    if (t_sample > cap) t_sample = cap;
    uint32_t* gmax = reinterpret_cast<uint32_t*>(fused_ws);
    unsigned* grid_bar = knn_tc_fused_grid_bar(fused_ws, device);
    pad_queries_f32_kernel<<<(NQ * D + 255) / 256, 256, 0, st>>>(q, nq_valid, NQ, (float*)qpad_ws, grid_bar);
    RDM_COUNT_LAUNCH();
    CUtensorMap ta, tb;
    RDM_TRY(make_map(&ta, db_f32, n, TM, true));
    RDM_TRY(make_map(&tb, qpad_ws, NQ, NQ, true));
    *slack_used = F32_SLACK;
    if (NQ == 16) return launch_fused<16, false, true>(ta, tb, inv, n, device, nq_valid, k_eff, (int)t_sample, cand, cand_cnt, overflow, gmax, grid_bar, grid, F32_SLACK, st);
    if (NQ == 32) return launch_fused<32, false, true>(ta, tb, inv, n, device, nq_valid, k_eff, (int)t_sample, cand, cand_cnt, overflow, gmax, grid_bar, grid, F32_SLACK, st);
    return launch_fused<64, false, true>(ta, tb, inv, n, device, nq_valid, k_eff, (int)t_sample, cand, cand_cnt, overflow, gmax, grid_bar, grid, F32_SLACK, st);
}
