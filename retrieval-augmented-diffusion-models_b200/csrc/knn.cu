// Exact cosine top-k over an in-HBM CLIP database (K1/K11 of SURVEY.md section 2.1).
//
// Pipeline per query batch of <= 16 queries (all stream-ordered, no host sync):
//   1. knn_scan_kernel<SAMPLE> -- a strided 1/16 sample of the row groups: every lane keeps the running maximum
//                           (score, row) key of the rows it owns.  knn_threshold_kernel takes the 32nd largest of
//                           those maxima per query: a VALID lower bound of the final 32nd-best key (they are 32
//                           distinct real rows), i.e. a threshold only ~16*32 rows of the whole DB exceed.
//   2. knn_scan_kernel<MAIN>   -- ONE pass over the database: coalesced 16-byte streaming loads, fp32 FMA dot
//                           products against the queries held in shared memory, transposing warp-shuffle
//                           reduction, ONE compare per (row, query) against the threshold; the rare survivors
//                           are appended to a small global candidate buffer.  HBM-bound for few queries:
//                           algorithmic bytes = n * d * sizeof(elem) (+4 B/row inverse norm).
//   3. knn_select_kernel -- the k-th best fp32 key of the survivors (warp bitonic networks) minus SCORE_SLACK is the cut; every
//                           survivor above the cut (<= 1024, usually ~k) is RE-RANKED with the exact definition of
//                           oracle/knn_ref.c (sequential fp64, separately rounded products) and the top-k is emitted
//                           by (score desc, index asc).  The fp32 / tensor-core scores only decide, with the slack as
//                           margin over their error (~1e-6), which rows reach the exact re-rank.
//   4. fallback (device-side conditional, normally two empty launches): if a candidate buffer overflowed (sample
//      unrepresentative of the DB), knn_scan_kernel<LOCKED> redoes the pass with per-CTA top-32 lists under a lock.
// Replaces ScaNN behind `searcher.search_batched` (dsetbuilder.py:490, ddpm.py:906-908).
#include "common.cuh"
#include "ptx.cuh"
#include "knn_tc.cuh"
#include <type_traits>
#include <cstdlib>
#include <algorithm>
#include "../../include/rdm_b200.h"
#include <math_constants.h>

namespace {

constexpr int LIST = 32;          // candidates kept per query (one per lane)
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int MAX_QP = 16;        // queries per CUDA-core scan pass
constexpr int MAX_TCQ = 128;      // queries per tensor-core pass (fp16 databases): 128 with the fused hi-only scan, 64 on the three-kernel path
constexpr int QHAT_ROWS = 1024;   // rdm_knn_search_raw normalises (and searches) this many raw queries at a time
constexpr unsigned FULL = 0xffffffffu;
typedef unsigned long long u64;

__device__ __forceinline__ uint32_t order_f32(float f) {
    uint32_t b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float unorder_f32(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
}
// larger key = better candidate: higher score first, then LOWER row index
__device__ __forceinline__ u64 make_key(float s, uint32_t idx) { return ((u64)order_f32(s) << 32) | (u64)(0xffffffffu - idx); }
// The scans rank rows by an fp32 (or tensor-core) score whose distance to the exact fp64 score of the result definition is bounded by
// e (a few 1e-6 for unit queries and normalised 512-d rows; the slack below allows e = 1.5e-5).  Wherever an fp32 score decides whether a
// row can still be one of the exact top-k, the comparison is relaxed by 2e: with f_k the k-th largest fp32 score seen, every row of the exact top-k has an fp32 score >= f_k - 2e.
// (Without the slack, a cluster of near-duplicate rows -- scores closer than the fp32 rounding -- could push true neighbours out.)
constexpr float SCORE_SLACK = 3e-5f;
__device__ __forceinline__ u64 relax_key(u64 key, float slack = SCORE_SLACK) {         // key of (score - slack) with the row part cleared; 0 stays "keep everything"
    if (key == 0ull) return 0ull;
    return (u64)order_f32(unorder_f32((uint32_t)(key >> 32)) - slack) << 32;
}

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <typename T> struct Elem;
template <> struct Elem<float> {
    static constexpr int EPV = 4;
    __device__ static __forceinline__ void unpack(const uint4& v, float* o) {
        o[0] = __uint_as_float(v.x); o[1] = __uint_as_float(v.y); o[2] = __uint_as_float(v.z); o[3] = __uint_as_float(v.w);
    }
    __device__ static __forceinline__ float to_f32(float v) { return v; }
};
template <> struct Elem<__half> {
    static constexpr int EPV = 8;
    __device__ static __forceinline__ void unpack(const uint4& v, float* o) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            o[2 * i] = f.x; o[2 * i + 1] = f.y;
        }
    }
    __device__ static __forceinline__ float to_f32(__half v) { return __half2float(v); }
};

__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int m) {
    uint32_t lo = __shfl_xor_sync(FULL, (uint32_t)v, m), hi = __shfl_xor_sync(FULL, (uint32_t)(v >> 32), m);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 shfl_u64(u64 v, int src) {
    uint32_t lo = __shfl_sync(FULL, (uint32_t)v, src), hi = __shfl_sync(FULL, (uint32_t)(v >> 32), src);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 warp_min_u64(u64 v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) { u64 o = shfl_xor_u64(v, m); v = o < v ? o : v; }
    return v;
}
// ascending bitonic sort of one key per lane
__device__ __forceinline__ u64 warp_sort_asc(u64 v, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            u64 o = shfl_xor_u64(v, j);
            bool asc = (lane & k) == 0, lower = (lane & j) == 0;
            bool keep_min = (asc == lower);
            v = keep_min ? (o < v ? o : v) : (o > v ? o : v);
        }
    }
    return v;
}
// cur: descending-sorted (lane 0 best); batch: arbitrary.  Returns descending-sorted top-32 of the union.
__device__ __forceinline__ u64 warp_merge_top32(u64 cur, u64 batch, int lane) {
    u64 b = warp_sort_asc(batch, lane);
    u64 v = cur > b ? cur : b;                 // half-cleaner of the bitonic sequence (batch asc | cur desc)
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {         // bitonic merge, descending
        u64 o = shfl_xor_u64(v, j);
        bool lower = (lane & j) == 0;
        v = lower ? (o > v ? o : v) : (o < v ? o : v);
    }
    return v;
}

// Transposing reduction: R per-lane partial sums -> the complete sum of row `rsel(lane)` in lane groups.
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
template <int R> __device__ __forceinline__ int row_of_lane(int lane) {
    int r = 0, width = 16;
#pragma unroll
    for (int cnt = R; cnt > 1; cnt >>= 1, width >>= 1) if (lane & width) r += cnt / 2;
    return r;
}

constexpr int SCAN_MAX_STAGES = 3;
struct ScanShared {
    u64 list[MAX_QP][LIST];
    float thr[MAX_QP];
    int lock[MAX_QP];
    uint64_t full[SCAN_WARPS][SCAN_MAX_STAGES];     // per-warp TMA ring barriers
};

// Warp-uniform insertion of `key` into the CTA-wide list of query `q` (rare path: ~LIST*ln(rows/LIST) times per CTA).
__device__ __noinline__ void list_insert(ScanShared* sh, int q, u64 key, int lane) {
    if (lane == 0) { while (atomicCAS(&sh->lock[q], 0, 1) != 0) {} }
    __syncwarp();
    __threadfence_block();
    volatile u64* lst = sh->list[q];
    u64 mine = lst[lane];
    u64 mn = warp_min_u64(mine);
    if (key > mn) {
        unsigned who = __ballot_sync(FULL, mine == mn);
        if (lane == __ffs(who) - 1) { lst[lane] = key; mine = key; }
        u64 mn2 = warp_min_u64(mine);
        if (lane == 0) *(volatile float*)&sh->thr[q] = (mn2 == 0ull) ? -CUDART_INF_F : unorder_f32((uint32_t)(mn2 >> 32));
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) atomicExch(&sh->lock[q], 0);
    __syncwarp();
}

enum { SCAN_SAMPLE = 0, SCAN_MAIN = 1, SCAN_LOCKED = 2 };
constexpr int CAND_CAP = 2048;     // survivors kept per query by the main scan
constexpr int KSEL = 32;           // rank of the threshold among the sample maxima

struct ScanArgs {
    u64* lists_out;        // LOCKED: [grid][QP][LIST]
    u64* maxima;           // SAMPLE: [QP][grid*SCAN_WARPS*32] per-lane maxima
    const u64* thr_key;    // MAIN: [QP] threshold keys
    u64* cand;             // MAIN: [QP][CAND_CAP]
    unsigned* cand_cnt;    // MAIN: [QP]
    const unsigned* overflow;   // LOCKED: per-query flags; a query group is re-scanned only if one of its flags is set
    int nq_total;          // LOCKED: queries of the whole pass (blockIdx.y = group of QP queries)
    int group_stride;      // SAMPLE: take every group_stride-th row group
};

// Rows reach the SM through per-warp shared-memory rings filled by 1-D TMA bulk copies (cp.async.bulk + mbarrier): one elected
// lane keeps NST groups of R rows (8 KB each) in flight per warp, decoupled from the registers that do the arithmetic.
// NST == 0: rows are loaded straight into registers with 16-byte streaming loads (fp32 rows: a ring stage would be 16 KB per warp).
template <typename T, int D, int QP, int R, int MODE, int NST>
__global__ void __launch_bounds__(SCAN_THREADS)
knn_scan_kernel(const T* __restrict__ db, const float* __restrict__ inv, long long n,
                const float* __restrict__ q, int nq_valid, ScanArgs args) {
    if (MODE == SCAN_LOCKED) {
        pdl_launch_dependents();                     // fallback pair: launched with programmatic serialisation behind the select kernel
        pdl_wait();
        // fallback pass: blockIdx.y selects a group of QP queries; it re-scans the database only if a query of the group overflowed
        const int grp = blockIdx.y;
        nq_valid = args.nq_total - grp * QP < QP ? args.nq_total - grp * QP : QP;
        unsigned any = 0u;
        for (int i = 0; i < nq_valid; i++) any |= args.overflow[grp * QP + i];
        if (any == 0u) return;
        q += (size_t)grp * QP * D;
        args.lists_out += (size_t)grp * gridDim.x * QP * LIST;
    }
    constexpr int STAGE_BYTES = R * D * (int)sizeof(T);
    constexpr int EPL = D / 32;              // elements per lane per row
    constexpr int EPV = Elem<T>::EPV;        // elements per 16-byte vector
    constexpr int NV = EPL / EPV;            // vectors per lane per row
    constexpr int NC = EPL / 4;              // float4 query chunks per lane
    static_assert(EPL % EPV == 0 && EPL % 4 == 0, "row width");
    extern __shared__ float4 smem_q[];       // [QP][NC][32], permuted so lane l reads consecutive float4s; then the row rings
    __shared__ ScanShared sh;
    uint8_t* ring = reinterpret_cast<uint8_t*>(smem_q + QP * NC * 32) + (size_t)(threadIdx.x >> 5) * (NST > 0 ? NST : 1) * STAGE_BYTES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < QP * NC * 32; i += SCAN_THREADS) {
        int l = i & 31, c = (i >> 5) % NC, qi = i / (32 * NC);
        float v[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            int j = 4 * c + t, vv = j / EPV, s = j % EPV;
            int e = vv * 32 * EPV + l * EPV + s;
            v[t] = qi < nq_valid ? q[(size_t)qi * D + e] : 0.f;
        }
        smem_q[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
    for (int i = tid; i < MAX_QP * LIST; i += SCAN_THREADS) sh.list[i / LIST][i % LIST] = 0ull;
    if (tid < MAX_QP) {
        sh.lock[tid] = 0;
        float t = -CUDART_INF_F;
        if (MODE == SCAN_MAIN && tid < QP) { u64 k = args.thr_key[tid]; t = k == 0ull ? -CUDART_INF_F : unorder_f32((uint32_t)(k >> 32)); }
        sh.thr[tid] = t;
    }
    if (NST > 0 && (tid & 31) == 0) {
        for (int s = 0; s < NST; s++) mbar_init(&sh.full[tid >> 5][s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    u64 best[QP];                                  // SAMPLE: running maximum key of the rows this lane owns
#pragma unroll
    for (int qi = 0; qi < QP; qi++) best[qi] = 0ull;
    const int gstride = MODE == SCAN_SAMPLE ? args.group_stride : 1;

    const int rsel = row_of_lane<R>(lane);
    const bool owner = (lane & (32 / R - 1)) == 0;
    const long long ngroups = (n + R - 1) / R;
    const long long nsteps = (ngroups + gstride - 1) / gstride;
    const long long gs0 = (long long)blockIdx.x * SCAN_WARPS + warp, gstep = (long long)gridDim.x * SCAN_WARPS;
    const long long nit = gs0 < nsteps ? (nsteps - gs0 + gstep - 1) / gstep : 0;
    uint64_t* bars = sh.full[warp];
    auto issue = [&](long long i) {                 // lane 0: one bulk copy of the (valid part of the) i-th row group of this warp
        const long long row0 = (gs0 + i * gstep) * gstride * R;
        const long long nrows = n - row0 < R ? n - row0 : R;
        const uint32_t bytes = (uint32_t)(nrows * D * (long long)sizeof(T));
        const int s = (int)(i % (NST > 0 ? NST : 1));
        mbar_expect_tx(&bars[s], bytes);
        bulk_load(ring + s * STAGE_BYTES, db + (size_t)row0 * D, bytes, &bars[s]);
    };
    if (NST > 0 && lane == 0) for (long long i = 0; i < NST && i < nit; i++) issue(i);
    for (long long it = 0; it < nit; it++) {
        const long long row0 = (gs0 + it * gstep) * gstride * R;
        uint4 raw[R][NV];
        if (NST > 0) {
            const int s = (int)(it % (NST > 0 ? NST : 1));
            mbar_wait(&bars[s], (uint32_t)((it / (NST > 0 ? NST : 1)) & 1));
#pragma unroll
            for (int r = 0; r < R; r++) {
                const uint4* p = reinterpret_cast<const uint4*>(ring + s * STAGE_BYTES + r * D * (int)sizeof(T));
#pragma unroll
                for (int v = 0; v < NV; v++) raw[r][v] = p[v * 32 + lane];
            }
            __syncwarp();                            // every lane has its rows in registers: the stage can be refilled
            if (lane == 0 && it + NST < nit) { fence_proxy_async(); issue(it + NST); }
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) {
                long long row = row0 + r; row = row < n ? row : n - 1;
                const uint4* p = reinterpret_cast<const uint4*>(db + (size_t)row * D);
#pragma unroll
                for (int v = 0; v < NV; v++) raw[r][v] = ldg_stream(p + v * 32 + lane);
            }
        }
        const long long myrow = row0 + rsel;
        const bool valid = owner && myrow < n;
        const float myinv = valid ? __ldg(inv + myrow) : 0.f;
        float x[R][EPL];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int v = 0; v < NV; v++) Elem<T>::unpack(raw[r][v], &x[r][v * EPV]);

#pragma unroll
        for (int qi = 0; qi < QP; qi++) {
            float acc[R];
#pragma unroll
            for (int r = 0; r < R; r++) acc[r] = 0.f;
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const float4 qv = smem_q[(qi * NC + c) * 32 + lane];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    acc[r] = fmaf(x[r][4 * c + 0], qv.x, acc[r]);
                    acc[r] = fmaf(x[r][4 * c + 1], qv.y, acc[r]);
                    acc[r] = fmaf(x[r][4 * c + 2], qv.z, acc[r]);
                    acc[r] = fmaf(x[r][4 * c + 3], qv.w, acc[r]);
                }
            }
            float s = reduce_rows<R>(acc, lane) * myinv;
            if (!(s == s)) s = -CUDART_INF_F;
            if (MODE == SCAN_SAMPLE) {
                if (valid) { u64 key = make_key(s, (uint32_t)myrow); best[qi] = key > best[qi] ? key : best[qi]; }
            } else if (MODE == SCAN_MAIN) {
                if (valid && qi < nq_valid && s >= sh.thr[qi]) {                 // rare: ~16*KSEL rows of the whole DB per query
                    const u64 key = make_key(s, (uint32_t)myrow);
                    if (key >= args.thr_key[qi]) {
                        unsigned pos = atomicAdd(&args.cand_cnt[qi], 1u);
                        if (pos < (unsigned)CAND_CAP) args.cand[(size_t)qi * CAND_CAP + pos] = key;
                    }
                }
            } else {
                const float thr = *(volatile float*)&sh.thr[qi];
                unsigned m = __ballot_sync(FULL, valid && qi < nq_valid && s >= thr);
                while (m) {
                    int src = __ffs(m) - 1; m &= m - 1;
                    float cs = __shfl_sync(FULL, s, src);
                    uint32_t ci = __shfl_sync(FULL, (uint32_t)myrow, src);
                    u64 key = make_key(cs, ci);
                    float tnow = __shfl_sync(FULL, *(volatile float*)&sh.thr[qi], 0);     // warp-uniform re-check
                    if (cs >= tnow) list_insert(&sh, qi, key, lane);
                }
            }
        }
    }
    if (MODE == SCAN_SAMPLE) {
        const size_t per_q = (size_t)gridDim.x * SCAN_THREADS;
#pragma unroll
        for (int qi = 0; qi < QP; qi++) args.maxima[qi * per_q + (size_t)blockIdx.x * SCAN_THREADS + tid] = best[qi];
    }
    if (MODE == SCAN_LOCKED) {
        __syncthreads();
        for (int i = tid; i < QP * LIST; i += SCAN_THREADS)
            args.lists_out[((size_t)blockIdx.x * QP + i / LIST) * LIST + (i % LIST)] = sh.list[i / LIST][i % LIST];
    }
}

// Tree merge of the 32 per-warp descending top-32 lists of a 1024-thread CTA (5 rounds instead of 31 serial merges); the result is
// warp 0's `cur`.  nlists: number of leading warps that hold a list (power of two).
__device__ __forceinline__ u64 cta_tree_merge(u64 cur, u64 (*s_keys)[LIST], int warp, int lane, int nlists) {
    s_keys[warp][lane] = cur;
    __syncthreads();
    for (int s = 1; s < nlists; s <<= 1) {
        if ((warp & (2 * s - 1)) == 0 && warp + s < nlists) {
            const u64 other = s_keys[warp + s][lane];
            if (__any_sync(FULL, other != 0ull)) cur = warp_merge_top32(cur, other, lane);
            s_keys[warp][lane] = cur;
        }
        __syncthreads();
    }
    return cur;
}

// threshold key[q] = (at least) the KSEL-th largest key of KSEL distinct sampled rows (0 if fewer exist); resets the counters.
// grid (queries, THR_P), block 1024.  Every thread first folds its strided share of the sample into ONE running maximum (a maximum
// over a group of rows is still the key of a real row, so the bound stays valid and is almost as tight: the top-32 sample rows fall into
// distinct groups with probability ~ 1 - 32^2 / (2 * 1024 * THR_P)); the CTA then takes the exact top-32 of its 1024 maxima (one bitonic
// sort per warp + a 5-round tree merge), and the last CTA of a query to finish merges the THR_P partial lists.
constexpr int THR_P = 8;
__global__ void __launch_bounds__(1024)
knn_threshold_kernel(const u64* __restrict__ maxima, size_t per_q, u64* __restrict__ thr_key, unsigned* __restrict__ cand_cnt, unsigned* __restrict__ overflow,
                     u64* __restrict__ part, unsigned* __restrict__ done) {
    __shared__ u64 s_keys[32][LIST];
    __shared__ unsigned s_last;
    const int qi = blockIdx.x, p = blockIdx.y, P = gridDim.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64* src = maxima + (size_t)qi * per_q;
    u64 best = 0ull;
    for (size_t i = (size_t)p * 1024 + tid; i < per_q; i += (size_t)P * 1024) { const u64 v = __ldcs(src + i); best = v > best ? v : best; }
    u64 cur = warp_merge_top32(0ull, best, lane);                      // sort this warp's 32 maxima, best first
    cur = cta_tree_merge(cur, s_keys, warp, lane, 32);
    if (warp == 0) {
        part[((size_t)qi * P + p) * LIST + lane] = cur;
        __threadfence();
        __syncwarp();
        if (lane == 0) s_last = atomicAdd(&done[qi], 1u) == (unsigned)(P - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    cur = warp < P ? __ldcg(part + ((size_t)qi * P + warp) * LIST + lane) : 0ull;
    cur = cta_tree_merge(cur, s_keys, warp, lane, P);
    if (warp == 0) {
        if (lane == KSEL - 1) thr_key[qi] = relax_key(cur);       // descending: lane 31 holds the 32nd largest (relaxed by the score slack)
        if (lane == 0) { cand_cnt[qi] = 0u; done[qi] = 0u; overflow[qi] = 0u; }
    }
}

__device__ __forceinline__ bool better_pair(double sa, long long ia, double sb, long long ib) { return sa > sb || (sa == sb && ia < ib); }

// grid = queries of this pass, block = 1024.  Merge the survivors to the 32 best fp32 keys, take f_k = the k-th of them, then re-rank
// EXACTLY (fp64, definition: oracle/knn_ref.c) every survivor whose fp32 score is within the slack of f_k -- at most one per thread -- and
// emit the top-k by (score desc, row asc): one warp-level bitonic sort when at most 32 survivors remain (the usual case), k rounds of a
// block-wide arg-best otherwise.
// FROM_LISTS = false: candidates of query qi are cand[qi*CAND_CAP .. + min(cnt, CAP)); sets overflow[qi] when cnt > CAP or when more than
//                     1024 survivors lie within the slack (a pathological cluster of near-identical rows: the fallback pass answers).
// FROM_LISTS = true : fallback, per-CTA lists [group][nblk][QP][LIST]; runs only for queries with overflow[qi] != 0 (every other query keeps
//                     the result of the main path); re-ranks the 32 best fp32 keys.
// Exact re-rank: a WARP fetches each candidate row with coalesced 16-byte loads (a thread walking its own row paid ~16 dependent DRAM round
// trips) and its 32 lanes leave the fp64 PRODUCTS q_j * d_j in shared memory -- every product is exact in fp64 (24 x 11 or 24 x 24
// significand bits), so computing them in parallel changes nothing; then one lane evaluates the oracle's SEQUENTIAL sum over them: 512
// dependent DADDs fed by plain shared-memory loads.  (Round 2, measured with ncu: with the unpacking, the two F2F conversions and the DMUL
// inside that one lane's loop the sum cost ~35 cycles per term and 44 % of the kernel's stall samples sat at the barrier behind it.)
constexpr int SEL_MAX = 1024;
template <typename T, int D> struct SelCfg {
    static constexpr int ROW_BYTES = D * (int)sizeof(T);
    static constexpr int ROWS_PAR = 16;                                                    // candidates re-ranked per round (one per warp)
    static constexpr int DYN_BYTES = ROWS_PAR * D * (int)sizeof(double);                   // their fp64 product rows
};
template <typename T, int D, bool FROM_LISTS>
__global__ void __launch_bounds__(1024)
knn_select_kernel(const u64* __restrict__ lists, int nblk, int QP, const unsigned* __restrict__ cand_cnt, unsigned* __restrict__ overflow,
                  const T* __restrict__ db, const float* __restrict__ inv,
                  long long n, const float* __restrict__ q, int k, long long idx_base,
                  long long* __restrict__ idx_out, float* __restrict__ dist_out, double* __restrict__ score_out, float slack) {
    __shared__ u64 s_keys[32][LIST];
    __shared__ float s_q[D];
    __shared__ u64 s_cand[SEL_MAX];
    __shared__ double s_score[SEL_MAX];
    extern __shared__ double s_prod[];                     // [ROWS_PAR][D]
    __shared__ u64 s_cut;
    __shared__ unsigned s_ncand;
    __shared__ double s_bs[32];
    __shared__ long long s_bi[32];
    __shared__ long long s_win;
    const int qi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();                         // the kernels of a search are chained with programmatic serialisation: the next one is
    pdl_wait();                                      // staged while this one runs and waits here for its predecessor's results
    if (FROM_LISTS) { if (overflow[qi] == 0u) return; }
    for (int i = tid; i < D; i += blockDim.x) s_q[i] = q[(size_t)qi * D + i];
    if (tid == 0) s_ncand = 0u;
    u64 cur = 0ull;
    unsigned cnt = 0u;
    if (FROM_LISTS) {
        const int grp = qi / QP, ql = qi - grp * QP;
        for (int b = warp; b < nblk; b += 32) {
            u64 batch = lists[(((size_t)grp * nblk + b) * QP + ql) * LIST + lane];
            if (__any_sync(FULL, batch != 0ull)) cur = warp_merge_top32(cur, batch, lane);
        }
    } else {
        cnt = cand_cnt[qi];
        if (cnt > (unsigned)CAND_CAP) { if (tid == 0) overflow[qi] = 1u; return; }            // the fallback pass redoes this query
        for (unsigned b = warp * 32; b < cnt; b += 32 * 32) {
            u64 batch = b + lane < cnt ? lists[(size_t)qi * CAND_CAP + b + lane] : 0ull;
            cur = warp_merge_top32(cur, batch, lane);
        }
    }
    cur = cta_tree_merge(cur, s_keys, warp, lane, 32);          // warp 0: lane i holds the i-th largest fp32 key
    if (warp == 0) {
        const u64 kth = __shfl_sync(FULL, cur, k - 1);           // 0 when fewer than k survivors exist: keep everything
        if (lane == 0) s_cut = relax_key(kth, slack);            // slack of the scan that produced the candidates (hi-only tensor-core passes: 1.05e-3)
        if (FROM_LISTS) { s_cand[lane] = cur; if (lane == 0) s_ncand = 32u; }
    }
    __syncthreads();
    if (!FROM_LISTS) {
        const u64 cut = s_cut;
        for (unsigned i = tid; i < cnt; i += blockDim.x) {
            const u64 key = lists[(size_t)qi * CAND_CAP + i];
            if (key != 0ull && key >= cut) { const unsigned pos = atomicAdd(&s_ncand, 1u); if (pos < (unsigned)SEL_MAX) s_cand[pos] = key; }
        }
        __syncthreads();
        if (s_ncand > (unsigned)SEL_MAX) { if (tid == 0) overflow[qi] = 1u; return; }
    }
    // exact re-rank: warp w stages candidate base + w, lane 0 sums it in the oracle's order; afterwards thread t owns candidate t
    {
        constexpr int RP = SelCfg<T, D>::ROWS_PAR, V = SelCfg<T, D>::ROW_BYTES / 16, EPV = Elem<T>::EPV;
        const int ncand = (int)s_ncand;
        for (int base = 0; base < ncand; base += RP) {
            const int c = base + warp;
            if (warp < RP && c < ncand) {
                const u64 key = s_cand[c];
                const long long r = key != 0ull ? (long long)(0xffffffffu - (uint32_t)key) : n;
                double acc = -CUDART_INF;
                if (r < n) {
                    const uint4* src = reinterpret_cast<const uint4*>(db + (size_t)r * D);
                    double* prod = s_prod + (size_t)warp * D;
                    for (int v = lane; v < V; v += 32) {
                        float f[EPV];
                        Elem<T>::unpack(ldg_stream(src + v), f);
#pragma unroll
                        for (int e = 0; e < EPV; e++) prod[v * EPV + e] = __dmul_rn((double)s_q[v * EPV + e], (double)f[e]);      // exact
                    }
                    __syncwarp();
                    if (lane == 0) {
                        acc = 0.0;
#pragma unroll 8
                        for (int j = 0; j < D; j++) acc = __dadd_rn(acc, prod[j]);                                                 // same order as the oracle
                        acc = __dmul_rn(acc, (double)inv[r]);
                        if (!(acc == acc)) acc = -CUDART_INF;
                    }
                    __syncwarp();
                }
                if (lane == 0) s_score[c] = acc;
            }
        }
        __syncthreads();
    }
    const u64 mykey = tid < (int)s_ncand ? s_cand[tid] : 0ull;
    const bool have = mykey != 0ull;
    long long row = have ? (long long)(0xffffffffu - (uint32_t)mykey) : 0x7fffffffffffffffLL;
    double s = -CUDART_INF;
    if (have && row < n) s = s_score[tid];
    else if (have) row = 0x7fffffffffffffffLL;
    if (s_ncand <= 32u) {
        // the usual case (and always the fallback pass): all candidates sit in warp 0 -- one bitonic sort of (score, row) pairs, best first
        if (warp != 0) return;
#pragma unroll
        for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
            for (int j = kk >> 1; j > 0; j >>= 1) {
                const double os = __shfl_xor_sync(FULL, s, j);
                const long long orow = __shfl_xor_sync(FULL, row, j);
                const bool desc = (lane & kk) == 0, lower = (lane & j) == 0;
                const bool keep_best = (desc == lower);
                const bool other_better = better_pair(os, orow, s, row);
                if (keep_best == other_better) { s = os; row = orow; }
            }
        }
        if (lane < k) {
            const bool ok = row != 0x7fffffffffffffffLL;
            idx_out[(size_t)qi * k + lane] = ok ? row + idx_base : -1;
            dist_out[(size_t)qi * k + lane] = ok ? (float)s : -CUDART_INF_F;
            if (score_out) score_out[(size_t)qi * k + lane] = ok ? s : -CUDART_INF;
        }
        return;
    }
    // more than 32 survivors within the slack: k rounds of a block-wide arg-best over the candidates not emitted yet
    bool taken = row == 0x7fffffffffffffffLL;
    for (int r = 0; r < k; r++) {
        double bs = taken ? -CUDART_INF : s;
        long long bi = taken ? 0x7fffffffffffffffLL : row;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const double os = __shfl_xor_sync(FULL, bs, m);
            const long long oi = __shfl_xor_sync(FULL, bi, m);
            if (oi != 0x7fffffffffffffffLL && (bi == 0x7fffffffffffffffLL || better_pair(os, oi, bs, bi))) { bs = os; bi = oi; }
        }
        if (lane == 0) { s_bs[warp] = bs; s_bi[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            bs = s_bs[lane]; bi = s_bi[lane];
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                const double os = __shfl_xor_sync(FULL, bs, m);
                const long long oi = __shfl_xor_sync(FULL, bi, m);
                if (oi != 0x7fffffffffffffffLL && (bi == 0x7fffffffffffffffLL || better_pair(os, oi, bs, bi))) { bs = os; bi = oi; }
            }
            if (lane == 0) {
                const bool ok = bi != 0x7fffffffffffffffLL;
                idx_out[(size_t)qi * k + r] = ok ? bi + idx_base : -1;
                dist_out[(size_t)qi * k + r] = ok ? (float)bs : -CUDART_INF_F;
                if (score_out) score_out[(size_t)qi * k + r] = ok ? bs : -CUDART_INF;
                s_win = bi;
            }
        }
        __syncthreads();
        if (!taken && row == s_win) taken = true;
        __syncthreads();                                          // s_bs / s_bi / s_win are rewritten by the next round
    }
}

template <typename T, int D>
__global__ void knn_inv_norm_kernel(const T* __restrict__ db, long long n, float* __restrict__ inv) {
    long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    constexpr int EPV = Elem<T>::EPV;
    const uint4* p = reinterpret_cast<const uint4*>(db + (size_t)row * D);
    double s = 0.0;
    for (int v = 0; v < D / EPV; v++) {
        float f[EPV];
        Elem<T>::unpack(__ldg(p + v), f);
#pragma unroll
        for (int j = 0; j < EPV; j++) { double e = (double)f[j]; s = __dadd_rn(s, __dmul_rn(e, e)); }
    }
    inv[row] = (float)(1.0 / sqrt(s));
}

template <typename T, int D>
__global__ void knn_gather_kernel(const T* __restrict__ db, long long n, long long idx_base, const long long* __restrict__ idx,
                                  long long count, float* __restrict__ out) {
    constexpr int EPV = Elem<T>::EPV;
    long long i = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (i >= count) return;
    const int lane = threadIdx.x & 31;
    long long row = idx[i] - idx_base;
    const bool inside = row >= 0 && row < n;
    float* o = out + (size_t)i * D;
    for (int v = lane; v < D / EPV; v += 32) {
        float f[EPV];
        if (inside) Elem<T>::unpack(__ldg(reinterpret_cast<const uint4*>(db + (size_t)row * D) + v), f);
        else {
#pragma unroll
            for (int j = 0; j < EPV; j++) f[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < EPV; j += 4) *reinterpret_cast<float4*>(o + v * EPV + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    }
}


// q_hat = q / ||q||_2 row-wise in fp32, bit-identical to the reference's NumPy statement `q / np.linalg.norm(q, axis=1)[:, np.newaxis]`
// (ddpm.py:297,907; dsetbuilder.py:487; base.py:82) for float32 rows: the squares are rounded separately, summed in NumPy's pairwise order
// (oracle/knn.py: pairwise_sum_f32 -- leaves of L = 128 (96 for d = 768) elements with 8 strided accumulators combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), leaves combined by a balanced binary tree), then an IEEE square root and an IEEE division.
// One CTA of 64 threads per query: thread t owns accumulator t % 8 of leaf t / 8; fp32 addition is commutative, so the xor-shuffle tree
// reproduces the association above exactly.
template <int D>
__global__ void __launch_bounds__(64) knn_normalize_kernel(const float* __restrict__ q, int nq, float* __restrict__ out,
                                                          __half* __restrict__ split, int NQ, unsigned* __restrict__ grid_bar) {
    constexpr int NLEAF = D == 768 ? 8 : D / 128, L = D / NLEAF;
    static_assert(NLEAF == 2 || NLEAF == 4 || NLEAF == 8, "leaf tree");
    __shared__ float s_w[2];
    const int qi = blockIdx.x, t = threadIdx.x, leaf = t >> 3, j = t & 7;
    // split != null (tensor-core scan follows): also emit the fp16 hi / lo rows of q_hat (rows [0, NQ) = hi, [NQ, 2 NQ) = lo, zero beyond nq)
    // and reset the grid-barrier counter of the fused scan -- the separate query-split launch disappears
    if (qi == 0 && t == 0 && grid_bar) *grid_bar = 0u;
    if (qi >= nq) {
        for (int i = t; i < D; i += 64) { split[(size_t)qi * D + i] = __float2half_rn(0.f); split[(size_t)(NQ + qi) * D + i] = __float2half_rn(0.f); }
        return;
    }
    const float* x = q + (size_t)qi * D;
    float r = 0.f;
    if (leaf < NLEAF) {
        const float* a = x + leaf * L + j;
        r = __fmul_rn(a[0], a[0]);
#pragma unroll 4
        for (int i = 8; i < L; i += 8) r = __fadd_rn(r, __fmul_rn(a[i], a[i]));
    }
    r = __fadd_rn(r, __shfl_xor_sync(FULL, r, 1));
    r = __fadd_rn(r, __shfl_xor_sync(FULL, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(FULL, r, 4));                  // leaf sums
    if (NLEAF >= 2) r = __fadd_rn(r, __shfl_xor_sync(FULL, r, 8));
    if (NLEAF >= 4) r = __fadd_rn(r, __shfl_xor_sync(FULL, r, 16));
    if ((t & 31) == 0) s_w[t >> 5] = r;                             // thread 0 holds the sum of leaves 0..3, thread 32 that of leaves 4..7
    __syncthreads();
    r = NLEAF == 8 ? __fadd_rn(s_w[0], s_w[1]) : s_w[0];
    r = __fadd_rn(0.f, r);                                          // add.reduce starts from the identity 0
    const float nrm = __fsqrt_rn(r);
    for (int i = t; i < D; i += 64) {
        const float v = __fdiv_rn(x[i], nrm);
        out[(size_t)qi * D + i] = v;
        if (split) {
            const __half hh = __float2half_rn(v);
            split[(size_t)qi * D + i] = hh;
            split[(size_t)(NQ + qi) * D + i] = __float2half_rn(v - __half2float(hh));
        }
    }
}

// one warp per query: k rounds of arg-best over parts*k candidates
__global__ void knn_merge_kernel(const long long* __restrict__ idx_in, const double* __restrict__ sc_in, int parts, int nq, int k,
                                 long long* __restrict__ idx_out, float* __restrict__ dist_out, double* __restrict__ sc_out) {
    const int qi = blockIdx.x, lane = threadIdx.x;
    const int total = parts * k;
    double last_s = CUDART_INF; long long last_i = -1;      // everything strictly worse than (last_s, last_i) is still available
    for (int t = 0; t < k; t++) {
        double bs = -CUDART_INF; long long bi = 0x7fffffffffffffffLL;
        for (int c = lane; c < total; c += 32) {
            int p = c / k, j = c % k;
            long long ci = idx_in[((size_t)p * nq + qi) * k + j];
            double cs = sc_in[((size_t)p * nq + qi) * k + j];
            if (ci < 0) continue;
            if (t > 0 && !better_pair(last_s, last_i, cs, ci)) continue;     // already emitted (or equal to it)
            if (better_pair(cs, ci, bs, bi)) { bs = cs; bi = ci; }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            double os = __shfl_xor_sync(FULL, bs, m); long long oi = __shfl_xor_sync(FULL, bi, m);
            if (better_pair(os, oi, bs, bi)) { bs = os; bi = oi; }
        }
        bool ok = bi != 0x7fffffffffffffffLL;
        if (lane == 0) {
            idx_out[(size_t)qi * k + t] = ok ? bi : -1;
            dist_out[(size_t)qi * k + t] = ok ? (float)bs : -CUDART_INF_F;
            if (sc_out) sc_out[(size_t)qi * k + t] = ok ? bs : -CUDART_INF;
        }
        last_s = bs; last_i = bi;
        if (!ok) { last_s = -CUDART_INF; last_i = 0x7fffffffffffffffLL; }
    }
}

}  // namespace

struct rdm_knn {
    int device = 0;
    long long n = 0;
    int d = 0, dtype = 0;
    long long idx_base = 0;
    const void* db = nullptr;      // device
    void* db_owned = nullptr;      // device allocation when copied from host
    float* inv = nullptr;
    u64* lists = nullptr;      // fallback per-CTA lists
    u64* maxima = nullptr;     // sample keys: [queries][per_q]
    size_t maxima_keys = 0;
    u64* cand = nullptr;       // [MAX_QP][CAND_CAP]
    u64* thr_key = nullptr;    // [MAX_QP]
    unsigned* cand_cnt = nullptr;   // [MAX_TCQ] survivor counters, then [MAX_TCQ] per-query overflow flags
    u64* thr_part = nullptr;        // threshold kernel: [MAX_TCQ][THR_P][LIST] partial top-32 lists
    unsigned* thr_done = nullptr;   // [MAX_TCQ] arrival counters (zero between searches)
    void* qsplit = nullptr;         // fp16 hi/lo query rows for the tensor-core scan
    float* qhat = nullptr;          // rdm_knn_search_raw: normalised queries [QHAT_ROWS][d]
    void* fused_ws = nullptr;       // fused tensor-core scan: group maxima + grid-barrier counter
    int max_grid = 0;
};

namespace {

// launch with programmatic stream serialisation: the kernel may be staged while its predecessor in the stream still runs; it calls
// griddepcontrol.wait before touching the predecessor's results (RDM_KNN_NO_PDL=1: ordinary stream order)
template <typename... KArgs, typename... Args>
cudaError_t launch_chained(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    static const int no_pdl = getenv("RDM_KNN_NO_PDL") ? 1 : 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

template <typename T, int D, int QP, int R, int MODE>
int launch_scan(rdm_knn* h, const float* q, int nq_valid, ScanArgs args, cudaStream_t st, int* grid_out, int groups = 1) {
    constexpr int STAGE = R * D * (int)sizeof(T);
    constexpr int QBYTES = QP * D * (int)sizeof(float);
    // 2-byte rows: TMA rings (3 stages if they fit next to the queries in 227 KB, else 2); 4-byte rows: direct loads
    constexpr int NST = sizeof(T) == 4 ? 0 : ((QBYTES + 3 * SCAN_WARPS * STAGE + 8192 <= 232448) ? 3 : 2);
    static_assert(QBYTES + NST * SCAN_WARPS * STAGE + 8192 <= 232448, "scan shared-memory budget");
    auto kern = knn_scan_kernel<T, D, QP, R, MODE, NST>;
    size_t smem = (size_t)QBYTES + (size_t)NST * SCAN_WARPS * STAGE;
    static thread_local int cached_grid[8] = {0};   // per device
    int& grid = cached_grid[h->device & 7];
    if (grid == 0) {
        RDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        RDM_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SCAN_THREADS, smem));
        if (per_sm < 1) per_sm = 1;
        grid = rdm_num_sms(h->device) * per_sm;
        if (grid > h->max_grid) grid = h->max_grid;
    }
    const long long ngroups = (h->n + R - 1) / R;
    const long long nsteps = (ngroups + args.group_stride - 1) / args.group_stride;
    int g = grid;
    long long need = (nsteps + SCAN_WARPS - 1) / SCAN_WARPS;
    if (need < g) g = (int)(need < 1 ? 1 : need);
    if (MODE == SCAN_LOCKED) {
        RDM_CHECK_CUDA(launch_chained(kern, dim3(g, groups), dim3(SCAN_THREADS), smem, st, (const T*)h->db, (const float*)h->inv, (long long)h->n, q, nq_valid, args));
    } else {
        kern<<<dim3(g, groups), SCAN_THREADS, smem, st>>>((const T*)h->db, h->inv, h->n, q, nq_valid, args);
    }
    RDM_COUNT_LAUNCH();
    RDM_CHECK_CUDA(cudaGetLastError());
    *grid_out = g;
    return RDM_OK;
}

template <typename T, int D, bool FROM_LISTS>
int launch_select(rdm_knn* h, int nq, const u64* lists, int nblk, int QP, const float* qp, int k, long long* idx_out, float* dist_out, double* sc_out, cudaStream_t st, float slack = SCORE_SLACK) {
    auto kern = knn_select_kernel<T, D, FROM_LISTS>;
    static bool configured[16] = {false};
    if (!configured[h->device & 15]) {
        RDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SelCfg<T, D>::DYN_BYTES));
        configured[h->device & 15] = true;
    }
    unsigned* overflow = h->cand_cnt + MAX_TCQ;
    RDM_CHECK_CUDA(launch_chained(kern, dim3(nq), dim3(1024), (size_t)SelCfg<T, D>::DYN_BYTES, st, lists, nblk, QP, (const unsigned*)h->cand_cnt, overflow, (const T*)h->db,
                                  (const float*)h->inv, (long long)h->n, qp, k, (long long)h->idx_base, idx_out, dist_out, sc_out, slack));
    RDM_COUNT_LAUNCH();
    RDM_CHECK_CUDA(cudaGetLastError());
    return RDM_OK;
}

template <typename T, int D, int QP, int R>
int search_pass(rdm_knn* h, const float* qp, int cnt, int k, long long* idx_out, float* dist_out, double* sc_out, cudaStream_t st) {
    unsigned* overflow = h->cand_cnt + MAX_TCQ;
    const long long ngroups = (h->n + R - 1) / R;
    // sample ~1/16 of the row groups, but keep at least ~8K groups in the sample (small DBs are sampled densely)
    int stride = (int)(ngroups / 8192); if (stride > 16) stride = 16; if (stride < 1) stride = 1;
    int g0 = 0, g1 = 0, g2 = 0;
    ScanArgs a{}; a.group_stride = stride; a.maxima = h->maxima;
    RDM_TRY((launch_scan<T, D, QP, R, SCAN_SAMPLE>(h, qp, cnt, a, st, &g0)));
    knn_threshold_kernel<<<dim3(QP, THR_P), 1024, 0, st>>>(h->maxima, (size_t)g0 * SCAN_THREADS, h->thr_key, h->cand_cnt, overflow, h->thr_part, h->thr_done);
    RDM_COUNT_LAUNCH();
    ScanArgs m{}; m.group_stride = 1; m.thr_key = h->thr_key; m.cand = h->cand; m.cand_cnt = h->cand_cnt;
    RDM_TRY((launch_scan<T, D, QP, R, SCAN_MAIN>(h, qp, cnt, m, st, &g1)));
    RDM_TRY((launch_select<T, D, false>(h, cnt, h->cand, 0, QP, qp, k, idx_out, dist_out, sc_out, st)));
    // device-side conditional fallback (both kernels return immediately unless a candidate buffer of one of the queries overflowed)
    ScanArgs f{}; f.group_stride = 1; f.lists_out = h->lists; f.overflow = overflow; f.nq_total = cnt;
    RDM_TRY((launch_scan<T, D, QP, R, SCAN_LOCKED>(h, qp, cnt, f, st, &g2)));
    RDM_TRY((launch_select<T, D, true>(h, cnt, h->lists, g2, QP, qp, k, idx_out, dist_out, sc_out, st)));
    return RDM_OK;
}

// fp16 / d=512, >= 8 queries: sample + main scans on the tensor cores, up to 64 queries per pass.
template <typename T, int D>
int search_pass_tc(rdm_knn* h, const float* qp, int cnt, int k, long long* idx_out, float* dist_out, double* sc_out, cudaStream_t st, int presplit = 0) {
    constexpr int R = 8;
    unsigned* overflow = h->cand_cnt + MAX_TCQ;
    // one cooperative launch (sample phase -> grid barrier -> thresholds -> main scan) where the database is large enough for it;
    // otherwise (or with RDM_KNN_NO_FUSED) the three-kernel sequence: sample scan, threshold kernel, main scan
    static const bool no_fused = getenv("RDM_KNN_NO_FUSED") != nullptr;
    float slack = SCORE_SLACK;
    int fused = no_fused ? 1 : knn_scan_tc_fused(h->db, h->inv, h->n, h->device, qp, cnt, k, h->qsplit, h->cand, h->cand_cnt, overflow, h->fused_ws, presplit, &slack, st);
    if (fused < 0) return fused;
    if (fused != RDM_OK && cnt > 64) {                    // the three-kernel path takes at most 64 queries: two half passes (they split the queries themselves)
        const int half = 64;
        RDM_TRY((search_pass_tc<T, D>(h, qp, half, k, idx_out, dist_out, sc_out, st, 0)));
        return search_pass_tc<T, D>(h, qp + (size_t)half * D, cnt - half, k, idx_out + (size_t)half * k, dist_out + (size_t)half * k, sc_out ? sc_out + (size_t)half * k : nullptr, st, 0);
    }
    if (fused != RDM_OK) {
        slack = SCORE_SLACK;
        const long long ntiles = (h->n + 127) / 128;
        int stride = (int)(ntiles / 512); if (stride > 16) stride = 16; if (stride < 1) stride = 1;       // >= ~64K sampled rows
        const long long per_q = knn_tc_sample_rows(h->n, stride);
        RDM_REQUIRE((size_t)per_q * cnt <= h->maxima_keys, RDM_ERR_STATE, "knn: sample buffer too small");
        RDM_TRY(knn_scan_tc(h->db, h->inv, h->n, h->device, qp, cnt, h->qsplit, 1, stride, h->maxima, per_q, nullptr, nullptr, nullptr, st));
        knn_threshold_kernel<<<dim3(cnt, THR_P), 1024, 0, st>>>(h->maxima, (size_t)per_q, h->thr_key, h->cand_cnt, overflow, h->thr_part, h->thr_done);
        RDM_COUNT_LAUNCH();
        RDM_TRY(knn_scan_tc(h->db, h->inv, h->n, h->device, qp, cnt, h->qsplit, 0, 1, nullptr, 0, h->thr_key, h->cand, h->cand_cnt, st));
    }
    RDM_TRY((launch_select<T, D, false>(h, cnt, h->cand, 0, MAX_QP, qp, k, idx_out, dist_out, sc_out, st, slack)));
    // device-side conditional fallback: ONE launch pair for all groups of 16 queries (blockIdx.y = group); normally two empty launches
    int g2 = 0;
    ScanArgs f{}; f.group_stride = 1; f.lists_out = h->lists; f.overflow = overflow; f.nq_total = cnt;
    RDM_TRY((launch_scan<T, D, MAX_QP, R, SCAN_LOCKED>(h, qp, cnt, f, st, &g2, (cnt + MAX_QP - 1) / MAX_QP)));
    RDM_TRY((launch_select<T, D, true>(h, cnt, h->lists, g2, MAX_QP, qp, k, idx_out, dist_out, sc_out, st)));
    return RDM_OK;
}

// fp32 / d=512 databases, >= 3 queries: the fused scan on kind::tf32 MMAs (knn_tc.cu), up to 64 queries per pass.  Returns 1 (nothing
// launched) when the database is too small for the fused scan: the caller then takes the CUDA-core passes.
template <typename T, int D>
int search_pass_tf32(rdm_knn* h, const float* qp, int cnt, int k, long long* idx_out, float* dist_out, double* sc_out, cudaStream_t st) {
    constexpr int R = 8;
    unsigned* overflow = h->cand_cnt + MAX_TCQ;
    float slack = SCORE_SLACK;
    const int fused = knn_scan_tc_fused_f32(h->db, h->inv, h->n, h->device, qp, cnt, k, h->qsplit, h->cand, h->cand_cnt, overflow, h->fused_ws, &slack, st);
    if (fused != RDM_OK) return fused;
    RDM_TRY((launch_select<T, D, false>(h, cnt, h->cand, 0, MAX_QP, qp, k, idx_out, dist_out, sc_out, st, slack)));
    int g2 = 0;
    ScanArgs f{}; f.group_stride = 1; f.lists_out = h->lists; f.overflow = overflow; f.nq_total = cnt;
    RDM_TRY((launch_scan<T, D, MAX_QP, R, SCAN_LOCKED>(h, qp, cnt, f, st, &g2, (cnt + MAX_QP - 1) / MAX_QP)));
    RDM_TRY((launch_select<T, D, true>(h, cnt, h->lists, g2, MAX_QP, qp, k, idx_out, dist_out, sc_out, st)));
    return RDM_OK;
}

template <typename T, int D>
int search_typed(rdm_knn* h, const float* q, int nq, int k, long long* idx_out, float* dist_out, double* sc_out, cudaStream_t st) {
    if constexpr (std::is_same<T, float>::value && D == 512) {
        static const bool no_tc = getenv("RDM_KNN_NO_TC") != nullptr;
        if (!no_tc && nq >= 3) {
            bool done = true;
            for (int q0 = 0; q0 < nq && done; q0 += 64) {
                const int cnt = nq - q0 < 64 ? nq - q0 : 64;
                const int rc = search_pass_tf32<T, D>(h, q + (size_t)q0 * D, cnt, k, idx_out + (size_t)q0 * k, dist_out + (size_t)q0 * k, sc_out ? sc_out + (size_t)q0 * k : nullptr, st);
                if (rc < 0) return rc;
                if (rc != RDM_OK) done = false;          // (only the first pass can say "not usable here": nothing was launched)
            }
            if (done) return RDM_OK;
        }
    }
    if constexpr (std::is_same<T, __half>::value && D == 512) {
        static const bool no_tc = getenv("RDM_KNN_NO_TC") != nullptr;
        if (!no_tc && nq >= 3) {          // measured: from 3 queries on, the (padded) tensor-core pass beats the FMA scan (0.28 vs 0.35 ms at 4 queries, 1.28 M rows)
            for (int q0 = 0; q0 < nq; q0 += MAX_TCQ) {
                int cnt = nq - q0 < MAX_TCQ ? nq - q0 : MAX_TCQ;
                RDM_TRY((search_pass_tc<T, D>(h, q + (size_t)q0 * D, cnt, k, idx_out + (size_t)q0 * k, dist_out + (size_t)q0 * k, sc_out ? sc_out + (size_t)q0 * k : nullptr, st)));
            }
            return RDM_OK;
        }
    }
    // rows per group: <= 8 KB per ring stage and <= 128 registers of row data per lane
    constexpr int RB = sizeof(T) == 4 ? 8 : 8192 / (D * (int)sizeof(T)), RR = 128 / (D / 32);
    constexpr int R = (RB >= 8 && RR >= 8) ? 8 : (RB >= 4 && RR >= 4) ? 4 : 2;
    for (int q0 = 0; q0 < nq; q0 += MAX_QP) {
        int cnt = nq - q0 < MAX_QP ? nq - q0 : MAX_QP;
        const float* qp = q + (size_t)q0 * D;
        long long* io = idx_out + (size_t)q0 * k; float* dist_o = dist_out + (size_t)q0 * k; double* so = sc_out ? sc_out + (size_t)q0 * k : nullptr;
        if (cnt <= 1)      RDM_TRY((search_pass<T, D, 1, R>(h, qp, cnt, k, io, dist_o, so, st)));
        else if (cnt <= 2) RDM_TRY((search_pass<T, D, 2, R>(h, qp, cnt, k, io, dist_o, so, st)));
        else if (cnt <= 4) RDM_TRY((search_pass<T, D, 4, R>(h, qp, cnt, k, io, dist_o, so, st)));
        else if (cnt <= 8) RDM_TRY((search_pass<T, D, 8, R>(h, qp, cnt, k, io, dist_o, so, st)));
        else               RDM_TRY((search_pass<T, D, 16, R>(h, qp, cnt, k, io, dist_o, so, st)));
    }
    return RDM_OK;
}

template <typename T, int D>
int create_typed(rdm_knn* h) {
    int blocks = (int)((h->n + 127) / 128);
    knn_inv_norm_kernel<T, D><<<blocks, 128>>>((const T*)h->db, h->n, h->inv);
    RDM_COUNT_LAUNCH();
    RDM_CHECK_CUDA(cudaGetLastError());
    RDM_CHECK_CUDA(cudaDeviceSynchronize());
    return RDM_OK;
}

template <typename T, int D>
int gather_typed(rdm_knn* h, const long long* idx, long long count, float* out, cudaStream_t st) {
    if (count == 0) return RDM_OK;
    int blocks = (int)((count + 7) / 8);
    knn_gather_kernel<T, D><<<blocks, 256, 0, st>>>((const T*)h->db, h->n, h->idx_base, idx, count, out);
    RDM_COUNT_LAUNCH();
    RDM_CHECK_CUDA(cudaGetLastError());
    return RDM_OK;
}

#define KNN_DISPATCH(h, FN, ...)                                                                    \
    do {                                                                                            \
        if ((h)->dtype == RDM_DTYPE_F16) {                                                          \
            switch ((h)->d) {                                                                       \
                case 256: return FN<__half, 256>(__VA_ARGS__);                                      \
                case 512: return FN<__half, 512>(__VA_ARGS__);                                      \
                case 768: return FN<__half, 768>(__VA_ARGS__);                                      \
                case 1024: return FN<__half, 1024>(__VA_ARGS__);                                    \
            }                                                                                       \
        } else {                                                                                    \
            switch ((h)->d) {                                                                       \
                case 256: return FN<float, 256>(__VA_ARGS__);                                       \
                case 512: return FN<float, 512>(__VA_ARGS__);                                       \
                case 768: return FN<float, 768>(__VA_ARGS__);                                       \
                case 1024: return FN<float, 1024>(__VA_ARGS__);                                     \
            }                                                                                       \
        }                                                                                           \
        rdm_set_error("knn: unsupported d=%d", (h)->d);                                             \
        return RDM_ERR_UNSUPPORTED;                                                                 \
    } while (0)

int do_create(rdm_knn* h) { KNN_DISPATCH(h, create_typed, h); }
int do_search(rdm_knn* h, const float* q, int nq, int k, long long* i, float* d, double* s, cudaStream_t st) {
    KNN_DISPATCH(h, search_typed, h, q, nq, k, i, d, s, st);
}
int do_gather(rdm_knn* h, const long long* idx, long long count, float* out, cudaStream_t st) {
    KNN_DISPATCH(h, gather_typed, h, idx, count, out, st);
}

int normalize_rows(const float* q, int nq, int d, float* out, cudaStream_t st, __half* split = nullptr, int NQ = 0, unsigned* grid_bar = nullptr) {
    if (nq == 0) return RDM_OK;
    const int grid = split && NQ > nq ? NQ : nq;
    switch (d) {
        case 256: knn_normalize_kernel<256><<<grid, 64, 0, st>>>(q, nq, out, split, NQ, grid_bar); break;
        case 512: knn_normalize_kernel<512><<<grid, 64, 0, st>>>(q, nq, out, split, NQ, grid_bar); break;
        case 768: knn_normalize_kernel<768><<<grid, 64, 0, st>>>(q, nq, out, split, NQ, grid_bar); break;
        case 1024: knn_normalize_kernel<1024><<<grid, 64, 0, st>>>(q, nq, out, split, NQ, grid_bar); break;
        default: rdm_set_error("knn normalize: unsupported d=%d", d); return RDM_ERR_UNSUPPORTED;
    }
    RDM_COUNT_LAUNCH();
    RDM_CHECK_CUDA(cudaGetLastError());
    return RDM_OK;
}

}  // namespace

extern "C" {

int rdm_knn_create(rdm_knn_t** out, const void* db, int64_t n, int32_t d, int32_t dtype, int32_t db_on_device,
                   int64_t idx_base, int32_t device) {
    RDM_REQUIRE(out && db, RDM_ERR_ARG, "rdm_knn_create: null argument");
    RDM_REQUIRE(n > 0 && n < 0xfffffff0LL, RDM_ERR_ARG, "rdm_knn_create: n=%lld out of range (1..2^32-16 rows per shard)", (long long)n);
    RDM_REQUIRE(d == 256 || d == 512 || d == 768 || d == 1024, RDM_ERR_UNSUPPORTED, "rdm_knn_create: d=%d not in {256,512,768,1024}", d);
    RDM_REQUIRE(dtype == RDM_DTYPE_F16 || dtype == RDM_DTYPE_F32, RDM_ERR_ARG, "rdm_knn_create: bad dtype %d", dtype);
    DeviceGuard guard(device);
    RDM_REQUIRE(guard.ok, RDM_ERR_CUDA, "rdm_knn_create: cannot select device %d", device);
    rdm_knn* h = new rdm_knn();
    h->device = device; h->n = n; h->d = d; h->dtype = dtype; h->idx_base = idx_base;
    size_t bytes = (size_t)n * d * (dtype == RDM_DTYPE_F16 ? 2 : 4);
    int rc = RDM_OK;
    do {
        if (db_on_device) {
            if (((uintptr_t)db & 15) != 0) { rdm_set_error("rdm_knn_create: device db pointer must be 16-byte aligned"); rc = RDM_ERR_ARG; break; }
            h->db = db;
        } else {
            if (cudaMalloc(&h->db_owned, bytes) != cudaSuccess) { rdm_set_error("rdm_knn_create: cudaMalloc(%zu) failed", bytes); rc = RDM_ERR_CUDA; break; }
            if (cudaMemcpy(h->db_owned, db, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { rdm_set_error("rdm_knn_create: H2D copy failed"); rc = RDM_ERR_CUDA; break; }
            h->db = h->db_owned;
        }
        h->max_grid = rdm_num_sms(device) * 8;
        const long long tc_tiles = (n + 127) / 128;
        int tc_stride = (int)(tc_tiles / 512); if (tc_stride > 16) tc_stride = 16; if (tc_stride < 1) tc_stride = 1;
        const long long tc_per_q = (dtype == RDM_DTYPE_F16 && d == 512) ? knn_tc_sample_rows(n, tc_stride) : 0;
        if (cudaMalloc(&h->inv, (size_t)n * sizeof(float)) != cudaSuccess ||
            cudaMalloc(&h->lists, (size_t)(MAX_TCQ / MAX_QP) * h->max_grid * MAX_QP * LIST * sizeof(u64)) != cudaSuccess ||
            cudaMalloc(&h->maxima, (h->maxima_keys = std::max((size_t)h->max_grid * SCAN_THREADS * MAX_QP, (size_t)tc_per_q * MAX_TCQ)) * sizeof(u64)) != cudaSuccess ||
            cudaMalloc(&h->cand, (size_t)MAX_TCQ * CAND_CAP * sizeof(u64)) != cudaSuccess ||
            cudaMalloc(&h->thr_key, (size_t)MAX_TCQ * sizeof(u64)) != cudaSuccess ||
            cudaMalloc(&h->cand_cnt, (size_t)(2 * MAX_TCQ) * sizeof(unsigned)) != cudaSuccess ||
            cudaMalloc(&h->thr_part, (size_t)MAX_TCQ * THR_P * LIST * sizeof(u64)) != cudaSuccess ||
            cudaMalloc(&h->thr_done, (size_t)MAX_TCQ * sizeof(unsigned)) != cudaSuccess ||
            cudaMemset(h->thr_done, 0, (size_t)MAX_TCQ * sizeof(unsigned)) != cudaSuccess ||
            cudaMalloc(&h->qsplit, (size_t)knn_tc_queries_bytes()) != cudaSuccess ||
            cudaMalloc(&h->qhat, (size_t)QHAT_ROWS * d * sizeof(float)) != cudaSuccess ||
            cudaMalloc(&h->fused_ws, knn_tc_fused_ws_bytes(device)) != cudaSuccess) {
            rdm_set_error("rdm_knn_create: workspace cudaMalloc failed"); rc = RDM_ERR_CUDA; break;
        }
        rc = do_create(h);
    } while (0);
    if (rc != RDM_OK) { rdm_knn_destroy(h); return rc; }
    *out = h;
    return RDM_OK;
}

void rdm_knn_destroy(rdm_knn_t* h) {
    if (!h) return;
    DeviceGuard guard(h->device);
    if (h->db_owned) cudaFree(h->db_owned);
    if (h->inv) cudaFree(h->inv);
    if (h->lists) cudaFree(h->lists);
    if (h->maxima) cudaFree(h->maxima);
    if (h->cand) cudaFree(h->cand);
    if (h->thr_key) cudaFree(h->thr_key);
    if (h->cand_cnt) cudaFree(h->cand_cnt);
    if (h->qsplit) cudaFree(h->qsplit);
    if (h->qhat) cudaFree(h->qhat);
    if (h->fused_ws) cudaFree(h->fused_ws);
    if (h->thr_part) cudaFree(h->thr_part);
    if (h->thr_done) cudaFree(h->thr_done);
    delete h;
}

int64_t rdm_knn_size(const rdm_knn_t* h) { return h ? h->n : 0; }
int rdm_knn_get_inv_norms(rdm_knn_t* h, float* out, void* stream) {
    RDM_REQUIRE(h && out, RDM_ERR_ARG, "rdm_knn_get_inv_norms: null argument");
    DeviceGuard guard(h->device);
    RDM_CHECK_CUDA(cudaMemcpyAsync(out, h->inv, (size_t)h->n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return RDM_OK;
}

int rdm_knn_search(rdm_knn_t* h, const float* q, int32_t nq, int32_t k, int64_t* idx_out, float* dist_out, double* score_out, void* stream) {
    RDM_REQUIRE(h && q && idx_out && dist_out, RDM_ERR_ARG, "rdm_knn_search: null argument");
    RDM_REQUIRE(k >= 1 && k <= RDM_KNN_MAX_K, RDM_ERR_ARG, "rdm_knn_search: k=%d outside 1..%d", k, RDM_KNN_MAX_K);
    RDM_REQUIRE(nq >= 0, RDM_ERR_ARG, "rdm_knn_search: nq=%d", nq);
    if (nq == 0) return RDM_OK;
    DeviceGuard guard(h->device);
    return do_search(h, q, nq, k, (long long*)idx_out, dist_out, score_out, (cudaStream_t)stream);
}

int rdm_knn_normalize(const float* q, int32_t nq, int32_t d, float* q_hat_out, int32_t device, void* stream) {
    RDM_REQUIRE(q && q_hat_out, RDM_ERR_ARG, "rdm_knn_normalize: null argument");
    RDM_REQUIRE(nq >= 0, RDM_ERR_ARG, "rdm_knn_normalize: nq=%d", nq);
    DeviceGuard guard(device);
    return normalize_rows(q, nq, d, q_hat_out, (cudaStream_t)stream);
}

int rdm_knn_search_raw(rdm_knn_t* h, const float* q_raw, int32_t nq, int32_t k, int64_t* idx_out, float* dist_out, double* score_out, void* stream) {
    RDM_REQUIRE(h && q_raw && idx_out && dist_out, RDM_ERR_ARG, "rdm_knn_search_raw: null argument");
    RDM_REQUIRE(k >= 1 && k <= RDM_KNN_MAX_K, RDM_ERR_ARG, "rdm_knn_search_raw: k=%d outside 1..%d", k, RDM_KNN_MAX_K);
    RDM_REQUIRE(nq >= 0, RDM_ERR_ARG, "rdm_knn_search_raw: nq=%d", nq);
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    static const bool no_tc = getenv("RDM_KNN_NO_TC") != nullptr;
    if (h->dtype == RDM_DTYPE_F16 && h->d == 512 && nq >= 3 && !no_tc) {
        // tensor-core passes of <= 64 queries: ONE launch normalises the pass, writes its fp16 hi / lo rows and resets the scan's barrier
        for (int q0 = 0; q0 < nq; q0 += MAX_TCQ) {
            const int cnt = nq - q0 < MAX_TCQ ? nq - q0 : MAX_TCQ;
            RDM_TRY(normalize_rows(q_raw + (size_t)q0 * 512, cnt, 512, h->qhat, st, (__half*)h->qsplit, knn_tc_pass_queries(cnt), knn_tc_fused_grid_bar(h->fused_ws, h->device)));
            RDM_TRY((search_pass_tc<__half, 512>(h, h->qhat, cnt, k, (long long*)idx_out + (size_t)q0 * k, dist_out + (size_t)q0 * k, score_out ? score_out + (size_t)q0 * k : nullptr, st, 1)));
        }
        return RDM_OK;
    }
    for (int q0 = 0; q0 < nq; q0 += QHAT_ROWS) {
        const int cnt = nq - q0 < QHAT_ROWS ? nq - q0 : QHAT_ROWS;
        RDM_TRY(normalize_rows(q_raw + (size_t)q0 * h->d, cnt, h->d, h->qhat, st));
        RDM_TRY(do_search(h, h->qhat, cnt, k, (long long*)idx_out + (size_t)q0 * k, dist_out + (size_t)q0 * k, score_out ? score_out + (size_t)q0 * k : nullptr, st));
    }
    return RDM_OK;
}

int rdm_knn_merge(const int64_t* idx_in, const double* score_in, int32_t parts, int32_t nq, int32_t k,
                  int64_t* idx_out, float* dist_out, double* score_out, int32_t device, void* stream) {
    RDM_REQUIRE(idx_in && score_in && idx_out && dist_out, RDM_ERR_ARG, "rdm_knn_merge: null argument");
    RDM_REQUIRE(parts >= 1 && k >= 1 && nq >= 0, RDM_ERR_ARG, "rdm_knn_merge: bad sizes");
    if (nq == 0) return RDM_OK;
    DeviceGuard guard(device);
    knn_merge_kernel<<<nq, 32, 0, (cudaStream_t)stream>>>((const long long*)idx_in, score_in, parts, nq, k, (long long*)idx_out, dist_out, score_out);
    RDM_COUNT_LAUNCH();
    RDM_CHECK_CUDA(cudaGetLastError());
    return RDM_OK;
}

int rdm_knn_gather(rdm_knn_t* h, const int64_t* idx, int64_t count, float* out, void* stream) {
    RDM_REQUIRE(h && idx && out, RDM_ERR_ARG, "rdm_knn_gather: null argument");
    RDM_REQUIRE(count >= 0, RDM_ERR_ARG, "rdm_knn_gather: count<0");
    DeviceGuard guard(h->device);
    return do_gather(h, (const long long*)idx, count, out, (cudaStream_t)stream);
}

}  // extern "C"
