/* Oracle: exact cosine top-k over a row-major CLIP database.  TEST INFRASTRUCTURE ONLY
 * (see oracle/__init__.py) -- never linked into the product library.
 *
 * Restates what the reference asks of its searcher:
 *   rdm/data/retrieval_dataset/dsetbuilder.py:574   DB rows are L2-normalised when the index is built
 *   rdm/data/retrieval_dataset/dsetbuilder.py:487-490, rdm/models/diffusion/ddpm.py:906-908
 *                                                   queries are L2-normalised (numpy fp32) by the caller,
 *                                                   searcher.search_batched(q_hat, final_num_neighbors=k)
 *                                                   returns (indices, dot-product scores) best first.
 * The reference's backend is ScaNN 1.2.4 (environment.yaml:33; approximate, not vendored, not installable
 * here), so this file defines the EXACT result ScaNN approximates -- PARITY UNPINNED by the reference:
 *
 *   inv_i      = (float)( 1.0 / sqrt( sum_j (double)d_ij * (double)d_ij ) )          j ascending
 *   score(q,i) = ( sum_j (double)q_j * (double)d_ij ) * (double)inv_i                j ascending
 *   result     = the k rows with the largest score, ties broken by the LOWEST row index,
 *                returned best first; the reported distance is (float)score.
 *
 * All sums are sequential IEEE double additions of separately rounded products (build with
 * -ffp-contract=off; the CUDA re-rank uses __dmul_rn/__dadd_rn in the same order), so scores are
 * bit-reproducible and independent of batching, sharding and thread count.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else { int e = -1; do { man <<= 1; e++; } while (!(man & 0x400u)); bits = sign | ((uint32_t)(112 - e) << 23) | ((man & 0x3ffu) << 13); }
    } else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
    else bits = sign | ((exp + 112) << 23) | (man << 13);
    float f; memcpy(&f, &bits, 4); return f;
}

static inline double elem(const void* db, int dtype, int64_t off) {
    return dtype == 0 ? (double)((const float*)db)[off] : (double)half_to_float(((const uint16_t*)db)[off]);
}

/* dtype: 0 = float32, 1 = float16 */
void knn_ref_inv_norms(const void* db, int64_t n, int d, int dtype, float* inv) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        double s = 0.0;
        for (int j = 0; j < d; j++) { double v = elem(db, dtype, i * d + j); s = s + v * v; }
        inv[i] = (float)(1.0 / sqrt(s));
    }
}

static inline int better(double sa, int64_t ia, double sb, int64_t ib) { return sa > sb || (sa == sb && ia < ib); }

/* q: [nq, d] float32, already normalised by the caller.  idx_base is added to reported indices (shards). */
void knn_ref_search(const void* db, const float* inv, int64_t n, int d, int dtype, const float* q, int nq, int k,
                    int64_t idx_base, int64_t* idx_out, float* dist_out, double* score_out) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int qi = 0; qi < nq; qi++) {
        double* bs = (double*)malloc(sizeof(double) * k);
        int64_t* bi = (int64_t*)malloc(sizeof(int64_t) * k);
        int cnt = 0;
        const float* qq = q + (int64_t)qi * d;
        for (int64_t i = 0; i < n; i++) {
            double s = 0.0;
            if (dtype == 0) { const float* r = (const float*)db + i * d; for (int j = 0; j < d; j++) s = s + (double)qq[j] * (double)r[j]; }
            else { const uint16_t* r = (const uint16_t*)db + i * d; for (int j = 0; j < d; j++) s = s + (double)qq[j] * (double)half_to_float(r[j]); }
            s = s * (double)inv[i];
            if (cnt == k && !better(s, i, bs[k - 1], bi[k - 1])) continue;
            int p = cnt < k ? cnt++ : k - 1;                       /* insertion into the sorted best-first list */
            while (p > 0 && better(s, i, bs[p - 1], bi[p - 1])) { bs[p] = bs[p - 1]; bi[p] = bi[p - 1]; p--; }
            bs[p] = s; bi[p] = i;
        }
        for (int j = 0; j < k; j++) {
            int ok = j < cnt;
            idx_out[(int64_t)qi * k + j] = ok ? bi[j] + idx_base : -1;
            dist_out[(int64_t)qi * k + j] = ok ? (float)bs[j] : -INFINITY;
            if (score_out) score_out[(int64_t)qi * k + j] = ok ? bs[j] : -INFINITY;
        }
        free(bs); free(bi);
    }
}
