/* librdm_b200 -- C ABI of the B200-native retrieval-augmented diffusion sampling hot path.
 *
 * The reference (CompVis/retrieval-augmented-diffusion-models @ 1017f2b) has no FFI of its own: its
 * boundary is YAML `target:` dotted paths resolved to Python classes.  Every
 * entry point below therefore cites the Python call site it replaces; the Python classes at the same
 * import paths (retrieval-augmented-diffusion-models_b200/rdm/...) bind these symbols via ctypes.
 *
 * Conventions: plain pointers and sizes only; all `*_dev` pointers are device pointers on the handle's
 * device, contiguous, 16-byte aligned, owned by the caller; work is stream-ordered on `stream`
 * (a cudaStream_t passed as void*; NULL = legacy default stream) with no hidden synchronisation unless
 * stated.  Every function returns 0 on success or a negative code and records a message retrievable
 * with rdm_last_error() (thread-local).  Handles are per-device and not thread-safe.
 */
#ifndef RDM_B200_H
#define RDM_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RDM_API __attribute__((visibility("default")))
#else
#define RDM_API
#endif

#define RDM_DTYPE_F32 0
#define RDM_DTYPE_F16 1

RDM_API const char* rdm_last_error(void);
/* ABI version of this header (bumped on any signature change). */
RDM_API int rdm_abi_version(void);
/* Number of kernel launches issued by this library so far in this process (bench.py `gpu_launches`). */
RDM_API unsigned long long rdm_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Retrieval: exact cosine top-k over the in-HBM CLIP database.
 * Replaces the ScaNN searcher the reference builds in
 *   rdm/data/retrieval_dataset/dsetbuilder.py:534-619 (train_searcher; DB rows L2-normalised at :574)
 * and queries at
 *   rdm/data/retrieval_dataset/dsetbuilder.py:490, rdm/models/diffusion/ddpm.py:298,906-908,
 *   rdm/models/autoregression/transformer.py:327-329, rdm/data/base.py:81-83
 *   (`searcher.search_batched(q_hat, final_num_neighbors=k) -> (indices, distances)`).
 * Result definition (bit-exact, batching/sharding independent): oracle/knn_ref.c.
 * ------------------------------------------------------------------------------------------------ */
typedef struct rdm_knn rdm_knn_t;

/* Create a searcher over `n` rows of dimension `d` (d in {256,512,768,1024}).  `db` is the RAW
 * (un-normalised) embedding matrix, row-major, dtype RDM_DTYPE_F16 or RDM_DTYPE_F32.
 * db_on_device != 0: `db` is a device pointer that the caller keeps alive for the handle's lifetime
 *                    (no copy; 20.9M x 512 fp16 = 21.4 GB stays resident once).
 * db_on_device == 0: `db` is a host pointer; the rows are copied to HBM in pinned-staged chunks.
 * `idx_base` is added to every reported index (row-sharded databases: rank r passes its first row).
 * Computes the per-row inverse L2 norms on the device (one pass over the DB) and synchronises. */
RDM_API int rdm_knn_create(rdm_knn_t** out, const void* db, int64_t n, int32_t d, int32_t dtype,
                   int32_t db_on_device, int64_t idx_base, int32_t device);
RDM_API void rdm_knn_destroy(rdm_knn_t* h);
RDM_API int64_t rdm_knn_size(const rdm_knn_t* h);
/* Copies the float32 inverse L2 norms [n] computed at create time into out_dev (device). */
RDM_API int rdm_knn_get_inv_norms(rdm_knn_t* h, float* out_dev, void* stream);

/* search_batched: q_hat_dev float32 [nq, d], ALREADY L2-normalised by the caller exactly as the
 * reference does (numpy fp32, ddpm.py:907).  1 <= k <= RDM_KNN_MAX_K.
 * Outputs (device): idx_out int64 [nq,k] (global indices, best first), dist_out float32 [nq,k]
 * (= (float)score), score_out float64 [nq,k] or NULL (exact scores, used by the shard merge). */
#define RDM_KNN_MAX_K 24
RDM_API int rdm_knn_search(rdm_knn_t* h, const float* q_hat_dev, int32_t nq, int32_t k,
                   int64_t* idx_out_dev, float* dist_out_dev, double* score_out_dev, void* stream);

/* Merge `parts` per-shard results (e.g. the all_gather of every rank's rdm_knn_search output):
 * idx_in int64 [parts, nq, k], score_in float64 [parts, nq, k] -> global top-k by (score desc, idx asc). */
RDM_API int rdm_knn_merge(const int64_t* idx_in_dev, const double* score_in_dev, int32_t parts, int32_t nq, int32_t k,
                  int64_t* idx_out_dev, float* dist_out_dev, double* score_out_dev, int32_t device, void* stream);

/* `data_pool['embedding'][nns]` then `.to(device).to(float)` (ddpm.py:921, dsetbuilder.py:493):
 * gathers RAW rows as float32.  idx_dev holds GLOBAL indices; rows outside this shard
 * ([idx_base, idx_base+n)) are written as zeros so shards can be summed. out_dev float32 [count, d]. */
RDM_API int rdm_knn_gather(rdm_knn_t* h, const int64_t* idx_dev, int64_t count, float* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif
