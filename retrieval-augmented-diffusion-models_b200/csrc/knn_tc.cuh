// Tensor-core scans for fp16 / d=512 databases (knn_tc.cu); same survivor-buffer contract as knn_scan_kernel<MAIN>.
#pragma once
#include <cuda_runtime.h>
int knn_tc_queries_bytes();
int knn_tc_pass_queries(int nq);                                  // 16, 32, 64 or 128 query columns for a pass of nq <= 128 queries
long long knn_tc_sample_rows(long long n, int tile_stride);       // keys per query written by the sample pass
// sample != 0: every tile_stride-th 128-row tile, writes maxima[q][per_q] (and prepares qsplit_ws from q).
// sample == 0: the main scan; survivors of thr_key go to cand / cand_cnt.  q: fp32 [nq_valid, 512] normalised queries (device).
int knn_scan_tc(const void* db_f16, const float* inv, long long n, int device, const float* q, int nq_valid, void* qsplit_ws, int sample, int tile_stride,
                unsigned long long* maxima, long long per_q, const unsigned long long* thr_key, unsigned long long* cand, unsigned* cand_cnt, cudaStream_t st);
// ONE launch per search (plus the query split): sample phase, grid barrier, thresholds and main scan in a persistent cooperative kernel.
// Resets cand_cnt / *overflow itself.  fused_ws: knn_tc_fused_ws_bytes(device) bytes of device memory owned by the searcher.
// Returns RDM_OK, a negative error code, or 1 if this path cannot run here (small database / no cooperative launch): use knn_scan_tc then.
size_t knn_tc_fused_ws_bytes(int device);
// presplit != 0: qsplit_ws already holds the fp16 hi / lo rows of the queries and the barrier counter was reset (fused normalisation kernel).
int knn_scan_tc_fused(const void* db_f16, const float* inv, long long n, int device, const float* q, int nq_valid, int k, void* qsplit_ws,
                      unsigned long long* cand, unsigned* cand_cnt, unsigned* overflow, void* fused_ws, int presplit, float* slack_used, cudaStream_t st);
// *slack_used: the score slack the pass decided with (hi + lo query rows: 3e-5; hi rows only, more than 16 queries: 1.05e-3) -- the select
// kernel's cut must use the same value.  Up to 128 queries per fused pass; knn_scan_tc (the three-kernel path) takes at most 64.
unsigned* knn_tc_fused_grid_bar(void* fused_ws, int device);
// fp32 databases (d = 512): the fused scan on kind::tf32 MMAs (rows and queries read as tf32, slack 4.2e-3), at most 64 queries per pass;
// qpad_ws: knn_tc_queries_bytes() bytes.  Returns like knn_scan_tc_fused (1: not usable here -> CUDA-core scan).
int knn_scan_tc_fused_f32(const void* db_f32, const float* inv, long long n, int device, const float* q, int nq_valid, int k, void* qpad_ws,
                          unsigned long long* cand, unsigned* cand_cnt, unsigned* overflow, void* fused_ws, float* slack_used, cudaStream_t st);
