// Tensor-core main scan for fp16 / d=512 databases (knn_tc.cu); same survivor-buffer contract as knn_scan_kernel<MAIN>.
#pragma once
#include <cuda_runtime.h>
int knn_tc_queries_bytes();
// q: fp32 [nq_valid, 512] normalised queries (device); qsplit_ws: knn_tc_queries_bytes() of device scratch.
int knn_scan_tc(const void* db_f16, const float* inv, long long n, int device, const float* q, int nq_valid, void* qsplit_ws,
                const unsigned long long* thr_key, unsigned long long* cand, unsigned* cand_cnt, cudaStream_t st);
