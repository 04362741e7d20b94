// The tensor-core engine (gemm_tc.cu, attention_mma.cu: tcgen05 / TMA / mma.sync PTX) cannot be emulated on the host.  These stand-ins make
// the SIMT translation units link; the strict fp32 mode of the executors never calls them, every other mode fails loudly.
#include "gemm_tc.cuh"
bool gemm_tc_supported(const TcA&) { return true; }     // shape support is a property of the engine, not of the emulation: the call itself reports the error
int gemm_tc(const TcA&, const TcW&, const GemmEpi&, __nv_bfloat16*, __nv_bfloat16*, int, int, int, cudaStream_t) {
    rdm_set_error("emulation: the tcgen05 GEMM engine is not available on the host (use RDM_UNET_MODE_FP32)");
    return RDM_ERR_UNSUPPORTED;
}
void gemm_tc_set_workspace_slot(int) {}
bool k_attention_mma_supported(int, int) { return false; }
int k_attention_mma(const __half*, const __half*, const __half*, int, int, int, int, int, float, Out4, cudaStream_t) {
    rdm_set_error("emulation: the warp-MMA attention kernel is not available on the host");
    return RDM_ERR_UNSUPPORTED;
}
// one-launch GroupNorm of the tensor-core modes (gn_fused.cu: thread-block clusters + distributed shared memory); the strict mode never asks for it
bool k_gn_fused_supported(int, int, int, bool) { return false; }
int k_gn_fused(View, int, int, int, const double*, int, float, const float*, const float*, int, Out4, Out4, cudaStream_t) {
    rdm_set_error("emulation: the cluster GroupNorm kernel is not available on the host (use RDM_UNET_MODE_FP32)");
    return RDM_ERR_UNSUPPORTED;
}

// tensor-core kNN scans (knn_tc.cu): the searcher is driven with RDM_KNN_NO_TC=1 under emulation
#include "knn_tc.cuh"
int knn_tc_queries_bytes() { return 256; }
int knn_tc_pass_queries(int) { return 16; }
long long knn_tc_sample_rows(long long, int) { return 0; }
int knn_scan_tc(const void*, const float*, long long, int, const float*, int, void*, int, int, unsigned long long*, long long, const unsigned long long*, unsigned long long*,
                unsigned*, cudaStream_t) {
    rdm_set_error("emulation: the tcgen05 kNN scan is not available on the host (set RDM_KNN_NO_TC=1)");
    return RDM_ERR_UNSUPPORTED;
}
size_t knn_tc_fused_ws_bytes(int) { return 256; }
unsigned* knn_tc_fused_grid_bar(void*, int) { return nullptr; }
int knn_scan_tc_fused(const void*, const float*, long long, int, const float*, int, int, void*, unsigned long long*, unsigned*, unsigned*, void*, int, float*, cudaStream_t) {
    rdm_set_error("emulation: the tcgen05 kNN scan is not available on the host (set RDM_KNN_NO_TC=1)");
    return RDM_ERR_UNSUPPORTED;
}
int knn_scan_tc_fused_f32(const void*, const float*, long long, int, const float*, int, int, void*, unsigned long long*, unsigned*, unsigned*, void*, float*, cudaStream_t) {
    return 1;       // not usable under emulation: the caller takes the CUDA-core scan
}

// ---- test-only entry points (tests/test_engine_identities_host.py): weight-load / first-conv kernels of kernels.cu that no strict-mode
// forward reaches through the C ABI in isolation.  Pointers are plain host memory (the emulation runs kernels on the host).
extern "C" int emu_fold_up_weights(const float* w, int N, int C, float* out) { return k_fold_up_weights(w, N, C, out, nullptr); }
extern "C" int emu_add_vec(const float* a, const float* b, float* out, int n) { return k_add_vec(a, b, out, n, nullptr); }
extern "C" int emu_conv_first(const float* x, int B, int H, int W, int Cin, const float* w, const float* bias, int Cout, float* out) {
    if (!k_conv_first_supported(Cin, Cout)) return RDM_ERR_UNSUPPORTED;
    return k_conv_first(View(const_cast<float*>(x), Cin, Cin), B, H, W, w, bias, Cout, View(out, Cout, Cout), nullptr);
}
