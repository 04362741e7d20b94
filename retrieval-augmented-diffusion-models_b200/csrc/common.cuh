// Shared helpers for librdm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>
#include <utility>

#define RDM_OK 0
#define RDM_ERR_ARG -1
#define RDM_ERR_CUDA -2
#define RDM_ERR_STATE -3
#define RDM_ERR_UNSUPPORTED -4

// thread-local last-error message (returned by rdm_last_error)
void rdm_set_error(const char* fmt, ...);

#define RDM_CHECK_CUDA(expr)                                                                         \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            rdm_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));     \
            return RDM_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

#define RDM_REQUIRE(cond, code, ...)     \
    do {                                 \
        if (!(cond)) {                   \
            rdm_set_error(__VA_ARGS__);  \
            return (code);               \
        }                                \
    } while (0)

#define RDM_TRY(expr)                \
    do {                             \
        int _rc = (expr);            \
        if (_rc != RDM_OK) return _rc; \
    } while (0)

// launch counter (bench.py reports gpu_launches from it)
extern unsigned long long g_rdm_launches;
#define RDM_COUNT_LAUNCH() (++g_rdm_launches)

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Launch with the programmatic-stream-serialization attribute (PDL).  A kernel launched this way MUST execute griddepcontrol.wait
// (pdl_wait) in every CTA before touching data of earlier kernels; RDM_PDL=0 disables the attribute globally.
extern int g_rdm_use_pdl, g_rdm_use_pdl_glue;       // glue kernels: measured slightly slower with PDL -> opt-in (RDM_PDL_GLUE=1)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = (g_rdm_use_pdl && g_rdm_use_pdl_glue) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int rdm_num_sms(int device);
