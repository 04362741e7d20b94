// HOST EMULATION of the slice of CUDA that csrc/rarm.cu uses -- TEST INFRASTRUCTURE (tests/test_rarm_emulated.py), never shipped.
//
// Purpose: run the UNMODIFIED kernel and host source of a translation unit on the CPU so that indexing, barrier placement, shuffle
// networks, shared-memory reuse, graph-replay argument baking and the C-ABI host logic are checked in a container without a GPU.
// tests/emu/build_emu.py rewrites only the launch syntax (`k<<<g, b, s, st>>>(args)` -> `emu::launch(k, g, b, s, st, args)`) and the
// `extern __shared__ T name[];` declarations; everything else compiles as written against this header.
//
// Execution model: blocks are distributed over a few OS threads; the threads of ONE block are cooperative fibers (ucontext) on one OS
// thread.  `__syncthreads()` and the `__shfl_*_sync` exchanges are yield points: a fiber parks until all live fibers of its block / all 32
// lanes of its warp have arrived.  `__shared__` becomes `static thread_local` (one block per OS thread at a time), dynamic shared memory
// is a per-block buffer, device memory is host memory.  A stream capture records launches (with their by-value arguments, exactly what a
// CUDA graph bakes) and `cudaGraphLaunch` replays them.  What this does NOT check: alignment faults, shared-memory / register limits,
// launch-configuration limits, memory-model races between warps of different blocks, performance.
#pragma once
#include <ucontext.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static thread_local


struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

// ---- runtime API subset ---------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmu = 1 };
typedef struct EmuStream* cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal, cudaStreamCaptureModeThreadLocal, cudaStreamCaptureModeRelaxed };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
struct EmuGraph { std::vector<std::function<void()>> nodes; };
typedef EmuGraph* cudaGraph_t;
typedef EmuGraph* cudaGraphExec_t;
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 1, cudaLaunchAttributeClusterDimension = 2 };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; struct { int programmaticStreamSerializationAllowed; struct { unsigned x, y, z; } clusterDim; } val; };
struct cudaLaunchConfig_t { dim3 gridDim, blockDim; size_t dynamicSmemBytes; cudaStream_t stream; cudaLaunchAttribute* attrs; unsigned numAttrs; };

cudaError_t cudaMalloc(void** p, size_t bytes);
template <typename T> cudaError_t cudaMalloc(T** p, size_t bytes) { return cudaMalloc((void**)p, bytes); }      // the runtime's C++ overload
cudaError_t cudaFree(void* p);
cudaError_t cudaMemset(void* p, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t st = nullptr);
typedef struct EmuEvent* cudaEvent_t;
enum { cudaEventRecordDefault = 0, cudaEventRecordExternal = 1 };
cudaError_t cudaEventCreate(cudaEvent_t* e);
enum { cudaEventDefault = 0, cudaEventDisableTiming = 2 };
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags);
// launches execute synchronously in issue order under emulation (and a capture records them in that order), which satisfies every
// cross-stream dependency an event wait can express
cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st = nullptr);
cudaError_t cudaEventRecordWithFlags(cudaEvent_t e, cudaStream_t st, unsigned flags);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaMemcpy(void* d, const void* s, size_t bytes, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t bytes, cudaMemcpyKind k, cudaStream_t st = nullptr);
cudaError_t cudaMemcpy2D(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind k);
cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind k, cudaStream_t st = nullptr);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* st, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t st);
cudaError_t cudaStreamSynchronize(cudaStream_t st);
cudaError_t cudaStreamBeginCapture(cudaStream_t st, cudaStreamCaptureMode mode);
cudaError_t cudaStreamEndCapture(cudaStream_t st, cudaGraph_t* g);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long flags);
cudaError_t cudaGraphDestroy(cudaGraph_t g);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e);
cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t st);
cudaError_t cudaGetLastError();
const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetDevice(int* d);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int dev);
template <typename F> cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ---- device-side intrinsics -------------------------------------------------------------------------------------------
namespace emu {
void sync_block();
uint64_t warp_exchange(uint64_t mine, int src_lane);       // returns the value posted by lane `src_lane` of the caller's warp (all 32 lanes call)
const uint64_t* warp_all(uint64_t mine);                   // posts `mine`, waits for the warp, returns the 32 posted values (valid until the warp's next exchange)
void yield();                                              // lets the other threads of the block run (spin loops on memory another thread will write)
int lane_id();
extern thread_local void* dyn_smem;
void run_launch(dim3 grid, dim3 block, size_t smem, cudaStream_t st, std::function<void()> thread_body);
template <typename... KArgs, typename... Args>
void launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    auto body = [=]() { kern(static_cast<KArgs>(args)...); };          // arguments captured BY VALUE at launch time
    run_launch(grid, block, smem, st, body);
}
}  // namespace emu
template <typename... KArgs, typename... Args> cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kern)(KArgs...), Args&&... args) {
    emu::launch(kern, cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, cfg->stream, static_cast<KArgs>(args)...);       // launch attributes (PDL) have no functional effect
    return cudaSuccess;
}
namespace emu {
template <typename T> inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, "shuffle of > 8 bytes"); memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
}  // namespace emu

inline void __syncthreads() { emu::sync_block(); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return emu::from_bits<T>(emu::warp_exchange(emu::to_bits(v), emu::lane_id() ^ lane_mask)); }
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return emu::from_bits<T>(emu::warp_exchange(emu::to_bits(v), src & 31)); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d) { int l = emu::lane_id(); return emu::from_bits<T>(emu::warp_exchange(emu::to_bits(v), l >= (int)d ? l - (int)d : l)); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned d) { int l = emu::lane_id(); return emu::from_bits<T>(emu::warp_exchange(emu::to_bits(v), l + (int)d < 32 ? l + (int)d : l)); }

template <typename T> inline T __ldg(const T* p) { return *p; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float __fsqrt_rn(float a) { volatile float r = sqrtf(a); return r; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
// atomics: shared-memory targets are only touched by the fibers of one block (never concurrent); GLOBAL targets can be hit by blocks on
// other OS threads, so the read-modify-write is a real atomic either way
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
template <typename F, typename U> inline F emu_atomic_add_fp(F* p, F v) {
    U* q = reinterpret_cast<U*>(p); U old = __atomic_load_n(q, __ATOMIC_RELAXED), neu; F f;
    do { memcpy(&f, &old, sizeof(F)); F r = f + v; memcpy(&neu, &r, sizeof(F)); } while (!__atomic_compare_exchange_n(q, &old, neu, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return f;
}
inline float atomicAdd(float* p, float v) { return emu_atomic_add_fp<float, uint32_t>(p, v); }
inline double atomicAdd(double* p, double v) { return emu_atomic_add_fp<double, uint64_t>(p, v); }
inline float __frcp_rn(float x) { volatile float r = 1.0f / x; return r; }
inline float __expf(float x) { return expf(x); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_exchange(0, 0); }      // a real rendezvous: lanes run one after the other between yield points
inline unsigned __ballot_sync(unsigned, int pred) { const uint64_t* v = emu::warp_all(pred ? 1 : 0); unsigned m = 0; for (int l = 0; l < 32; l++) m |= (unsigned)(v[l] & 1u) << l; return m; }
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
// a failed compare-and-swap is a yield point, so spin locks make progress under cooperative scheduling
template <typename T> inline T emu_atomic_cas(T* p, T cmp, T val) {
    T expected = cmp;
    if (__atomic_compare_exchange_n(p, &expected, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) return cmp;
    emu::yield();
    return expected;
}
inline int atomicCAS(int* p, int c, int v) { return emu_atomic_cas(p, c, v); }
inline unsigned atomicCAS(unsigned* p, unsigned c, unsigned v) { return emu_atomic_cas(p, c, v); }
inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long c, unsigned long long v) { return emu_atomic_cas(p, c, v); }
inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <typename T> inline T __ldcs(const T* p) { return *p; }
template <typename T> inline T __ldcg(const T* p) { return *(const volatile T*)p; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline long long clock64() { return 0; }
inline void __trap() { fprintf(stderr, "emu: __trap()\n"); abort(); }
template <typename F> cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return cudaSuccess; }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
#define __align__(n) __attribute__((aligned(n)))
inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
