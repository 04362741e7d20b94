// Scheduler and runtime of the host emulation declared in include/cuda_runtime.h (test infrastructure).
#include "cuda_runtime.h"
#include <mutex>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim(1), gridDim(1);

namespace emu {

void check_device_guards(const char* when);
thread_local void* dyn_smem = nullptr;

namespace {
constexpr size_t STACK_BYTES = 64 * 1024;
enum State { READY, WAIT_BLOCK, WAIT_WARP, DONE };
struct Fiber {
    ucontext_t ctx;
    State state = READY;
    uint3 tid{0, 0, 0};
    int linear = 0;
};
struct Warp { uint64_t slot[2][32]; int parity = 0; int arrived = 0; int live = 0; };
struct BlockRun {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    ucontext_t main;
    int cur = -1, live = 0, block_arrived = 0;
    const std::function<void()>* body = nullptr;
};
thread_local BlockRun* t_run = nullptr;
struct StackPool { std::vector<char*> v; ~StackPool() { for (char* p : v) free(p); } };      // freed when the worker thread of a launch exits
thread_local StackPool t_stacks;

void fiber_entry() {
    BlockRun* r = t_run;
    (*r->body)();
    Fiber& f = r->fibers[r->cur];
    f.state = DONE;
    r->live--;
    r->warps[f.linear / 32].live--;
    // a thread that exits releases barriers the remaining threads are waiting on (CUDA: exited threads do not participate)
    if (r->live > 0 && r->block_arrived == r->live) { for (auto& g : r->fibers) if (g.state == WAIT_BLOCK) g.state = READY; r->block_arrived = 0; }
    Warp& w = r->warps[f.linear / 32];
    if (w.live > 0 && w.arrived == w.live) { for (int l = 0; l < 32; l++) { int i = (f.linear / 32) * 32 + l; if (i < (int)r->fibers.size() && r->fibers[i].state == WAIT_WARP) r->fibers[i].state = READY; } w.arrived = 0; w.parity ^= 1; }
    swapcontext(&f.ctx, &r->main);
}
void yield_to_scheduler() {
    BlockRun* r = t_run;
    swapcontext(&r->fibers[r->cur].ctx, &r->main);
}

void run_block(const std::function<void()>& body, dim3 grid, dim3 block, uint3 bidx, size_t smem) {
    const int n = (int)(block.x * block.y * block.z);
    while ((int)t_stacks.v.size() < n) t_stacks.v.push_back((char*)aligned_alloc(64, STACK_BYTES));
    std::vector<char> shared(smem + 64 + 64, (char)0x5A);                      // 64-byte guard after the dynamic shared memory of the block
    dyn_smem = (void*)(((uintptr_t)shared.data() + 63) & ~(uintptr_t)63);
    BlockRun run;
    run.fibers.resize(n); run.warps.resize((n + 31) / 32); run.live = n; run.body = &body;
    for (int i = 0; i < n; i++) {
        Fiber& f = run.fibers[i];
        f.linear = i; f.tid = uint3{(unsigned)(i % block.x), (unsigned)((i / block.x) % block.y), (unsigned)(i / (block.x * block.y))};
        run.warps[i / 32].live++;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = t_stacks.v[i]; f.ctx.uc_stack.ss_size = STACK_BYTES; f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, fiber_entry, 0);
    }
    t_run = &run;
    blockIdx = bidx; blockDim = block; gridDim = grid;
    // Order in which runnable threads are resumed within a scheduling round (EMU_SCHEDULE = forward | reverse | random:<seed>): results must
    // not depend on it -- a missing barrier (read-after-write or write-after-read across threads) shows up as a result that changes with the order.
    const char* sched = getenv("EMU_SCHEDULE");
    const int mode = !sched || !strncmp(sched, "forward", 7) ? 0 : !strncmp(sched, "reverse", 7) ? 1 : 2;
    uint64_t rng = mode == 2 ? (uint64_t)strtoull(strchr(sched, ':') ? strchr(sched, ':') + 1 : "1", nullptr, 10) * 0x9E3779B97F4A7C15ull + bidx.x * 1315423911u + bidx.y * 2654435761u + 1 : 0;
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = mode == 1 ? n - 1 - i : i;
    int idle_rounds = 0;
    while (run.live > 0) {
        bool progressed = false;
        if (mode == 2) for (int i = n - 1; i > 0; i--) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; std::swap(order[i], order[(int)(rng % (uint64_t)(i + 1))]); }
        for (int oi = 0; oi < n && run.live > 0; oi++) {
            const int i = order[oi];
            Fiber& f = run.fibers[i];
            if (f.state != READY) continue;
            run.cur = i; threadIdx = f.tid;
            swapcontext(&run.main, &f.ctx);
            progressed = true;
        }
        if (!progressed && ++idle_rounds > 2) { fprintf(stderr, "emu: deadlock in block (%u,%u,%u): %d live threads, none runnable (divergent barrier?)\n", bidx.x, bidx.y, bidx.z, run.live); abort(); }
        if (progressed) idle_rounds = 0;
    }
    for (int i = 0; i < 64; i++)
        if (((const char*)dyn_smem)[smem + i] != (char)0x5A) { fprintf(stderr, "emu: write past the %zu bytes of dynamic shared memory in block (%u,%u,%u)\n", smem, bidx.x, bidx.y, bidx.z); abort(); }
    t_run = nullptr; dyn_smem = nullptr;
}

struct Capture { bool on = false; EmuGraph* graph = nullptr; };
thread_local Capture t_capture;
std::atomic<int> g_workers{0};
}  // namespace

void sync_block() {
    BlockRun* r = t_run;
    Fiber& f = r->fibers[r->cur];
    if (++r->block_arrived == r->live) {
        for (auto& g : r->fibers) if (g.state == WAIT_BLOCK) g.state = READY;
        r->block_arrived = 0;
        return;                                       // the last arriver continues without parking
    }
    f.state = WAIT_BLOCK;
    yield_to_scheduler();
}
int lane_id() { return t_run->fibers[t_run->cur].linear & 31; }
uint64_t warp_exchange(uint64_t mine, int src_lane) {
    BlockRun* r = t_run;
    Fiber& f = r->fibers[r->cur];
    const int wi = f.linear / 32, lane = f.linear & 31;
    Warp& w = r->warps[wi];
    const int par = w.parity;
    w.slot[par][lane] = mine;
    if (++w.arrived == w.live) {
        for (int l = 0; l < 32; l++) { int i = wi * 32 + l; if (i < (int)r->fibers.size() && r->fibers[i].state == WAIT_WARP) r->fibers[i].state = READY; }
        w.arrived = 0; w.parity ^= 1;
    } else {
        f.state = WAIT_WARP;
        yield_to_scheduler();
    }
    return w.slot[par][src_lane];                     // double-buffered: the next exchange of this warp writes the other parity
}

const uint64_t* warp_all(uint64_t mine) {
    BlockRun* r = t_run;
    Fiber& f = r->fibers[r->cur];
    const int wi = f.linear / 32, lane = f.linear & 31;
    Warp& w = r->warps[wi];
    const int par = w.parity;
    w.slot[par][lane] = mine;
    if (++w.arrived == w.live) {
        for (int l = 0; l < 32; l++) { int i = wi * 32 + l; if (i < (int)r->fibers.size() && r->fibers[i].state == WAIT_WARP) r->fibers[i].state = READY; }
        w.arrived = 0; w.parity ^= 1;
    } else {
        f.state = WAIT_WARP;
        yield_to_scheduler();
    }
    return w.slot[par];
}
void yield() { yield_to_scheduler(); }                 // stays READY: resumed on the scheduler's next round

void run_launch(dim3 grid, dim3 block, size_t smem, cudaStream_t, std::function<void()> body) {
    if (t_capture.on) {                               // stream capture: record, do not execute
        t_capture.graph->nodes.push_back([=]() { Capture saved = t_capture; t_capture.on = false; run_launch(grid, block, smem, nullptr, body); t_capture = saved; });
        return;
    }
    const long nblocks = (long)grid.x * grid.y * grid.z;
    static const int maxw = [] { const char* e = getenv("EMU_WORKERS"); int n = e ? atoi(e) : (int)std::thread::hardware_concurrency(); return n < 1 ? 1 : (n > 16 ? 16 : n); }();
    const int nw = (int)std::min<long>(maxw, nblocks);
    std::atomic<long> next{0};
    auto worker = [&]() {
        for (;;) {
            long b = next.fetch_add(1);
            if (b >= nblocks) break;
            uint3 bidx{(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((long)grid.x * grid.y))};
            run_block(body, grid, block, bidx, smem);
        }
    };
    if (nw <= 1) worker();
    else {
        std::vector<std::thread> ts;
        for (int i = 0; i < nw; i++) ts.emplace_back(worker);
        for (auto& t : ts) t.join();
    }
    check_device_guards("after a kernel launch");
}

}  // namespace emu

// ---- runtime API -------------------------------------------------------------------------------------------------------
struct EmuStream { int id; };
// "Device" allocations carry a 256-byte guard zone on both sides; every launch ends with a check that no kernel wrote into one
// (an out-of-bounds WRITE aborts with the allocation size; out-of-bounds reads are not detected).
namespace {
constexpr size_t GUARD = 256;
constexpr unsigned char GUARD_BYTE = 0xA5;
struct Alloc { char* raw; size_t bytes; };
std::mutex g_alloc_mu;
std::vector<Alloc> g_allocs;
void check_guards(const char* when) {
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    for (const Alloc& a : g_allocs) {
        for (size_t i = 0; i < GUARD; i++) {
            if ((unsigned char)a.raw[i] != GUARD_BYTE || (unsigned char)a.raw[GUARD + a.bytes + i] != GUARD_BYTE) {
                fprintf(stderr, "emu: out-of-bounds write %s a %zu-byte device allocation (detected %s)\n", (unsigned char)a.raw[i] != GUARD_BYTE ? "before" : "after", a.bytes, when);
                abort();
            }
        }
    }
}
}  // namespace
namespace emu { void check_device_guards(const char* when) { check_guards(when); } }
cudaError_t cudaMalloc(void** p, size_t bytes) {
    char* raw = (char*)aligned_alloc(256, ((bytes + 255) & ~(size_t)255) + 2 * GUARD);
    if (!raw) return cudaErrorEmu;
    memset(raw, GUARD_BYTE, GUARD); memset(raw + GUARD, 0xCD, bytes); memset(raw + GUARD + bytes, GUARD_BYTE, GUARD);      // payload poisoned: fresh device memory is arbitrary
    { std::lock_guard<std::mutex> lk(g_alloc_mu); g_allocs.push_back({raw, bytes}); }
    *p = raw + GUARD;
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
    if (!p) return cudaSuccess;
    check_guards("at cudaFree");
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    for (size_t i = 0; i < g_allocs.size(); i++) if (g_allocs[i].raw + GUARD == (char*)p) { free(g_allocs[i].raw); g_allocs.erase(g_allocs.begin() + i); return cudaSuccess; }
    fprintf(stderr, "emu: cudaFree of an unknown pointer\n"); abort();
}
cudaError_t cudaMemset(void* p, int v, size_t bytes) { memset(p, v, bytes); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t) {
    if (emu::t_capture.on) { emu::t_capture.graph->nodes.push_back([=]() { memset(p, v, bytes); }); return cudaSuccess; }      // a memset node of the captured graph
    memset(p, v, bytes); return cudaSuccess;
}
struct EmuEvent { int id; };
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new EmuEvent{0}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new EmuEvent{0}; return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventRecordWithFlags(cudaEvent_t, cudaStream_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t bytes, cudaMemcpyKind) { memmove(d, s, bytes); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t bytes, cudaMemcpyKind, cudaStream_t) {
    if (emu::t_capture.on) { emu::t_capture.graph->nodes.push_back([=]() { memmove(d, s, bytes); }); return cudaSuccess; }     // a memcpy node: pointers baked, data read at replay
    memmove(d, s, bytes); return cudaSuccess;
}
cudaError_t cudaMemcpy2D(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind) {
    for (size_t r = 0; r < height; r++) memmove((char*)d + r * dpitch, (const char*)s + r * spitch, width);
    return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy2D(d, dpitch, s, spitch, width, height, k); }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* st, unsigned) { *st = new EmuStream{1}; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t st) { delete st; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { emu::t_capture.on = true; emu::t_capture.graph = new EmuGraph(); return cudaSuccess; }
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = emu::t_capture.graph; emu::t_capture.on = false; emu::t_capture.graph = nullptr; return cudaSuccess; }
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long) { *e = new EmuGraph(*g); return cudaSuccess; }
cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { delete e; return cudaSuccess; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t) { for (auto& n : e->nodes) n(); return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulation error"; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 148; return cudaSuccess; }
