// Emulation stand-in for csrc/ptx.cuh (the part the SIMT translation units use): programmatic-dependent-launch fences (no functional
// effect when launches run one after the other) and the mbarrier / 1-D bulk-copy pair of the kNN scan's per-warp rings.
// An mbarrier is modelled as the number of completed phases: a bulk copy is performed at issue time and completes the phase that
// `mbar_expect_tx` armed; `mbar_wait(bar, parity)` returns once the phase of that parity has completed, yielding to the other threads
// of the block until then -- the protocol (who arms, who waits, which parity) is checked, the asynchrony is not.
#pragma once
#include <stdint.h>
#include <string.h>
inline void pdl_wait() {}
inline void pdl_launch_dependents() {}
inline void fence_barrier_init() {}
inline void fence_proxy_async() {}
inline void mbar_init(uint64_t* bar, uint32_t) { *bar = 0; }
inline void mbar_expect_tx(uint64_t*, uint32_t) {}
inline void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    if ((bytes & 15u) || ((uintptr_t)smem_dst & 15u) || ((uintptr_t)gsrc & 15u)) { fprintf(stderr, "emu: cp.async.bulk needs 16-byte sizes and addresses\n"); abort(); }
    memcpy(smem_dst, gsrc, bytes);
    *(volatile uint64_t*)bar = *bar + 1;
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    long spins = 0;
    while (((*(volatile uint64_t*)bar) & 1u) == parity) {              // the phase with this parity has not completed yet
        emu::yield();
        if (++spins > 100000000L) { fprintf(stderr, "emu: mbarrier wait never completed\n"); abort(); }
    }
}
