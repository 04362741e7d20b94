// Emulation stand-in for csrc/ptx.cuh: the SIMT translation units only use the programmatic-dependent-launch fences from it, which have
// no functional effect when launches are executed one after the other.
#pragma once
inline void pdl_wait() {}
inline void pdl_launch_dependents() {}
