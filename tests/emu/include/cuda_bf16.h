// bfloat16 storage type with round-to-nearest-even conversion (what __float2bfloat16_rn does on the device).
#pragma once
#include <cstdint>
#include <cstring>
struct __nv_bfloat16 { uint16_t x; };
inline __nv_bfloat16 __float2bfloat16_rn(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    __nv_bfloat16 o;
    if ((u & 0x7fffffffu) > 0x7f800000u) { o.x = (uint16_t)((u >> 16) | 0x40u); return o; }       // nan stays nan
    u += 0x7fffu + ((u >> 16) & 1u);
    o.x = (uint16_t)(u >> 16); return o;
}
inline float __bfloat162float(__nv_bfloat16 b) { uint32_t u = (uint32_t)b.x << 16; float f; memcpy(&f, &u, 4); return f; }
inline unsigned short __bfloat16_as_ushort(__nv_bfloat16 b) { return b.x; }
