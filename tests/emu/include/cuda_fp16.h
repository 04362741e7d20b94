// IEEE binary16 storage type with round-to-nearest-even conversion (what __float2half_rn / __half2float do on the device).
#pragma once
#include "cuda_runtime.h"
struct __half { uint16_t x; };
struct alignas(4) __half2 { __half x, y; };
inline __half __float2half_rn(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    const uint32_t sign = (u >> 16) & 0x8000u; u &= 0x7fffffffu;
    uint16_t h;
    if (u >= 0x7f800000u) h = (uint16_t)(0x7c00u | (u > 0x7f800000u ? 0x200u : 0));                  // inf / nan
    else if (u >= 0x477ff000u) h = 0x7c00u;                                                           // rounds to inf (>= 65520)
    else if (u >= 0x38800000u) {                                                                      // normal half
        const uint32_t m = u - 0x38000000u;                                                           // rebias exponent by 112
        uint32_t r = m >> 13; const uint32_t rem = m & 0x1fffu;
        if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) r++;
        h = (uint16_t)r;
    } else if (u >= 0x33000000u) {                                                                    // subnormal half
        const int e = (int)(u >> 23);                                                                 // 102 .. 112
        const uint32_t mant = (u & 0x7fffffu) | 0x800000u;
        const int shift = 126 - e;                                                                    // 14 .. 24
        uint32_t r = mant >> shift; const uint32_t rem = mant & ((1u << shift) - 1u), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (r & 1u))) r++;
        h = (uint16_t)r;
    } else h = 0;
    __half o; o.x = (uint16_t)(h | sign); return o;
}
inline float __half2float(__half hh) {
    const uint32_t h = hh.x, sign = (h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    uint32_t u;
    if (e == 0) {
        if (m == 0) u = sign;
        else { int s = 0; uint32_t mm = m; while (!(mm & 0x400u)) { mm <<= 1; s++; } u = sign | ((uint32_t)(113 - s) << 23) | ((mm & 0x3ffu) << 13); }
    } else if (e == 31) u = sign | 0x7f800000u | (m << 13);
    else u = sign | ((e + 112u) << 23) | (m << 13);
    float f; memcpy(&f, &u, 4); return f;
}
inline float2 __half22float2(__half2 h) { return make_float2(__half2float(h.x), __half2float(h.y)); }
inline unsigned short __half_as_ushort(__half h) { return h.x; }
inline __half __ushort_as_half(unsigned short u) { __half h; h.x = u; return h; }
inline __half2 __floats2half2_rn(float a, float b) { __half2 r; r.x = __float2half_rn(a); r.y = __float2half_rn(b); return r; }
inline float __low2float(__half2 h) { return __half2float(h.x); }
inline float __high2float(__half2 h) { return __half2float(h.y); }
