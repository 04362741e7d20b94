#pragma once
#include <cstdint>
struct __nv_bfloat16 { uint16_t x; };      // only named by headers; the emulated translation units do no bf16 arithmetic
