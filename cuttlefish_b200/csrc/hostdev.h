// Lets the lane-local encoder cores (*_core.cuh) compile both as device code and as plain host
// C++ for the developer-side emulators under tools/.  On the host the few CUDA intrinsics they
// use are restated with their documented integer/float semantics.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define CFX_HD __device__ __forceinline__
#define CFX_HD_NOINLINE static __device__ __noinline__
#define CFX_CONST __device__ __constant__
#define CFX_TABLE static __device__ const
#else
#include <algorithm>
#include <cmath>
#include <cstring>
#define CFX_HD inline
#define CFX_HD_NOINLINE inline
#define CFX_CONST static const
#define CFX_TABLE static const
struct float4 { float x, y, z, w; };
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
using std::max;
using std::min;
inline int __float2int_rn(float f) { return (int)lrintf(f); }
inline int __float2int_rz(float f) { return (int)f; }
inline float rsqrtf(float x) { return 1.0f/sqrtf(x); }
inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t __vabsdiffu4(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        int x = (a >> (8*i)) & 0xFF, y = (b >> (8*i)) & 0xFF;
        r |= (uint32_t)(x > y ? x - y : y - x) << (8*i);
    }
    return r;
}
inline uint32_t __dp4a(uint32_t a, uint32_t b, uint32_t c)
{
    for (int i = 0; i < 4; ++i) c += ((a >> (8*i)) & 0xFF)*((b >> (8*i)) & 0xFF);
    return c;
}
inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s)
{
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8*((s >> (4*i)) & 7))) & 0xFF) << (8*i);
    return r;
}
inline int __popc(uint32_t x) { return __builtin_popcount(x); }
inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
#endif
