// Shared device-side plumbing for the block encoders: kernel parameters, tile staging
// (coalesced 128-bit loads of RGBA8 / RGBA16F / RGBA32F texels into shared memory with the
// reference's clamp-to-edge gather) and coalesced stores of the packed blocks.
//
// Replaces, on the device, the per-block gather + quantise glue of
//   lib/src/S3tcConverter.cpp:242-255 (4x4 gather, min(coord, dim-1) clamp)
//   lib/src/S3tcConverter.cpp:97-111  (u8 = round(clamp01(v)*255), half away from zero)
//   lib/src/HalfFloat.h:96-134        (f32 -> f16 round-to-nearest-even)
//   lib/src/AstcConverter.cpp:210-222 (NxM gather with the same clamp)
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cfx {

enum { SRC_RGBA8 = 0, SRC_RGBA16F = 1, SRC_RGBA32F = 2 };

struct EncodeParams {
    const uint8_t* src;   // device pointer, row 0 = top
    uint8_t* dst;         // device pointer, blocks row-major
    uint64_t pitch;       // bytes
    uint32_t width, height;
    uint32_t src_format;
    uint32_t blocks_x, blocks_y, total_blocks;
    uint32_t block_w, block_h, block_bytes;
    uint32_t format, type, quality, alpha_type, color_mask, color_space;
    uint32_t aligned16;   // src pointer and pitch are 16-byte multiples
};

constexpr int kThreads = 256;          // 8 warps per CTA
constexpr int kWarps = kThreads/32;

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t warp_id() { return threadIdx.x >> 5; }

// round(clamp(v,0,1)*255) with std::round semantics (half away from zero); NaN -> 0 like the
// host's std::max/std::min based clamp followed by a cast would give on x86 for NaN inputs is
// undefined, so NaN is simply mapped to 0 here.
__device__ __forceinline__ uint32_t f32_to_unorm8(float v)
{
    v = fminf(fmaxf(v, 0.0f), 1.0f);
    return static_cast<uint32_t>(roundf(__fmul_rn(v, 255.0f)));
}

__device__ __forceinline__ uint32_t pack_rgba8(float4 v)
{
    return f32_to_unorm8(v.x) | (f32_to_unorm8(v.y) << 8) | (f32_to_unorm8(v.z) << 16) |
        (f32_to_unorm8(v.w) << 24);
}

__device__ __forceinline__ float4 unpack_rgba8_f32(uint32_t p)
{
    // k/255 with IEEE division: identical to the float the host holds for an 8-bit-snapped image.
    return make_float4(__fdiv_rn(static_cast<float>(p & 0xFF), 255.0f),
        __fdiv_rn(static_cast<float>((p >> 8) & 0xFF), 255.0f),
        __fdiv_rn(static_cast<float>((p >> 16) & 0xFF), 255.0f),
        __fdiv_rn(static_cast<float>(p >> 24), 255.0f));
}

__device__ __forceinline__ float4 load_texel_f32(const EncodeParams& p, uint32_t x, uint32_t y)
{
    const uint8_t* row = p.src + static_cast<uint64_t>(y)*p.pitch;
    if (p.src_format == SRC_RGBA32F)
        return __ldg(reinterpret_cast<const float4*>(row) + x);
    if (p.src_format == SRC_RGBA16F) {
        uint2 h = __ldg(reinterpret_cast<const uint2*>(row) + x);
        __half2 a = *reinterpret_cast<__half2*>(&h.x), b = *reinterpret_cast<__half2*>(&h.y);
        float2 fa = __half22float2(a), fb = __half22float2(b);
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    }
    return unpack_rgba8_f32(__ldg(reinterpret_cast<const uint32_t*>(row) + x));
}

__device__ __forceinline__ uint32_t load_texel_u8(const EncodeParams& p, uint32_t x, uint32_t y)
{
    if (p.src_format == SRC_RGBA8)
        return __ldg(reinterpret_cast<const uint32_t*>(p.src + static_cast<uint64_t>(y)*p.pitch) + x);
    return pack_rgba8(load_texel_f32(p, x, y));
}

// Stage the RGBA8 view of `nblocks` consecutive 4x4 blocks (linear block index first..first+n)
// into shared memory: s_px[b*16 + r*4 + c].  Work item = (texel row r, block b) so that
// consecutive threads read consecutive 16-byte pieces of the same image row.
// `stride` = words between consecutive blocks in s_px (a multiple of 4; 20 keeps the 4 or 8 blocks a
// warp works on at once in different banks).
__device__ __forceinline__ void stage_tile_u8(uint32_t* s_px, const EncodeParams& p, uint32_t first,
    uint32_t nblocks, uint32_t stride = 16)
{
    for (uint32_t i = threadIdx.x; i < nblocks*4; i += blockDim.x) {
        uint32_t r = i / nblocks, b = i - r*nblocks;
        uint32_t blk = first + b;
        uint32_t by = blk / p.blocks_x, bx = blk - by*p.blocks_x;
        uint32_t y = min(by*4 + r, p.height - 1);
        uint32_t x0 = bx*4;
        uint4 v;
        if (p.src_format == SRC_RGBA8 && p.aligned16 && x0 + 3 < p.width) {
            v = __ldg(reinterpret_cast<const uint4*>(p.src + static_cast<uint64_t>(y)*p.pitch) + bx);
        } else {
            uint32_t xm = p.width - 1;
            v.x = load_texel_u8(p, min(x0, xm), y);
            v.y = load_texel_u8(p, min(x0 + 1, xm), y);
            v.z = load_texel_u8(p, min(x0 + 2, xm), y);
            v.w = load_texel_u8(p, min(x0 + 3, xm), y);
        }
        *reinterpret_cast<uint4*>(s_px + b*stride + r*4) = v;
    }
}

// SNORM view for BC4S / BC5S: per channel round(clamp(v,-1,1)*127) (std::round, like Bc4Converter::compressBlock,
// lib/src/S3tcConverter.cpp:404-409) biased by +128 into a byte: s_px[b*stride + r*4 + c].
__device__ __forceinline__ uint32_t f32_to_snorm8_biased(float v)
{
    v = fminf(fmaxf(v, -1.0f), 1.0f);
    return static_cast<uint32_t>(static_cast<int>(roundf(__fmul_rn(v, 127.0f))) + 128);
}

__device__ __forceinline__ void stage_tile_s8(uint32_t* s_px, const EncodeParams& p, uint32_t first, uint32_t nblocks)
{
    for (uint32_t i = threadIdx.x; i < nblocks*16; i += blockDim.x) {
        uint32_t t = i / nblocks, b = i - t*nblocks;
        uint32_t r = t >> 2, c = t & 3;
        uint32_t blk = first + b;
        uint32_t by = blk / p.blocks_x, bx = blk - by*p.blocks_x;
        uint32_t y = min(by*4 + r, p.height - 1);
        uint32_t x = min(bx*4 + c, p.width - 1);
        const float4 f = load_texel_f32(p, x, y);
        s_px[b*16 + t] = f32_to_snorm8_biased(f.x) | (f32_to_snorm8_biased(f.y) << 8) | (f32_to_snorm8_biased(f.z) << 16) |
            (f32_to_snorm8_biased(f.w) << 24);
    }
}

// Same for the float view: s_px[b*16 + r*4 + c] as float4 (ETC / BC6H inputs).
__device__ __forceinline__ void stage_tile_f32(float4* s_px, const EncodeParams& p, uint32_t first,
    uint32_t nblocks)
{
    for (uint32_t i = threadIdx.x; i < nblocks*16; i += blockDim.x) {
        uint32_t t = i / nblocks, b = i - t*nblocks;   // t = texel within block (r*4+c)
        // re-order so consecutive threads walk x: item -> (r, b, c)
        uint32_t r = t >> 2, c = t & 3;
        uint32_t blk = first + b;
        uint32_t by = blk / p.blocks_x, bx = blk - by*p.blocks_x;
        uint32_t y = min(by*4 + r, p.height - 1);
        uint32_t x = min(bx*4 + c, p.width - 1);
        s_px[b*16 + t] = load_texel_f32(p, x, y);
    }
}

// Copy the CTA's packed output tile from shared memory to global memory with 16-byte stores.
__device__ __forceinline__ void store_tile(const EncodeParams& p, const uint32_t* s_out,
    uint32_t first, uint32_t nblocks)
{
    uint32_t words = nblocks*(p.block_bytes/4);
    uint32_t* dst = reinterpret_cast<uint32_t*>(p.dst + static_cast<uint64_t>(first)*p.block_bytes);
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        uint32_t vec = words/4;
        for (uint32_t i = threadIdx.x; i < vec; i += blockDim.x)
            reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(s_out)[i];
        for (uint32_t i = vec*4 + threadIdx.x; i < words; i += blockDim.x)
            dst[i] = s_out[i];
    } else {
        for (uint32_t i = threadIdx.x; i < words; i += blockDim.x)
            dst[i] = s_out[i];
    }
}

} // namespace cfx
