// BC1_RGB / BC1_RGBA / BC2 / BC3 kernels.  A warp owns 32 consecutive blocks: it stages their
// texels in shared memory with 16-byte row loads (32 lanes x 16 B = 512 contiguous bytes per image
// row), every LANE encodes the colour half of ITS block (bc1_core.cuh), and for BC3 the warp then
// walks the 32 blocks cooperatively for the alpha half with the exact BC4 search of bc4_device.cuh
// (bit-identical to rgbcx::encode_bc4_hq, as Bc3Converter uses it through encode_bc3_hq).
//
// Replaces Bc1Converter / Bc1AConverter / Bc2Converter / Bc3Converter::compressBlock,
// lib/src/S3tcConverter.cpp:263-376.  Colour halves: PSNR parity (see bc1_core.cuh); BC3/BC2 alpha
// halves: bit-exact (rgbcx.cpp:2730-2884; packBc2Alpha, S3tcConverter.cpp:131-143).
#include "bc1_core.cuh"
#ifdef CFX_HAVE_RGBCX_TABLES
#include "bc1_exact.cuh"
#endif
#include "bc4_device.cuh"
#include "kernels.h"

namespace cfx {

namespace {
constexpr int kBc1Warps = 8;
constexpr int kBlkStride = 20;     // words per block in shared memory (16 texels + pad: no bank conflicts)
}

template <int FORMAT>   // 29 BC1_RGB, 30 BC1_RGBA, 31 BC2, 32 BC3
__global__ void __launch_bounds__(kBc1Warps*32) bc123_kernel(const EncodeParams p, int descent, uint32_t radius, uint32_t hq,
    bool exact)
{
    __shared__ __align__(16) uint32_t s_px[kBc1Warps][32*kBlkStride];
    __shared__ __align__(16) uint32_t s_tab[FORMAT == 32 ? kBc1Warps : 1][kBc4TableWords];
    const uint32_t lane = lane_id(), warp = warp_id();
    uint32_t* sp = s_px[warp];
    const uint32_t groups = (p.total_blocks + 31)/32;
    const uint32_t inv = bc4_trial_inv(radius);
    for (uint32_t grp = blockIdx.x*kBc1Warps + warp; grp < groups; grp += gridDim.x*kBc1Warps) {
        const uint32_t first = grp*32;
        const uint32_t nblk = min(32u, p.total_blocks - first);
        __syncwarp();
        // stage: row r of block (first + lane)
        {
            const uint32_t blk = min(first + lane, p.total_blocks - 1);
            const uint32_t by = blk / p.blocks_x, bx = blk - by*p.blocks_x;
#pragma unroll
            for (uint32_t r = 0; r < 4; ++r) {
                const uint32_t y = min(by*4 + r, p.height - 1), x0 = bx*4;
                uint4 v;
                if (p.src_format == SRC_RGBA8 && p.aligned16 && x0 + 3 < p.width) {
                    v = __ldg(reinterpret_cast<const uint4*>(p.src + static_cast<uint64_t>(y)*p.pitch) + bx);
                } else {
                    const uint32_t xm = p.width - 1;
                    v.x = load_texel_u8(p, min(x0, xm), y); v.y = load_texel_u8(p, min(x0 + 1, xm), y);
                    v.z = load_texel_u8(p, min(x0 + 2, xm), y); v.w = load_texel_u8(p, min(x0 + 3, xm), y);
                }
                *reinterpret_cast<uint4*>(sp + lane*kBlkStride + r*4) = v;
            }
        }
        __syncwarp();
        uint32_t px[16];
        // like the reference (Bc1/Bc2/Bc3Converter never look at the colour mask) all channels are encoded
#pragma unroll
        for (int i = 0; i < 16; ++i) px[i] = sp[lane*kBlkStride + i];
        uint32_t flags = 0;
        if (FORMAT == 29) flags = bc1::kAllow3 | bc1::kAllowBlack;
        if (FORMAT == 30) flags = bc1::kAllow3 | bc1::kPunchThrough;
        uint2 color;
#ifdef CFX_HAVE_RGBCX_TABLES
        // Every Texture::Quality: byte-exact rgbcx levels 0 / 4 / 9 / 13 / 18 (bc1_exact.cuh).  BC1_RGBA blocks with a texel of
        // alpha < 0.5 go through libsquish in the reference (S3tcConverter.cpp:283-330); those keep our
        // punch-through search.
        bool transparent = false;
        if (FORMAT == 30) {
#pragma unroll
            for (int i = 0; i < 16; ++i) transparent = transparent || (px[i] >> 24) < 128u;
        }
        if (exact && !transparent)
            color = rgbcx9::encode_bc1_exact(px, p.quality, FORMAT == 29 || FORMAT == 30, FORMAT == 29);
        else
#endif
            color = bc1::encode_color_block(px, flags, descent);
        if (FORMAT == 29 || FORMAT == 30) {
            if (lane < nblk) reinterpret_cast<uint2*>(p.dst)[first + lane] = color;
        } else if (FORMAT == 31) {
            // explicit 4-bit alpha: round(a * 15/255), half away from zero (packBc2Alpha)
            uint32_t a_lo = 0, a_hi = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint32_t a = px[i] >> 24;
                const uint32_t q = static_cast<uint32_t>(roundf(__fmul_rn(static_cast<float>(a), 15.0f/255.0f))) & 15u;
                if (i < 8) a_lo |= q << (4*i); else a_hi |= q << (4*(i - 8));
            }
            if (lane < nblk) reinterpret_cast<uint4*>(p.dst)[first + lane] = make_uint4(a_lo, a_hi, color.x, color.y);
        } else {
            uint2 mine = make_uint2(0, 0);
            for (uint32_t b = 0; b < nblk; ++b) {
                const uint2 a = bc4_encode_warp(sp + b*kBlkStride, 3, radius, inv, hq != 0, s_tab[FORMAT == 32 ? warp : 0]);
                if (b == lane) mine = a;
            }
            if (lane < nblk) reinterpret_cast<uint4*>(p.dst)[first + lane] = make_uint4(mine.x, mine.y, color.x, color.y);
        }
    }
}

// Whether BC1-family colour blocks are byte-identical to rgbcx at this quality: all five levels (rgbcx levels 0, 4, 9,
// 13, 18) are restated, but only when the reference's tables were generated into the build.
bool bc1_color_is_exact(uint32_t quality)
{
#ifdef CFX_HAVE_RGBCX_TABLES
    return quality <= 4;
#else
    (void)quality;
    return false;
#endif
}

int launch_bc123(const EncodeParams& p, cudaStream_t stream)
{
    static const uint32_t radii[5] = {3, 3, 5, 16, 32};   // getSearchRadius, S3tcConverter.cpp:80-95
    static const int descents[5] = {0, 1, 2, 3, 4};
    const uint32_t radius = radii[p.quality], hq = p.quality > 1;
    const int descent = descents[p.quality];
    const uint32_t groups = (p.total_blocks + 31)/32;
    const uint32_t ctas = (groups + kBc1Warps - 1)/kBc1Warps;
    const void* k;
    switch (p.format) {
        case 29: k = reinterpret_cast<const void*>(&bc123_kernel<29>); break;
        case 30: k = reinterpret_cast<const void*>(&bc123_kernel<30>); break;
        case 31: k = reinterpret_cast<const void*>(&bc123_kernel<31>); break;
        default: k = reinterpret_cast<const void*>(&bc123_kernel<32>); break;
    }
    const uint32_t grid = min(ctas, persistent_ctas(k, kBc1Warps*32));
    bool exact = bc1_color_is_exact(p.quality);
    void* args[] = {const_cast<EncodeParams*>(&p), const_cast<int*>(&descent), const_cast<uint32_t*>(&radius),
        const_cast<uint32_t*>(&hq), &exact};
    if (cudaLaunchKernel(k, dim3(grid), dim3(kBc1Warps*32), args, 0, stream) != cudaSuccess) return -4;
    return 1;
}

} // namespace cfx
