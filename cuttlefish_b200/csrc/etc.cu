// ETC1 / ETC2 RGB / ETC2 RGBA8 kernels: ONE LANE OWNS ONE 4x4 BLOCK (etc_core.cuh); a warp encodes
// 32 consecutive blocks, so its texel loads walk contiguous bytes of each image row and its output
// is one contiguous store.  Texels are kept as floats (0..255) in shared memory, lane-interleaved:
// like the reference (lib/src/EtcConverter.cpp:145, float RGBA into Etc::Image) the encoder sees
// unquantised values when the source is RGBA16F / RGBA32F.
//
// Replaces EtcConverter::process, lib/src/EtcConverter.cpp:120-152, for ETC1, ETC2_R8G8B8, ETC2_R8G8B8A1,
// ETC2_R8G8B8A8 and EAC_R11 / EAC_R11G11 (UNorm and SNorm).
#include "common.cuh"
#include "etc_core.cuh"
#include "etc1_exact.cuh"
#include "kernels.h"

namespace cfx {

namespace { constexpr int kEtcWarps = 4; }

// PERC: the sRGB instantiation of the ETC2 colour search (REC709 error, etc_core.cuh) -- its own kernel, so that the linear
// one carries none of its code (both searches in one kernel cost the linear path 5 %)
template <int FORMAT, bool SIGNED = false, bool PERC = false>   // 37 ETC1, 38 ETC2 RGB, 39 ETC2 RGB8A1, 40 ETC2 RGBA8, 41 EAC R11, 42 EAC RG11
__global__ void __launch_bounds__(kEtcWarps*32) etc_kernel(const EncodeParams p, int rounds, int alpha_radius, bool exact)
{
    __shared__ float s_x[kEtcWarps][16*4*32];
    const uint32_t lane = lane_id(), warp = warp_id();
    float* xs = s_x[warp];
    const uint32_t groups = (p.total_blocks + 31)/32;
    for (uint32_t grp = blockIdx.x*kEtcWarps + warp; grp < groups; grp += gridDim.x*kEtcWarps) {
        const uint32_t blk = grp*32 + lane;
        const bool live = blk < p.total_blocks;
        const uint32_t b = live ? blk : p.total_blocks - 1;
        const uint32_t by = b / p.blocks_x, bx = b - by*p.blocks_x;
        if (FORMAT == 37 && exact) {
            // byte-exact etc2comp (etc1_exact.cuh): texels in the reference's column-major
            // block order, [0,1] floats, alpha forced to 1, texels outside the image marked with NaN alpha
            etc1x::Px src[16];
#pragma unroll
            for (uint32_t x = 0; x < 4; ++x)
#pragma unroll
                for (uint32_t y = 0; y < 4; ++y) {
                    const uint32_t sx = bx*4 + x, sy = by*4 + y;
                    etc1x::Px q;
                    if (sx >= p.width || sy >= p.height) { q.r = q.g = q.b = 0.0f; q.a = __int_as_float(0x7FC00000); }
                    else {
                        const float4 f = load_texel_f32(p, sx, sy);
                        q.r = etc1x::clamp01(f.x); q.g = etc1x::clamp01(f.y); q.b = etc1x::clamp01(f.z); q.a = 1.0f;
                    }
                    src[x*4 + y] = q;
                }
            // etc2comp's effort for the quality level (lib/src/EtcConverter.cpp:34-51): up to Normal only encoding iteration 0
            // runs, High (70) goes on to the radius-1 tries and the first degenerate set, Highest (100) runs all nine
            const float effort = p.quality == 3u ? 70.0f : (p.quality >= 4u ? 100.0f : 40.0f);
            // sRGB textures: etc2comp's REC709 metric instead of RGBX (lib/src/EtcConverter.cpp:61-64)
            const uint2 color = etc1x::encode_etc1_exact(src, effort, p.color_space == 1u);
            if (live) reinterpret_cast<uint2*>(p.dst)[blk] = color;
            continue;
        }
        // texels of a partial edge block that lie outside the image: clamp-to-edge replicas in xs, no weight in the search
        // (EtcConverter hands etc2comp a smaller image for these blocks, lib/src/EtcConverter.cpp:122-150)
        uint32_t vm = 0;
        for (uint32_t t = 0; t < 16; ++t) {
            if (bx*4 + (t & 3) < p.width && by*4 + (t >> 2) < p.height) vm |= 1u << t;
            const uint32_t x = min(bx*4 + (t & 3), p.width - 1), y = min(by*4 + (t >> 2), p.height - 1);
            float4 v;
            if (p.src_format == SRC_RGBA8) {
                const uint32_t q = __ldg(reinterpret_cast<const uint32_t*>(p.src + static_cast<uint64_t>(y)*p.pitch) + x);
                v = make_float4(static_cast<float>(q & 0xFF), static_cast<float>((q >> 8) & 0xFF), static_cast<float>((q >> 16) & 0xFF),
                    static_cast<float>(q >> 24));
            } else {
                const float4 f = load_texel_f32(p, x, y);
                const float lowest = SIGNED ? -1.0f : 0.0f;      // signed EAC keeps [-1,1] (EtcConverter.cpp:139-143)
                v = make_float4(fminf(fmaxf(f.x, lowest), 1.0f)*255.0f, fminf(fmaxf(f.y, lowest), 1.0f)*255.0f,
                    fminf(fmaxf(f.z, 0.0f), 1.0f)*255.0f, fminf(fmaxf(f.w, 0.0f), 1.0f)*255.0f);
            }
            // EtcConverter never looks at the colour mask or the alpha type: all four channels as they are
            etc::px(xs, lane, t, 0) = v.x; etc::px(xs, lane, t, 1) = v.y; etc::px(xs, lane, t, 2) = v.z; etc::px(xs, lane, t, 3) = v.w;
        }
        if (FORMAT == 41 || FORMAT == 42) {
            const uint2 r = etc::encode_eac_r11<SIGNED>(xs, lane, 0, alpha_radius, vm);
            if (FORMAT == 41) {
                if (live) reinterpret_cast<uint2*>(p.dst)[blk] = r;
            } else {
                const uint2 g = etc::encode_eac_r11<SIGNED>(xs, lane, 1, alpha_radius, vm);
                if (live) reinterpret_cast<uint4*>(p.dst)[blk] = make_uint4(r.x, r.y, g.x, g.y);
            }
            continue;
        }
        // (PERC, sRGB textures: the reference's perceptual REC709 colour error instead of plain squared RGB distance)
        const uint2 color = FORMAT == 39 ? etc::encode_color_a1<PERC>(xs, lane, rounds, vm) : etc::encode_color<PERC>(xs, lane, FORMAT != 37, rounds, vm);
        if (FORMAT == 40) {
            const uint2 alpha = etc::encode_eac_alpha(xs, lane, alpha_radius, vm);
            if (live) reinterpret_cast<uint4*>(p.dst)[blk] = make_uint4(alpha.x, alpha.y, color.x, color.y);
        } else {
            if (live) reinterpret_cast<uint2*>(p.dst)[blk] = color;
        }
    }
}

bool etc1_is_exact(uint32_t quality) { return quality <= 4; }       // every level, linear and sRGB

int launch_etc(const EncodeParams& p, cudaStream_t stream)
{
    static const int rounds_by_quality[5] = {0, 1, 1, 3, 5};
    static const int radius_by_quality[5] = {0, 1, 2, 3, 4};
    const int rounds = rounds_by_quality[p.quality], radius = radius_by_quality[p.quality];
    const uint32_t groups = (p.total_blocks + 31)/32;
    const uint32_t ctas = (groups + kEtcWarps - 1)/kEtcWarps;
    const bool sn = p.type == 1;                          // Texture::Type::SNorm
    const bool srgb = p.color_space == 1;                 // the REC709 instantiation of the colour search
    const void* k = nullptr;
    switch (p.format) {
        case 37: k = srgb ? reinterpret_cast<const void*>(&etc_kernel<37, false, true>) : reinterpret_cast<const void*>(&etc_kernel<37>); break;
        case 38: k = srgb ? reinterpret_cast<const void*>(&etc_kernel<38, false, true>) : reinterpret_cast<const void*>(&etc_kernel<38>); break;
        case 39: k = srgb ? reinterpret_cast<const void*>(&etc_kernel<39, false, true>) : reinterpret_cast<const void*>(&etc_kernel<39>); break;
        case 40: k = srgb ? reinterpret_cast<const void*>(&etc_kernel<40, false, true>) : reinterpret_cast<const void*>(&etc_kernel<40>); break;
        case 41: k = sn ? reinterpret_cast<const void*>(&etc_kernel<41, true>) : reinterpret_cast<const void*>(&etc_kernel<41, false>); break;
        case 42: k = sn ? reinterpret_cast<const void*>(&etc_kernel<42, true>) : reinterpret_cast<const void*>(&etc_kernel<42, false>); break;
        default: return -2;
    }
    const uint32_t grid = min(ctas, persistent_ctas(k, kEtcWarps*32));
    // ETC1 is the byte-exact restatement at every quality level, in both colour spaces
    bool exact = etc1_is_exact(p.quality);
    void* args[] = {const_cast<EncodeParams*>(&p), const_cast<int*>(&rounds), const_cast<int*>(&radius), &exact};
    if (cudaLaunchKernel(k, dim3(grid), dim3(kEtcWarps*32), args, 0, stream) != cudaSuccess) return -4;
    return 1;
}

} // namespace cfx
