// BC4 / BC5 UNorm kernels: a persistent grid of 8-warp CTAs; each CTA stages a tile of 64
// blocks in shared memory with coalesced 16-byte loads, each warp encodes one block at a
// time (bc4_device.cuh), and the packed tile is written back with 16-byte stores.
//
// Replaces Bc4Converter::compressBlock / Bc5Converter::compressBlock, lib/src/S3tcConverter.cpp:400-429, :453-490:
// UNorm bit-exact with rgbcx; SNorm (Compressonator in the reference) by the same search in a biased domain.
#include "bc4_device.cuh"
#include "kernels.h"

namespace cfx {

constexpr int kTile = 64;

template <int CHANNELS, bool SIGNED>
__global__ void __launch_bounds__(kThreads) bc45_kernel(const EncodeParams p, uint32_t radius, uint32_t hq)
{
    __shared__ __align__(16) uint32_t s_px[kTile*16];
    __shared__ __align__(16) uint32_t s_out[kTile*2*CHANNELS];
    const uint32_t tiles = (p.total_blocks + kTile - 1)/kTile;
    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        uint32_t first = tile*kTile;
        uint32_t n = min(static_cast<uint32_t>(kTile), p.total_blocks - first);
        __syncthreads();
        if (SIGNED) stage_tile_s8(s_px, p, first, n); else stage_tile_u8(s_px, p, first, n);
        __syncthreads();
        for (uint32_t b = warp_id(); b < n; b += kWarps) {
#pragma unroll
            for (int c = 0; c < CHANNELS; ++c) {
                uint2 r = bc4_encode_warp<SIGNED>(s_px + b*16, c, radius, hq != 0);
                if (lane_id() == 0) {
                    s_out[(b*CHANNELS + c)*2] = r.x;
                    s_out[(b*CHANNELS + c)*2 + 1] = r.y;
                }
            }
        }
        __syncthreads();
        store_tile(p, s_out, first, n);
    }
}

int launch_bc45(const EncodeParams& p, cudaStream_t stream)
{
    static const uint32_t radii[5] = {3, 3, 5, 16, 32};   // getSearchRadius, S3tcConverter.cpp:80-95
    uint32_t radius = radii[p.quality];
    uint32_t hq = p.quality > 1;                          // Quality <= Low -> encode_bc4 / encode_bc5
    uint32_t tiles = (p.total_blocks + kTile - 1)/kTile;
    const bool sn = p.type == 1;                          // Texture::Type::SNorm
    const void* k = p.format == 33 ? (sn ? reinterpret_cast<const void*>(&bc45_kernel<1, true>) : reinterpret_cast<const void*>(&bc45_kernel<1, false>))
                                   : (sn ? reinterpret_cast<const void*>(&bc45_kernel<2, true>) : reinterpret_cast<const void*>(&bc45_kernel<2, false>));
    const uint32_t grid = min(tiles, persistent_ctas(k, kThreads));
    void* args[] = {const_cast<EncodeParams*>(&p), &radius, &hq};
    if (cudaLaunchKernel(k, dim3(grid), dim3(kThreads), args, 0, stream) != cudaSuccess) return -4;
    return 1;
}

} // namespace cfx
