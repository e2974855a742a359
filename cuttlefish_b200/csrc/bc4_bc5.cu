// BC4 / BC5 UNorm kernels: a persistent grid of 8-warp CTAs; each CTA stages a tile of 64
// blocks in shared memory with coalesced 16-byte loads, each warp encodes one block at a
// time (bc4_device.cuh), and the packed tile is written back with 16-byte stores.
//
// Replaces Bc4Converter::compressBlock / Bc5Converter::compressBlock, lib/src/S3tcConverter.cpp:400-429, :453-490:
// UNorm bit-exact with rgbcx; SNorm (Compressonator in the reference) by the same search in a biased domain.
#include "bc4_device.cuh"
#include "kernels.h"

#include <cuda.h>
#include <cstdlib>

namespace cfx {

constexpr int kTile = 64;

// Resident CTAs per SM the kernels are compiled for (measured on 8192^2, radius 5): BC4 is fastest at 48 registers /
// 5 CTAs (17.5 GTexel/s; 17.3 at 6, 16.8 at 8 CTAs of 32 registers), BC5 at 40 registers / 6 CTAs (9.4; 8.9 at 4 CTAs
// of 59 registers, 8.7 at 8).
#ifndef CFX_BC45_MINB
#define CFX_BC45_MINB(channels) ((channels) == 2 ? 6 : 5)
#endif

template <int CHANNELS, bool SIGNED>
__global__ void __launch_bounds__(kThreads, CFX_BC45_MINB(CHANNELS)) bc45_kernel(const EncodeParams p, uint32_t radius, uint32_t hq)
{
    __shared__ __align__(16) uint32_t s_px[kTile*16];
    __shared__ __align__(16) uint32_t s_out[kTile*2*CHANNELS];
    __shared__ __align__(16) uint32_t s_tab[kWarps][kBc4TableWords];
    const uint32_t tiles = (p.total_blocks + kTile - 1)/kTile;
    const uint32_t inv = bc4_trial_inv(radius);
    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        uint32_t first = tile*kTile;
        uint32_t n = min(static_cast<uint32_t>(kTile), p.total_blocks - first);
        __syncthreads();
        if (SIGNED) stage_tile_s8(s_px, p, first, n); else stage_tile_u8(s_px, p, first, n);
        __syncthreads();
        for (uint32_t b = warp_id(); b < n; b += kWarps) {
#pragma unroll
            for (int c = 0; c < CHANNELS; ++c) {
                uint2 r = bc4_encode_warp<SIGNED>(s_px + b*16, c, radius, inv, hq != 0, s_tab[warp_id()]);
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

// ---- TMA-staged variant ---------------------------------------------------------------------------
// When the surface is RGBA8 with whole 4x4 blocks, a 16-byte aligned pitch and a block-row that is a multiple of the
// tile (64 blocks), a tile is a 256 x 4 texel box of the image: ONE cp.async.bulk.tensor (TMA) request per tile,
// double-buffered so that the next tile lands in shared memory while the warps search the current one.  The box
// arrives row-major (4 rows of 256 texels), which bc4_encode_warp reads through its row_stride argument.
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a bad tensor map must surface as an error, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, uint32_t x, uint32_t y)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}

} // namespace

template <int CHANNELS>
__global__ void __launch_bounds__(kThreads, CFX_BC45_MINB(CHANNELS)) bc45_tma_kernel(const EncodeParams p, const __grid_constant__ CUtensorMap tmap,
    uint32_t radius, uint32_t hq)
{
    __shared__ __align__(128) uint32_t s_px[2][kTile*16];          // two 256 x 4 texel boxes
    __shared__ __align__(16) uint32_t s_out[kTile*2*CHANNELS];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ __align__(16) uint32_t s_tab[kWarps][kBc4TableWords];
    const uint32_t tiles = p.total_blocks/kTile;                   // the launcher guarantees whole tiles
    const uint32_t tiles_x = p.blocks_x/kTile;
    const uint32_t inv = bc4_trial_inv(radius);
    constexpr uint32_t kBoxBytes = kTile*16*4;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](uint32_t tile, uint32_t buf) {
        const uint32_t ty = tile / tiles_x, tx = tile - ty*tiles_x;
        mbar_expect_tx(&s_bar[buf], kBoxBytes);
        tma_load_2d(s_px[buf], &tmap, &s_bar[buf], tx*kTile*4, ty*4);
    };
    if (threadIdx.x == 0 && blockIdx.x < tiles) issue(blockIdx.x, 0);
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1u;
        // the other buffer was last read in the previous iteration (every thread passed its trailing barrier)
        if (threadIdx.x == 0 && tile + gridDim.x < tiles) issue(tile + gridDim.x, buf ^ 1u);
        mbar_wait(&s_bar[buf], (it >> 1) & 1u);
        const uint32_t first = tile*kTile;
        for (uint32_t b = warp_id(); b < kTile; b += kWarps) {
#pragma unroll
            for (int c = 0; c < CHANNELS; ++c) {
                uint2 r = bc4_encode_warp<false>(s_px[buf] + b*4, c, radius, inv, hq != 0, s_tab[warp_id()], kTile*4);
                if (lane_id() == 0) {
                    s_out[(b*CHANNELS + c)*2] = r.x;
                    s_out[(b*CHANNELS + c)*2 + 1] = r.y;
                }
            }
        }
        __syncthreads();
        store_tile(p, s_out, first, kTile);
        __syncthreads();
    }
}

namespace {

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult res;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &res) != cudaSuccess ||
            res != cudaDriverEntryPointSuccess)
            ptr = nullptr;
        return reinterpret_cast<EncodeTiledFn>(ptr);
    }();
    return fn;
}

// 2-D map of the RGBA8 surface (u32 texels), box = 256 texels x 4 rows.
bool make_tile_map(const EncodeParams& p, CUtensorMap& map)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {p.width, p.height};
    const cuuint64_t strides[1] = {p.pitch};
    const cuuint32_t box[2] = {kTile*4, 4};
    const cuuint32_t elem[2] = {1, 1};
    return fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint8_t*>(p.src), dims, strides, box, elem,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

} // namespace

bool bc45_uses_tma(const EncodeParams& p)
{
    return p.type == 0 && (p.pitch >> 40) == 0 && p.src_format == SRC_RGBA8 && p.aligned16 && p.width % 4 == 0 && p.height % 4 == 0 &&
        p.blocks_x % kTile == 0 && p.total_blocks >= static_cast<uint32_t>(kTile);
}

int launch_bc45(const EncodeParams& p, cudaStream_t stream)
{
    static const uint32_t radii[5] = {3, 3, 5, 16, 32};   // getSearchRadius, S3tcConverter.cpp:80-95
    uint32_t radius = radii[p.quality];
    uint32_t hq = p.quality > 1;                          // Quality <= Low -> encode_bc4 / encode_bc5
    uint32_t tiles = (p.total_blocks + kTile - 1)/kTile;
    if (bc45_uses_tma(p)) {
        CUtensorMap map;
        if (make_tile_map(p, map)) {
            const void* kt = p.format == 33 ? reinterpret_cast<const void*>(&bc45_tma_kernel<1>) : reinterpret_cast<const void*>(&bc45_tma_kernel<2>);
            const uint32_t grid = min(p.total_blocks/kTile, persistent_ctas(kt, kThreads));
            void* targs[] = {const_cast<EncodeParams*>(&p), &map, &radius, &hq};
            if (cudaLaunchKernel(kt, dim3(grid), dim3(kThreads), targs, 0, stream) != cudaSuccess) return -4;
            return 1;
        }
    }
    const bool sn = p.type == 1;                          // Texture::Type::SNorm
    const void* k = p.format == 33 ? (sn ? reinterpret_cast<const void*>(&bc45_kernel<1, true>) : reinterpret_cast<const void*>(&bc45_kernel<1, false>))
                                   : (sn ? reinterpret_cast<const void*>(&bc45_kernel<2, true>) : reinterpret_cast<const void*>(&bc45_kernel<2, false>));
    const uint32_t grid = min(tiles, persistent_ctas(k, kThreads));
    void* args[] = {const_cast<EncodeParams*>(&p), &radius, &hq};
    if (cudaLaunchKernel(k, dim3(grid), dim3(kThreads), args, 0, stream) != cudaSuccess) return -4;
    return 1;
}

} // namespace cfx
