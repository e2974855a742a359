// BC7 UNorm encoder for sm_100a -- our own search, not a port of bc7enc / bc7e.
//
// Replaces Bc7Converter::compressBlock (lib/src/S3tcConverter.cpp:593-646), whose CPU encoder
// in the ISPC=0 configuration is bc7enc_compress_block (lib/bc7enc_rdo/bc7enc.cpp:2402-2438:
// modes 6+1 for opaque blocks, 6+5+7 for blocks with alpha, one estimated partition refined).
// Parity contract: RGB PSNR >= reference - 0.1 dB on the same input (tests/test_parity_gpu.py).
//
// Mapping: a group of G lanes (G = 8, 16 or 32, by quality) owns one 4x4 block, so a warp
// encodes 32/G blocks at once.  Work inside a group is organised as LANE = CANDIDATE:
//   phase 1  every lane scores 64/G of the 64 two-subset partition shapes: per-subset 4x4
//            scatter matrices, score = sum_s (trace - lambda_max) = squared distance of the
//            subset's texels from their best-fit line (3 power iterations);
//   phase 2  the G-1 best shapes are picked with butterfly min-reductions over
//            (score | shape) keys;
//   phase 3  lane 0 takes mode 6, the other lanes take mode 1 / mode 3 (opaque) or mode 7 (alpha) on the ranked
//            shapes, or modes 5 / 4 (two index sets, channel rotation: fit_dual); each lane runs the whole fit for ITS candidate:
//            PCA endpoints -> p-bit aware quantisation -> index assignment with exact integer
//            palette error (packed 2x16 IMAD interpolation, vabsdiff4 + dp4a SSE) -> two rounds
//            of least-squares endpoint refinement, each subset keeping its own best;
//   phase 4  butterfly argmin over (SSE | lane) picks the winner lane, which packs the 128 bits.
// Texels come from a shared-memory tile staged with coalesced 16-byte loads; packed blocks go
// back through shared memory and leave as 16-byte stores.
#include "bc7_core.cuh"
#include "common.cuh"
#include "kernels.h"

namespace cfx {

using namespace bc7;

namespace {

constexpr int kTileBc7 = 64;
constexpr int kPxStride = 20;     // words per block in s_px: blocks of one warp land in different banks
constexpr int kPxfStride = 17;    // float4 per block in s_pxf, same reason

template <int G>
__device__ __forceinline__ uint32_t group_min(uint32_t v)
{
#pragma unroll
    for (int o = G/2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}

} // namespace

// Candidate sets per lanes-per-block G (index 0: G=4, 1: G=8, 2: G=16, 3: G=32) and per opaque / alpha.
#define C CFX_BC7_CAND
// Modes 4 / 5 entries: C(mode, rotation, index mode, rounds).  Blocks with alpha give one lane to mode 5 at Normal and
// three to modes 5 / 4 at High (an alpha channel that runs independently of the colour is what they are for); the
// widest set also tries the rotations, on opaque blocks too (one colour channel on the scalar index set).
__device__ __constant__ uint16_t kBc7Cand[4][2][32] = {
    {{C(6,0,0,2), C(1,0,0,2), C(3,0,0,2), C(1,1,0,2)},
     {C(6,0,0,2), C(7,0,0,2), C(7,1,0,2), C(5,0,0,2)}},
    {{C(6,0,0,2), C(6,0,1,2), C(1,0,0,2), C(3,0,0,2), C(1,0,1,2), C(3,0,1,2), C(1,1,0,2), C(3,1,0,2)},
     {C(6,0,0,2), C(6,0,1,2), C(7,0,0,2), C(7,0,1,2), C(7,1,0,2), C(5,0,0,2), C(4,0,0,2), C(4,0,1,2)}},
    {{C(6,0,0,2), C(6,0,1,2), C(1,0,0,2), C(3,0,0,2), C(1,0,1,2), C(3,0,1,2), C(1,1,0,2), C(3,1,0,2),
      C(1,1,1,2), C(3,1,1,2), C(1,2,0,2), C(3,2,0,2), C(1,2,1,2), C(3,2,1,2), C(5,1,0,2), C(5,3,0,2)},
     {C(6,0,0,2), C(6,0,1,2), C(7,0,0,2), C(7,0,1,2), C(7,1,0,2), C(7,1,1,2), C(7,2,0,2), C(7,2,1,2),
      C(7,3,0,2), C(7,3,1,2), C(5,0,0,2), C(4,0,0,2), C(4,0,1,2), C(5,1,0,2), C(5,2,0,2), C(5,3,0,2)}},
    {{C(6,0,0,2), C(6,0,1,2), C(1,0,0,2), C(3,0,0,2), C(1,0,1,2), C(3,0,1,2), C(1,1,0,2), C(3,1,0,2),
      C(1,1,1,2), C(3,1,1,2), C(1,2,0,2), C(3,2,0,2), C(1,2,1,2), C(3,2,1,2), C(1,3,0,2), C(3,3,0,2),
      C(1,3,1,2), C(3,3,1,2), C(1,4,0,2), C(3,4,0,2), C(1,4,1,2), C(3,4,1,2), C(1,5,0,2), C(3,5,0,2),
      C(1,5,1,2), C(3,5,1,2), C(1,6,0,2), C(3,6,0,2), C(5,1,0,2), C(5,2,0,2), C(5,3,0,2), C(4,1,1,2)},
     {C(6,0,0,2), C(6,0,1,2), C(7,0,0,2), C(7,0,1,2), C(7,1,0,2), C(7,1,1,2), C(7,2,0,2), C(7,2,1,2),
      C(7,3,0,2), C(7,3,1,2), C(7,4,0,2), C(7,4,1,2), C(7,5,0,2), C(7,5,1,2), C(7,6,0,2), C(7,6,1,2),
      C(7,7,0,2), C(7,7,1,2), C(7,8,0,2), C(7,8,1,2), C(7,9,0,2), C(7,9,1,2), C(5,0,0,2), C(4,0,0,2),
      C(4,0,1,2), C(5,1,0,2), C(5,2,0,2), C(5,3,0,2), C(4,1,0,2), C(4,1,1,2), C(4,2,0,2), C(4,3,0,2)}},
};
#undef C
// number of ranked shapes each set needs (opaque, alpha)
__device__ __constant__ uint8_t kBc7Ranks[4][2] = {{2, 2}, {2, 2}, {3, 4}, {7, 10}};

#ifndef CFX_BC7_MIN_CTAS
#define CFX_BC7_MIN_CTAS 2      // 3 (80 registers, 236 B of spills) was measured: 3.08 vs 3.16 GTexel/s
#endif
template <int G>
__global__ void __launch_bounds__(kThreads, CFX_BC7_MIN_CTAS) bc7_kernel(const EncodeParams p)
{
    constexpr int kCandSet = G == 4 ? 0 : (G == 8 ? 1 : (G == 16 ? 2 : 3));
    constexpr int kPerWarp = 32/G;            // blocks a warp encodes at once
    constexpr int kShapes = 64/G;             // partition shapes scored per lane
    __shared__ __align__(16) uint32_t s_px[kTileBc7*kPxStride];
    __shared__ __align__(16) float4 s_pxf[kTileBc7*kPxfStride];
    __shared__ __align__(16) uint32_t s_out[kTileBc7*4];

    const uint32_t lane = lane_id();
    const uint32_t sub = lane % G, grp = lane / G;
    uint32_t chmask = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) if (p.color_mask & (1u << c)) chmask |= 0xFFu << (8*c);

    const uint32_t tiles = (p.total_blocks + kTileBc7 - 1)/kTileBc7;
    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint32_t first = tile*kTileBc7;
        const uint32_t n = min(static_cast<uint32_t>(kTileBc7), p.total_blocks - first);
        __syncthreads();
        stage_tile_u8(s_px, p, first, n, kPxStride);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n*16; i += blockDim.x) {
            uint32_t x = s_px[(i >> 4)*kPxStride + (i & 15)];
            s_pxf[(i >> 4)*kPxfStride + (i & 15)] = make_float4(static_cast<float>(x & 0xFF), static_cast<float>((x >> 8) & 0xFF),
                static_cast<float>((x >> 16) & 0xFF), static_cast<float>(x >> 24));
        }
        __syncthreads();

        for (uint32_t b0 = warp_id()*kPerWarp; b0 < n; b0 += kWarps*kPerWarp) {
            const uint32_t b = min(b0 + grp, n - 1);     // idle groups redo the last block, unstored
            const uint32_t* bx = s_px + b*kPxStride;
            const float4* bxf = s_pxf + b*kPxfStride;

            // block-level facts
            uint32_t amin = 255, amax = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) { amin = min(amin, bx[i] >> 24); amax = max(amax, bx[i] >> 24); }
            const bool has_alpha = amin < 255 && (p.color_mask & 8u);
            const bool flat_alpha = amin == amax;        // alpha adds nothing to any scatter matrix

            // ---- phase 1: score 64/G two-subset shapes per lane
            float sT[4] = {0, 0, 0, 0};
            float cT[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 2
            for (int i = 0; i < 16; ++i) {
                float4 x = bxf[i];
                sT[0] += x.x; sT[1] += x.y; sT[2] += x.z; sT[3] += x.w;
                cT[0] += x.x*x.x; cT[1] += x.x*x.y; cT[2] += x.x*x.z; cT[3] += x.x*x.w;
                cT[4] += x.y*x.y; cT[5] += x.y*x.z; cT[6] += x.y*x.w;
                cT[7] += x.z*x.z; cT[8] += x.z*x.w; cT[9] += x.w*x.w;
            }
            const float sT3[3] = {sT[0], sT[1], sT[2]};
            const float cT6[6] = {cT[0], cT[1], cT[2], cT[4], cT[5], cT[7]};
            uint32_t keys[kShapes];
#pragma unroll 1
            for (int j = 0; j < kShapes; ++j) {
                const uint32_t shape = sub*kShapes + j;
                const uint32_t key = flat_alpha ? score_shape_rgb(bxf, sT3, cT6, shape) : score_shape(bxf, sT, cT, shape);
#pragma unroll
                for (int jj = 0; jj < kShapes; ++jj) if (jj == j) keys[jj] = key;
            }

            // ---- phase 2+3: ranked shapes -> candidates
            // lanes 0,1: mode 6 (plain / extrapolating variant). Remaining lanes j = l-2:
            //   opaque: rank j/4, (mode 1, mode 3) x (plain, extrapolating) by j%4
            //   alpha : rank j/2, mode 7 x (plain, extrapolating)
            const uint32_t desc = kBc7Cand[kCandSet][has_alpha ? 1 : 0][sub];
            const uint32_t my_rank = cand_rank(desc);
            uint32_t my_shape = 0;
            // trip count must be warp-uniform (shuffles inside)
            const uint32_t need = __any_sync(0xFFFFFFFFu, has_alpha) ? kBc7Ranks[kCandSet][1] : kBc7Ranks[kCandSet][0];
            for (uint32_t r = 0; r < need; ++r) {
                uint32_t local = keys[0];
#pragma unroll
                for (int jj = 1; jj < kShapes; ++jj) local = min(local, keys[jj]);
                uint32_t win = group_min<G>(local);
#pragma unroll
                for (int jj = 0; jj < kShapes; ++jj) if (keys[jj] == win) keys[jj] = 0xFFFFFFFFu;
                if (my_rank == r) my_shape = win & 63u;
            }
            const uint32_t mode = cand_mode(desc);
            const uint32_t m1 = (mode == 6 || cand_is_dual(desc)) ? 0u : kBc7Part2[my_shape];

            Fit fit;
            const bool dual = cand_is_dual(desc);           // modes 4 / 5: their own fit (divergent lanes of the group)
            if (dual) fit_dual(bx, mode, cand_rotation(desc), cand_variant(desc) & 1u, cand_rounds(desc), chmask, fit);
            else fit_candidate(bxf, bx, mode, m1, cand_variant(desc), cand_rounds(desc), chmask, fit);

            // ---- phase 4: winner
            uint32_t total = fit.err[0] + fit.err[1];
            uint32_t key = (min(total, 0x03FFFFFFu) << 5) | sub;
            uint32_t win = group_min<G>(key);
            if (key == win && b0 + grp < n) {
                uint4 blk = dual ? pack_dual(mode, cand_rotation(desc), cand_variant(desc) & 1u, fit) : pack_block(mode, my_shape, m1, fit);
                *reinterpret_cast<uint4*>(s_out + b*4) = blk;
            }
        }
        __syncthreads();
        store_tile(p, s_out, first, n);
    }
}

int launch_bc7(const EncodeParams& p, cudaStream_t stream)
{
    uint32_t tiles = (p.total_blocks + kTileBc7 - 1)/kTileBc7;
    // lanes per block by quality: more lanes = more (mode, shape) candidates refined per block
    const void* k;
    switch (p.quality) {
        case 0: case 1: case 2: k = reinterpret_cast<const void*>(&bc7_kernel<4>); break;
        case 3: k = reinterpret_cast<const void*>(&bc7_kernel<8>); break;
        default: k = reinterpret_cast<const void*>(&bc7_kernel<16>); break;
    }
    uint32_t grid = min(tiles, persistent_ctas(k, kThreads));
    void* args[] = {const_cast<EncodeParams*>(&p)};
    if (cudaLaunchKernel(k, dim3(grid), dim3(kThreads), args, 0, stream) != cudaSuccess) return -4;
    return 1;
}

} // namespace cfx
