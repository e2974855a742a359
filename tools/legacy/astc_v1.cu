// ASTC LDR encoder kernel for sm_100a: one warp owns one NxM block and runs the whole search of
// astc_core.cuh with LANE = CANDIDATE:
//   load     lanes = texels: coalesced reads of the block's rows, clamp-to-edge gather, colour mask
//   init     lane 0 fits the single-subset line; lanes 1/2 cluster the texels into 2/3 groups;
//            lanes 4..7 set up the dual-plane hypotheses (one channel on its own weight plane)
//   rank     every lane scans 32 of the 1024 partition seeds against the clustering (popcount of
//            mask XOR), keeps its best, and scores it exactly (per-subset line-fit residual)
//   slots    lanes 0..3 turn the two best two-subset and three-subset seeds into slots
//   search   for every slot, every lane evaluates one block mode (weight grid x quantisation level)
//            per round, exactly: decimate ideal weights onto the grid (one least-squares step),
//            quantise, infill, quantise end points at the colour level the left-over bits allow,
//            least-squares end points, decoded squared error; each lane keeps its best candidate
//   refine   each lane alternates end-point solves and weight re-derivation on its best candidate
//   pack     warp argmin; the winning lane BISE-packs the 128-bit block
// Tables (block modes, infill/decimation, quantisation, partitions) are built on the host from the
// ASTC specification (astc_tables.hpp) once per footprint and live in global memory (L1/L2 resident).
//
// Replaces AstcConverter::process (lib/src/AstcConverter.cpp:208-230) for LDR profiles.
#include "astc3_tables.hpp"
#include "astc_core.cuh"
#include "common.cuh"
#include "kernels.h"

#include <cstdlib>
#include <map>
#include <mutex>

namespace cfx {

using namespace astc;

namespace {

constexpr int kAstcWarps = 4;
constexpr uint32_t kUScr = 64*32;        // bytes of grid-weight scratch per warp
constexpr uint32_t kWScr = 128*32;       // bytes of texel-weight scratch per warp (two planes)
constexpr uint32_t kWarpBytes = (sizeof(BlockState) + 15)/16*16 + kUScr + kWScr;

struct DeviceTables {
    Ctx ctx;
    Astc3Tab t3;
};

std::mutex g_mutex;
std::map<std::pair<int, int>, DeviceTables> g_tables;   // per device the library is bound to one GPU at a time

} // namespace

__global__ void __launch_bounds__(kAstcWarps*32) astc_kernel(const EncodeParams p, const Ctx ctx, const Plan plan)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = lane_id(), warp = warp_id();
    uint8_t* base = smem + warp*kWarpBytes;
    BlockState& st = *reinterpret_cast<BlockState*>(base);
    uint8_t* u_scr = base + (sizeof(BlockState) + 15)/16*16;
    uint8_t* w_scr = u_scr + kUScr;
    const uint32_t T = ctx.tab.texels, bw = ctx.tab.bw, bh = ctx.tab.bh;
    const bool alpha_off = p.alpha_type == 0;      // Alpha::None -> swizzle a = 1 (AstcConverter.cpp:143-146)

    for (uint32_t blk = blockIdx.x*kAstcWarps + warp; blk < p.total_blocks; blk += gridDim.x*kAstcWarps) {
        const uint32_t by = blk / p.blocks_x, bx = blk - by*p.blocks_x;
        __syncwarp();
        // ---- load
        bool differs = false, alpha = false;
        for (uint32_t i = lane; i < T; i += 32) {
            const uint32_t ty = i / bw, tx = i - ty*bw;
            const uint32_t x = min(bx*bw + tx, p.width - 1), y = min(by*bh + ty, p.height - 1);
            float4 v;
            if (p.src_format == SRC_RGBA8) {
                const uint32_t px = __ldg(reinterpret_cast<const uint32_t*>(p.src + static_cast<uint64_t>(y)*p.pitch) + x);
                v = make_float4(static_cast<float>(px & 0xFF), static_cast<float>((px >> 8) & 0xFF),
                    static_cast<float>((px >> 16) & 0xFF), static_cast<float>(px >> 24));
            } else {
                const float4 f = load_texel_f32(p, x, y);
                v = make_float4(fminf(fmaxf(f.x, 0.0f), 1.0f)*255.0f, fminf(fmaxf(f.y, 0.0f), 1.0f)*255.0f,
                    fminf(fmaxf(f.z, 0.0f), 1.0f)*255.0f, fminf(fmaxf(f.w, 0.0f), 1.0f)*255.0f);
            }
            if (!(p.color_mask & 1u)) v.x = 0.0f;
            if (!(p.color_mask & 2u)) v.y = 0.0f;
            if (!(p.color_mask & 4u)) v.z = 0.0f;
            if (!(p.color_mask & 8u)) v.w = 0.0f; else if (alpha_off) v.w = 255.0f;
            st.cf[i] = v;
            alpha |= v.w != 255.0f;
        }
        __syncwarp();
        const float4 first = st.cf[0];
        for (uint32_t i = lane; i < T; i += 32) {
            const float4 v = st.cf[i];
            differs |= v.x != first.x || v.y != first.y || v.z != first.z || v.w != first.w;
        }
        const bool constant = !__any_sync(0xFFFFFFFFu, differs);
        const bool has_alpha = __any_sync(0xFFFFFFFFu, alpha);
        uint4* dst = reinterpret_cast<uint4*>(p.dst) + blk;
        if (constant) {
            if (lane == 0) *dst = pack_void_extent(first);
            continue;
        }
        if (lane == 0) st.has_alpha = has_alpha ? 1u : 0u;
        if (lane < kSlots) st.slots[lane].valid = 0;
        __syncwarp();
        step_init(ctx, st, lane);
        __syncwarp();
        if (plan.slots > 1) {
            step_rank(ctx, st, lane);
            __syncwarp();
            step_score(ctx, st, lane);
            __syncwarp();
            step_slots(ctx, st, lane);
            __syncwarp();
        }
        // ---- search
        float best_err = 3.0e38f;
        uint32_t best_mode = 0, best_slot = 0;
        for (uint32_t s = 0; s < plan.slots; ++s) {
            if (!st.slots[s].valid) continue;
            const uint32_t type = slot_type(s);
            const uint32_t n = plan.n_cand[type];
            for (uint32_t c0 = 0; c0 < n; c0 += 32) {
                if (c0 + lane < n) {
                    const uint32_t mi = tab_u16(ctx, ctx.tab.off_cand[type] + (c0 + lane)*2u);
                    Enc e;
                    evaluate(ctx, st.cf, st.slots[s], tab_mode(ctx, mi), has_alpha, u_scr, w_scr, lane, -1, e);
                    if (e.err < best_err) { best_err = e.err; best_mode = mi; best_slot = s; }
                }
                __syncwarp();
            }
        }
        // ---- refine each lane's best, pick the winner, pack
        Enc enc;
        enc.err = 3.0e38f;
        const ModeInfo bm = tab_mode(ctx, best_mode);
        if (best_err < 3.0e38f)
            evaluate(ctx, st.cf, st.slots[best_slot], bm, has_alpha, u_scr, w_scr, lane, static_cast<int>(plan.refine), enc);
        uint32_t key = (__float_as_uint(fmaxf(enc.err, 0.0f)) & ~31u) | lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xFFFFFFFFu, key, o));
        if ((key & 31u) == lane)
            *dst = pack_block(ctx, st.slots[best_slot], bm, enc, has_alpha, u_scr, lane);
    }
}

int launch_astc2(const EncodeParams& p, const Ctx& ctx, cudaStream_t stream);   // astc2.cu: warp-cooperative search
int launch_astc3(const EncodeParams& p, const Ctx& ctx, const Astc3Tab& t3, cudaStream_t stream);   // astc3.cu: two-phase search

int launch_astc(const EncodeParams& p, cudaStream_t stream)
{
    int device = 0;
    cudaGetDevice(&device);
    Ctx ctx;
    Astc3Tab t3;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        auto key = std::make_pair(device, static_cast<int>(p.block_w*16 + p.block_h));
        auto it = g_tables.find(key);
        if (it == g_tables.end()) {
            Built b = build_tables(static_cast<int>(p.block_w), static_cast<int>(p.block_h));
            const Astc3Tab b3 = build_tables3(b);
            if (b.tab.n_grids > static_cast<uint32_t>(kMaxGrids3)) return -2;
            uint8_t* d = nullptr;
            if (cudaMalloc(&d, b.blob.size()) != cudaSuccess) return -4;
            if (cudaMemcpy(d, b.blob.data(), b.blob.size(), cudaMemcpyHostToDevice) != cudaSuccess) return -4;
            DeviceTables dt;
            dt.ctx.blob = d; dt.ctx.tab = b.tab; dt.t3 = b3;
            it = g_tables.insert(std::make_pair(key, dt)).first;
        }
        ctx = it->second.ctx;
        t3 = it->second.t3;
    }
    // Default: the two-phase tensor-core kernel of astc3.cu.  The earlier implementations are kept as
    // cross-checks: CFX_ASTC_V=2 (exhaustive warp-cooperative search, astc2.cu), CFX_ASTC_V=1 (lane per
    // candidate, below).
    static const int version = getenv("CFX_ASTC_V") ? atoi(getenv("CFX_ASTC_V")) : (getenv("CFX_ASTC_V1") ? 1 : 3);
    if (version >= 3) return launch_astc3(p, ctx, t3, stream);
    if (p.block_w*p.block_h > static_cast<uint32_t>(kMaxTexels) || p.type != 0u) return -2;     // the older kernels: LDR, <= 64 texels
    if (version == 2) return launch_astc2(p, ctx, stream);
    const Plan plan = make_plan(p.quality, ctx.tab);
    const size_t smem = kAstcWarps*kWarpBytes;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(astc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
            return -4;
        attr_set = true;
    }
    const uint32_t ctas_needed = (p.total_blocks + kAstcWarps - 1)/kAstcWarps;
    const uint32_t grid = min(ctas_needed, persistent_ctas(reinterpret_cast<const void*>(&astc_kernel), kAstcWarps*32, smem));
    astc_kernel<<<grid, kAstcWarps*32, smem, stream>>>(p, ctx, plan);
    return 1;
}

} // namespace cfx
