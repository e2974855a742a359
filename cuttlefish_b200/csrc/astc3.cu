// ASTC LDR encoder, two-phase warp-cooperative search (v3).  One warp owns one block.
//
//   setup    partition hypotheses ("slots": one subset; 2 x two subsets; 2 x three subsets; 4 x dual plane), each
//            with per-subset principal lines and per-texel ideal weights t in [0,1]
//   phase 1  EVERY (slot, block mode) of the footprint gets a model-based error estimate
//              e_line(slot) + sum_planes [ D(slot plane, grid) + S(slot plane, grid)*qvar(level) ] + cq(colour level)
//            D = | L (I - P_g M_g) t |^2 is what decimation to weight grid g loses.  It is computed for all 13 slot
//            planes x all grids at once on the TENSOR CORES: A = t^T (16 slot planes x texels, fp16, per warp),
//            B = R_g^T (texels x texels, fp16, per footprint table in mma fragment order), mma.sync.m16n8k16.
//   phase 2  the N best estimates (N by quality) are evaluated exactly: least-squares decimation g = M_g t (again one
//            small tensor-core GEMM), quantise, infill, least-squares end points from integer moment sums
//            (redux.sync), exact decoded error; the winner is refined (re-project, re-decimate, re-solve) and packed.
//
// Replaces AstcConverter::process -> astcenc compress_block (lib/src/AstcConverter.cpp:208-230,
// lib/astc-encoder/Source/astcenc_compress_symbolic.cpp:1163-1454); the estimate plays the role of astcenc's
// compute_ideal_endpoint_formats ranking (astcenc_pick_best_endpoint_format.cpp:1090), the exact pass that of its
// candidate refinement loop.  Our own search: PSNR parity with the reference, not byte parity.
#include "astc3_tables.hpp"
#include "astc_core.cuh"
#include "astc_hdr.cuh"
#include "common.cuh"
#include "kernels.h"

#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

namespace cfx {

using namespace astc;

namespace {

// Launch shape: W warps per CTA, CTAS resident CTAs per SM (register budget = 64K/(W*32*CTAS)); chosen per footprint in
// launch_one().  The kernel is ~15k SASS instructions, far beyond the 32 KB instruction cache, so the warps of a CTA are
// kept in the same phase of the search by CTA-wide barriers at the phase boundaries: one fetched instruction line then
// serves all of them.
// Finer-grained barriers were measured too (6x6): one per trip of the phase-2 loop is +10 % with two CTAs per SM but
// -8 % with the single CTA the kernel now runs as (the trips differ in length: the sum of the slowest warps' trips
// outweighs what the tighter lockstep saves), one per slot in the setup 7 / phase 1b' / 1c loops another -3 %.  Off.
#ifdef CFX_ASTC3_TUNE
#define CFX_ASTC3_TUNE_KEEP_ALL && !(tb.flags & 256u)      // developer: the candidate dump wants every slot's estimates
#else
#define CFX_ASTC3_TUNE_KEEP_ALL
#endif
// TMA-staged phase-1a operands (astc3_kernel<..., STAGE = true>: ONE thread fetches the R / M fragments of a weight grid
// into shared memory with cp.async.bulk, double buffered on mbarriers, every warp's MMAs read them from there).  Measured
// on the single-CTA launch shape (6x6, 4092^2): bit-identical output, 430.6 against 439.9 MTexel/s with every warp pulling
// its own fragments through L1 (4x4 342 / 352, 8x6 445 / 462) -- the fragments hit in L1 (95 %) once one CTA owns the SM,
// and the barrier per grid that hands the buffers back costs more than the long-scoreboard stalls it removes.  Off.
#ifndef CFX_ASTC3_STAGE
#define CFX_ASTC3_STAGE 0
#endif
#ifndef CFX_ASTC3_LOOPSYNC
#define CFX_ASTC3_LOOPSYNC 0
#endif
#ifndef CFX_ASTC3_SLOTSYNC
#define CFX_ASTC3_SLOTSYNC 0
#endif
#define PHASE_SYNC() do { if (LOCK) __syncthreads(); else __syncwarp(); } while (0)
constexpr int FX = 8;                          // texel fixed point scale
constexpr int kMaxCand = 16;
constexpr int kHdrAlphaShift = 3;              // HDR blocks keep alpha at 1/8 of the LDR scale (error weight 1/64): about the
constexpr float kHdrAlphaScale = 1.0f/static_cast<float>(1 << kHdrAlphaShift);     // weight the reference's HDR-alpha metric gives it

struct Slot3 {
    float4 e0[3], e1[3];       // line end points (0..255 per channel), sum(e1.rgb) >= sum(e0.rgb)
    float len2[3];             // squared length of each subset's line (first plane)
    float len2b;               // squared length of the second plane's line (dual plane slots)
    float e_line;              // squared distance of the texels from their lines: error floor of the slot
    uint32_t pc, seed, valid;
    int32_t dual_ch;
};
// Slots 0..8 as in astc_core.cuh (RGB(A) end points); 9: one subset with LUMINANCE end points (CEM 0); 10, 11 / 12, 13:
// the two-subset / three-subset partitionings of slots 1, 2 / 3, 4 with luminance end points.  Luminance slots are
// for opaque blocks only, which is why 12, 13 can live in the A operand rows of the alpha dual-plane slot (8, 12).
// 14 / 15, 16 / 17, 18: slot 0 / the two-subset slots 1, 2 / the three-subset slots 3, 4 with RGB BASE + SCALE end
// points on every subset (CEM 6: e1 = (r, g, b), e0 = e1*s/256 -- a line through black, which is what shaded surfaces of
// one material are; 4 / 8 / 12 stored values instead of 6 / 12 / 18 buy a finer weight grid or finer weights, and are
// what makes multi-subset encodings affordable at all on the larger footprints: astcenc picks CEM 6 for 20-50 % of the
// blocks of photographic content and on at least one subset of most of its two-subset blocks). In blocks with alpha every subset keeps a direct alpha pair (CEM 10), which is also what makes three subsets possible there (18 values).
// They are VIRTUAL: partition, ideal weights (A operand row), line lengths and the measured quantisation loss are those
// of the RGB sibling slot -- where a subset suits a line through black its free line nearly is one -- and only the
// error floor (distance from the through-black lines, Warp3T::scale_eline), the colour level class and the end point
// solve differ.
constexpr int kSlots3 = kSlots + 5;            // real slots (own Slot3 entry)
constexpr int kScaleSlot = kSlots3;            // first virtual slot
// 19, 20: the two-subset partitionings of slots 1, 2 with MIXED end point modes -- opaque blocks: the subset that loses
// least by it goes through black (CEM 6), the other keeps RGB (CEM 8), 10 stored values; blocks with alpha: a subset
// whose alpha is 255 throughout stores RGB only (CEM 8), the other RGBA (CEM 12), 14 values instead of 16 (valid when
// exactly one subset is opaque). Which subset is the cheap one: Warp3T::mix_mask.
constexpr int kMixSlot = kSlots3 + 5;
constexpr int kSlotsAll = kSlots3 + 7;         // + the virtual slots 14..20
constexpr int kLumSlot = 9, kLumRow = 13;
__device__ __forceinline__ bool slot_is_scale(uint32_t s) { return s >= static_cast<uint32_t>(kScaleSlot); }
__device__ __forceinline__ bool slot_is_lum(uint32_t s) { return s >= static_cast<uint32_t>(kLumSlot) && s < static_cast<uint32_t>(kScaleSlot); }
__device__ __forceinline__ uint32_t slot_base(uint32_t s) { return s < static_cast<uint32_t>(kScaleSlot) ? s : (s < static_cast<uint32_t>(kMixSlot) ? s - static_cast<uint32_t>(kScaleSlot) : s - 18u); }   // the slot whose Slot3 entry describes s
__device__ __forceinline__ uint32_t slot_kind(uint32_t s, bool has_alpha) { return s < 9u ? slot_type(s) : (s == 9u ? 4u : (s < 12u ? 5u : (s < 14u ? 6u : (s == 14u ? 7u : (s < 17u ? 8u : (s < 19u ? 9u : (has_alpha ? 11u : 10u))))))); }   // est list / colour level class
__device__ __forceinline__ uint32_t slot_row(uint32_t s) { return s < 9u ? s : (s < 12u ? s + 4u : (s == 12u ? 8u : (s == 13u ? 12u : slot_base(s)))); }        // A operand row of its first plane
__device__ __forceinline__ uint32_t slot_part(uint32_t s) { return s >= 19u ? s - 19u : (s >= 15u ? s - 15u : (s >= 10u ? s - 10u : (s - 1u) & 3u)); }          // index into Warp3T::part
constexpr float kMismatchWeight = 0.05f; // partition ranking: cost of one texel off the clustering, in mean squared spreads
constexpr int kDataLevels = 6;          // weight levels 2,3,4,5,6,8: their quantisation loss is MEASURED on the slot's ideal
                                        // weights (text and edges are bimodal: the uniform model is far off there)

// Moments of a set of texels about the block centre (FX units): count, sums, upper triangle of products.
struct Mom { float n, s[4], p[10]; };
struct LineFit { float m[4], v[4], resid, pad[3]; };     // mean (about the centre), unit direction, trace - lambda

// Per-warp working set, sized by the footprint: TP = texel capacity (NT*8), GP = weight grid capacity.
constexpr int grid_capacity(int NT) { return NT <= 3 ? 12 : (NT == 4 ? 20 : (NT == 5 ? 28 : (NT <= 7 ? 36 : (NT == 8 ? 52 : kMaxGrids3)))); }
constexpr int ta_stride(int TP) { return ((TP + 8)/2) % 8 == 0 ? TP + 16 : TP + 8; }   // halfs; conflict-free A fragment loads

template <int NT>
struct Warp3T {
    static constexpr int TP = NT*8;
    static constexpr int GP = grid_capacity(NT);
    static constexpr int TS = ta_stride(TP);
    int4 v[TP];                             // texels, FX fixed point
    Slot3 slots[kSlots3];
    float scale_eline[7];                   // error floor of the virtual slots 14..20
    uint32_t scale_valid[7];
    uint32_t mix_mask[2];                   // slots 19, 20: bit p = subset p takes the cheaper end point mode
    uint32_t mix_kind[2];                   // ... in a block with alpha: 0 = the cheap subset is opaque (CEM 8 + 12), 1 = it goes
                                            // through black with its own alpha pair (CEM 10 + 12)
    int sc[3], best_sc[3];                  // base + scale subsets: the quantised scale
    uint32_t contr, best_contr;             // bit p: subset p's direct end points are stored blue-contracted (values in epv)
    uint8_t part[4][TP];                    // subset of every texel for slots 1..4
    __half ta[kRows3][TS];                  // A operand: ideal weights per slot plane (rows 0..8 first planes,
                                            // 9..12 second planes of slots 5..8, 13..15 luminance slots 9..11;
                                            // refinement re-uses rows 0/1 once the candidates are done)
    struct Est {
        float qn[kSlots3][kDataLevels];     // measured quantisation loss of the slot's ideal weights at the coarse levels
        float D[16][GP];                    // decimation loss per slot plane and grid
        float Sm[8][GP];                    // sum_i len2_i kappa_gi for the multi-subset slots 1..4 and 10..13
    };
    struct Setup {
        LineFit lines[15];                  // slot 0, dual-plane slots 5..8, then the 10 subsets of slots 1..4
        Mom moms[11];                       // moments of those 10 subsets; [10] = the whole block
    };
    union { Est est; Setup setup; } u;      // the setup scratch is dead before phase 1 writes D
    float g[2][TP];                         // decimated ideal grid weights of the current candidate, per plane
    int ep[24];                             // quantised end points of the candidate: [subset][e0 rgba, e1 rgba]
    int best_ep[24];                        // (HDR: 12-bit values as end point mode 11 decodes them)
    float epf[24];                          // HDR: least-squares end points before mode-11 packing, 8-bit-like units
    int epv[24], best_epv[24];              // HDR: per subset the six packed mode-11 values + two LDR alpha end points;
                                            // LDR: the stored values v0..v7 of a blue-contracted subset
    uint32_t hkey[28];                      // HDR: (error, sub-mode) keys of the 9 packing candidates per subset
    uint8_t su[2*TP];                       // candidate grid weights (unquantised values 0..64), bit-stream order
    uint8_t sk[2*TP];                       // ... and their rank in the quantisation level's value table
    uint8_t best_su[2*TP];
    uint8_t best_sk[2*TP];
};

// model constants (fitted on the host with tools/emu_astc3.py)
constexpr float kLine = 1.1f, kDec = 1.0f, kQuant = 0.9f, kColor = 0.5f;

struct Tab3 { Ctx ctx; Astc3Tab t3; uint32_t flags; uint32_t hdr; float mis_w; int dbg_slot, dbg_level, dbg_nw; uint32_t cw; float4 sw, isw; };   // sw / isw: square roots of the channel weights and their reciprocals (kernel parameters: constant-bank operands, no registers)   // cw: channel error weights r, g, b, a in sixteenths, one byte each (0x10101010 = plain squared error)   // dbg_*: developer builds only (-1 = any)   // hdr: Texture::Type::UFloat (texels searched as LNS, end point mode 11)   // flags (developer): 1 = no luminance slot, 2 = model-only quantisation term

__device__ __forceinline__ int redux_add(int v) { return __reduce_add_sync(0xFFFFFFFFu, v); }
__device__ __forceinline__ uint32_t redux_addu(uint32_t v) { return __reduce_add_sync(0xFFFFFFFFu, v); }

// ---- TMA bulk staging of the phase-1a operands (STAGE variant of the kernel) ------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a copy that never lands must surface as an error, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
// One contiguous piece of the table blob -> shared memory (cp.async.bulk, the 1-D TMA copy); 16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint2 b)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

template <int KS, typename WS>
__device__ __forceinline__ void load_a(const WS& ws, uint32_t (&a)[KS][4], uint32_t lane)
{
    const uint32_t gq = lane >> 2, tq = lane & 3u;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        a[ks][0] = *reinterpret_cast<const uint32_t*>(&ws.ta[gq][ks*16 + 2*tq]);
        a[ks][1] = *reinterpret_cast<const uint32_t*>(&ws.ta[gq + 8][ks*16 + 2*tq]);
        a[ks][2] = *reinterpret_cast<const uint32_t*>(&ws.ta[gq][ks*16 + 2*tq + 8]);
        a[ks][3] = *reinterpret_cast<const uint32_t*>(&ws.ta[gq + 8][ks*16 + 2*tq + 8]);
    }
}

// g[plane][j] = clamp((M_grid t_row)[j], 0, 1) for the rows (slot planes) row0 / row1 (row1 < 0: single plane).
template <int KS, typename WS>
__device__ __forceinline__ void decimate_mma(const Tab3& tb, WS& ws, const uint32_t (&a)[KS][4], uint32_t grid, uint32_t nw,
    int row0, int row1, uint32_t lane)
{
    const uint32_t gq = lane >> 2, tq = lane & 3u;
    const uint32_t off = __ldg(reinterpret_cast<const uint32_t*>(tb.ctx.blob + tb.t3.off_mfrag_idx) + grid);
    const uint2* frag = reinterpret_cast<const uint2*>(tb.ctx.blob + off) + lane;
    const uint32_t ntw = (nw + 7u) >> 3;
    for (uint32_t nt = 0; nt < ntw; ++nt) {
        float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) mma16816(c, a[ks], __ldg(frag + (nt*KS + ks)*32u));
        const uint32_t j = nt*8u + 2u*tq;
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
            const int row = pl ? row1 : row0;
            if (row < 0) continue;
            if (static_cast<int>(gq) == (row & 7)) {
                const float x = row < 8 ? c[0] : c[2], y = row < 8 ? c[1] : c[3];
                ws.g[pl][j] = fminf(fmaxf(x, 0.0f), 1.0f);
                ws.g[pl][j + 1] = fminf(fmaxf(y, 0.0f), 1.0f);
            }
        }
    }
    __syncwarp();
}

// Multi-subset slots: M_g is the UNWEIGHTED least-squares decimation, but a grid weight shared by texels of two
// subsets should follow the subset whose line is long -- a texel of a nearly constant subset does not care what its
// weight is (astcenc weights its ideal-weight decimation the same way, compute_ideal_weights_for_decimation,
// lib/astc-encoder/Source/astcenc_ideal_endpoints_and_weights.cpp:845).  Two damped Jacobi steps of the weighted
// problem, starting from M_g t: residual per texel through the infill table (lane = texel), update per grid weight
// through the transposed table (lane = grid weight).  Single plane only (multi-subset slots have no second plane).
template <typename WS>
__device__ __forceinline__ void reweight_decimation(const Ctx& c, WS& ws, uint32_t s, uint32_t grid, uint32_t nw, uint32_t row, uint32_t lane)
{
    const uint32_t T = c.tab.texels;
    const Slot3& slot = ws.slots[slot_base(s)];
    const uint8_t* parts = ws.part[slot_part(s)];
    const float wmax = fmaxf(fmaxf(slot.len2[0], slot.len2[1]), slot.pc > 2 ? slot.len2[2] : 0.0f);
    if (!(wmax > 0.0f)) return;
    const float iw = 1.0f/wmax;
    const uint32_t inf_off = c.tab.off_infill + grid*T*8u;
    const uint32_t start_off = c.tab.off_csr_start + grid*(kMaxTexels + 2)*2u;
    const uint32_t ent_off = c.tab.off_csr_ent + grid*4u*T*2u;
    float* res = ws.g[1];                                  // scratch: weighted residual per texel
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
        for (uint32_t i = lane; i < T; i += 32) {
            const uint2 inf = tab_u32x2(c, inf_off + i*8u);
            float r = 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) r += static_cast<float>((inf.y >> (8*k)) & 0xFFu)*ws.g[0][(inf.x >> (8*k)) & 0xFFu];
            res[i] = (__half2float(ws.ta[row][i]) - r*(1.0f/16.0f))*(slot.len2[parts[i]]*iw + 1e-3f);
        }
        __syncwarp();
        for (uint32_t j = lane; j < nw; j += 32) {
            const uint32_t e0 = tab_u16(c, start_off + j*2u), e1 = tab_u16(c, start_off + (j + 1u)*2u);
            float num = 0.0f, den = 0.0f;
            for (uint32_t e = e0; e < e1; ++e) {
                const uint32_t ent = tab_u16(c, ent_off + e*2u);
                const uint32_t i = ent & 0xFFu;
                const float f = static_cast<float>(ent >> 8);
                num += f*res[i]; den += f*f*(slot.len2[parts[i]]*iw + 1e-3f);
            }
            if (den > 0.0f) ws.g[0][j] = fminf(fmaxf(ws.g[0][j] + 12.0f*num/den, 0.0f), 1.0f);
        }
        __syncwarp();
    }
}

// Evaluate block mode m on slot s with the decimated weights in ws.g; on success ws.su / ws.ep hold the candidate and
// its exact decoded error (FX^2 units) is returned.
// quantise = false: ws.su / ws.sk already hold the candidate's quantised weights (realign_weights).
template <int K, bool hdr, typename WS>
__device__ __forceinline__ float evaluate3(const Ctx& c, WS& ws, uint32_t s, const ModeInfo& m, uint32_t cl, bool has_alpha,
    uint32_t lane, bool quantise = true, bool solve = true, uint32_t cw = 0x10101010u)
{
    const bool lum = slot_is_lum(s);
    // per-subset end point mode: 0 = direct (RGB / RGBA), 1 = base + scale (+ alpha pair), 2 = RGB only in a block with alpha
    const uint32_t mixm = s >= static_cast<uint32_t>(kMixSlot) ? ws.mix_mask[s - kMixSlot] : 0u;
    auto sub_mode = [&](uint32_t p) -> uint32_t {
        if (s < static_cast<uint32_t>(kScaleSlot)) return 0u;
        if (s < static_cast<uint32_t>(kMixSlot)) return 1u;
        return (mixm >> p) & 1u ? (has_alpha && ws.mix_kind[s - kMixSlot] == 0u ? 2u : 1u) : 0u;
    };
    const uint32_t T = c.tab.texels;
    const Slot3& slot = ws.slots[slot_base(s)];
    const uint32_t pc = slot.pc;
    const int dc = slot.dual_ch;
    const uint32_t planes = dc >= 0 ? 2u : 1u;
    const uint32_t L = m.level, nw = m.nw;
    const float nm1 = static_cast<float>(kWqN[L] - 1);
    // quantise (lane = grid weight)
    for (uint32_t pl = 0; quantise && pl < planes; ++pl)
        for (uint32_t j = lane; j < nw; j += 32) {
            // nearest level by VALUE: the trit and quint levels are not evenly spaced (24 levels: 0 2 5 8 11 13 ...)
            const float target = ws.g[pl][j]*64.0f;
            int k = min(max(__float2int_rn(ws.g[pl][j]*nm1), 0), static_cast<int>(kWqN[L]) - 1);
            int u = static_cast<int>(tab_u8(c, c.tab.off_wq_val + L*32u + static_cast<uint32_t>(k)));
            if (kWqN[L] > 2u) {
                const int kn = static_cast<float>(u) > target ? max(k - 1, 0) : min(k + 1, static_cast<int>(kWqN[L]) - 1);
                const int un = static_cast<int>(tab_u8(c, c.tab.off_wq_val + L*32u + static_cast<uint32_t>(kn)));
                if (fabsf(static_cast<float>(un) - target) < fabsf(static_cast<float>(u) - target)) { k = kn; u = un; }
            }
            ws.su[j*planes + pl] = static_cast<uint8_t>(u);
            ws.sk[j*planes + pl] = static_cast<uint8_t>(k);
        }
    __syncwarp();
    // infill (lane = texel)
    int w[K][2];
    uint32_t part[K];
    bool live[K];
    const uint32_t inf_off = c.tab.off_infill + static_cast<uint32_t>(m.grid)*T*8u;
    const uint8_t* parts = ws.part[slot_part(s)];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t i = lane + 32u*k;
        live[k] = i < T;
        const uint32_t ii = live[k] ? i : 0u;
        part[k] = pc > 1 ? parts[ii] : 0u;
        const uint2 inf = tab_u32x2(c, inf_off + ii*8u);
#pragma unroll
        for (uint32_t pl = 0; pl < 2; ++pl) {
            if (pl >= planes) { w[k][pl] = w[k][0]; continue; }
            uint32_t acc = 8;
#pragma unroll
            for (int q = 0; q < 4; ++q) acc += ((inf.y >> (8*q)) & 0xFFu)*ws.su[((inf.x >> (8*q)) & 0xFFu)*planes + pl];
            w[k][pl] = static_cast<int>(acc >> 4);
        }
    }
    // least-squares end points per subset from integer moment sums (solve == false: ws.ep / ws.epv / ws.sc / ws.contr are
    // kept as the caller left them and only the error is measured)
    for (uint32_t p = 0; solve && p < pc; ++p) {
        int A = 0, B = 0, C = 0, P0 = 0, P1 = 0, P2 = 0, P3 = 0, Q0 = 0, Q1 = 0, Q2 = 0, Q3 = 0;
        int A2 = 0, B2 = 0, C2 = 0, PD = 0, QD = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (!live[k] || part[k] != p) continue;
            const int4 x = ws.v[lane + 32u*k];
            const int ww = w[k][0], iw = 64 - ww;
            A += iw*iw; B += iw*ww; C += ww*ww;
            P0 += iw*x.x; P1 += iw*x.y; P2 += iw*x.z; P3 += iw*x.w;
            Q0 += ww*x.x; Q1 += ww*x.y; Q2 += ww*x.z; Q3 += ww*x.w;
            if (dc >= 0) {
                const int w2 = w[k][1], i2 = 64 - w2;
                const int xd = dc == 0 ? x.x : (dc == 1 ? x.y : (dc == 2 ? x.z : x.w));
                A2 += i2*i2; B2 += i2*w2; C2 += w2*w2; PD += i2*xd; QD += w2*xd;
            }
        }
        A = redux_add(A); B = redux_add(B); C = redux_add(C);
        P0 = redux_add(P0); P1 = redux_add(P1); P2 = redux_add(P2);
        Q0 = redux_add(Q0); Q1 = redux_add(Q1); Q2 = redux_add(Q2);
        if (has_alpha) { P3 = redux_add(P3); Q3 = redux_add(Q3); }
        if (dc >= 0) { A2 = redux_add(A2); B2 = redux_add(B2); C2 = redux_add(C2); PD = redux_add(PD); QD = redux_add(QD); }
        // lane c (0..7) solves and quantises component c: e0 r,g,b,a then e1 r,g,b,a
        if (lane < 8) {
            const uint32_t ch = lane & 3u, which = lane >> 2;
            float fA = static_cast<float>(A), fB = static_cast<float>(B), fC = static_cast<float>(C);
            float fP = static_cast<float>(ch == 0 ? P0 : (ch == 1 ? P1 : (ch == 2 ? P2 : P3)));
            float fQ = static_cast<float>(ch == 0 ? Q0 : (ch == 1 ? Q1 : (ch == 2 ? Q2 : Q3)));
            if (dc >= 0 && static_cast<int>(ch) == dc) {
                fA = static_cast<float>(A2); fB = static_cast<float>(B2); fC = static_cast<float>(C2);
                fP = static_cast<float>(PD); fQ = static_cast<float>(QD);
            }
            const float det = fA*fC - fB*fB;
            float val;
            if (fabsf(det) < 1e-4f*(fA + fC)*(fA + fC) + 1e-6f) {
                const float4 e = which ? slot.e1[p] : slot.e0[p];
                val = ch == 0 ? e.x : (ch == 1 ? e.y : (ch == 2 ? e.z : e.w));
            } else {
                val = (which ? (fA*fQ - fB*fP) : (fC*fP - fB*fQ))*(64.0f/static_cast<float>(FX))/det;
            }
            if (lum) {
                // gray end points: the least-squares solution is the mean of the per-channel ones (same system matrix)
                // (blocks with alpha, CEM 4: the alpha pair keeps its own solution)
                const uint32_t b4 = which*4u;
                const float gray = (__shfl_sync(0xFFu, val, b4) + __shfl_sync(0xFFu, val, b4 + 1u) + __shfl_sync(0xFFu, val, b4 + 2u))*(1.0f/3.0f);
                if (ch < 3u) val = gray;
            }
            int q = 255;
            const uint32_t smode = sub_mode(p);
            if (smode == 1u) {
                // base + scale: s = where the unconstrained e0 falls along e1, snapped to the colour level; with the scale
                // fixed the least-squares base is closed form in the same moment sums: every texel is a_i*E with
                // a_i = s(1 - w_i) + w_i, so E = sum a_i x_i / sum a_i^2 = (s P + Q)/(s^2 A + 2 s B + C)
                const float e0r = __shfl_sync(0xFFu, val, 0), e0g = __shfl_sync(0xFFu, val, 1), e0b = __shfl_sync(0xFFu, val, 2);
                const float e1r = __shfl_sync(0xFFu, val, 4), e1g = __shfl_sync(0xFFu, val, 5), e1b = __shfl_sync(0xFFu, val, 6);
                const float dd = e1r*e1r + e1g*e1g + e1b*e1b;
                const float sp = dd > 0.0f ? (e0r*e1r + e0g*e1g + e0b*e1b)/dd : 1.0f;
                const int s8 = min(max(__float2int_rn(sp*256.0f), 0), 255);
                const int sq = static_cast<int>(tab_u8(c, c.tab.off_cq_val + cl*256u + tab_u8(c, c.tab.off_cq_near + cl*256u + static_cast<uint32_t>(s8))));
                const float sf = static_cast<float>(sq)*(1.0f/256.0f);
                const float den = sf*sf*fA + 2.0f*sf*fB + fC;
                const float E = den > 0.0f ? (sf*fP + fQ)*(64.0f/static_cast<float>(FX))/den : val;
                const int iv = min(max(__float2int_rn(E), 0), 255);
                const int qe = static_cast<int>(tab_u8(c, c.tab.off_cq_val + cl*256u + tab_u8(c, c.tab.off_cq_near + cl*256u + static_cast<uint32_t>(iv))));
                // e0 as the decoder derives it; alpha (CEM 10, blocks with alpha) keeps its own direct pair
                if (ch == 3) {
                    q = 255;
                    if (has_alpha) {
                        const int av = min(max(__float2int_rn(val), 0), 255);
                        q = static_cast<int>(tab_u8(c, c.tab.off_cq_val + cl*256u + tab_u8(c, c.tab.off_cq_near + cl*256u + static_cast<uint32_t>(av))));
                    }
                } else q = which ? qe : (qe*sq) >> 8;
                if (lane == 0) ws.sc[p] = sq;
            }
            else if (hdr) ws.epf[p*8u + lane] = val;
            else if (ch < 3 || (has_alpha && smode != 2u)) {
                const int iv = min(max(__float2int_rn(val), 0), 255);
                const uint32_t rank = tab_u8(c, c.tab.off_cq_near + cl*256u + static_cast<uint32_t>(iv));
                q = static_cast<int>(tab_u8(c, c.tab.off_cq_val + cl*256u + rank));
            }
            if (!hdr && !lum && smode != 1u) {
                // BLUE CONTRACTION (ASTC spec, LDR RGB(A) direct; astcenc try_quantize_rgb_blue_contract,
                // lib/astc-encoder/Source/astcenc_color_quantize.cpp): the end points may be stored as (2r - b, 2g - b, b)
                // in swapped order -- the decoder recognises it by the order of the two sums and takes r = (r' + b) >> 1 --
                // which halves the quantisation step of red and green wherever 2r - b and 2g - b stay inside 0..255,
                // i.e. on near-gray colours.  Taken when it fits, keeps the order strict and reproduces the least-squares
                // end points better than the direct values do.
                const float bval = __shfl_sync(0xFFu, val, which*4u + 2u);
                const float cval = ch < 2u ? 2.0f*val - bval : val;
                const bool fits = ch >= 2u || (cval >= -0.5f && cval <= 255.5f);
                bool use = false;
                if (__ballot_sync(0xFFu, !fits) == 0u) {          // (uniform over the eight lanes; colourful blocks stop here)
                    int cq = q;
                    if (ch < 3u) {
                        const int civ = min(max(__float2int_rn(cval), 0), 255);
                        cq = static_cast<int>(tab_u8(c, c.tab.off_cq_val + cl*256u + tab_u8(c, c.tab.off_cq_near + cl*256u + static_cast<uint32_t>(civ))));
                    }
                    const int cqb = __shfl_sync(0xFFu, cq, which*4u + 2u);
                    const int dec = ch < 2u ? (cq + cqb) >> 1 : cq;
                    float dd = 0.0f, dq = 0.0f;
                    int ssum = 0;
                    if (ch < 3u) {
                        dd = (static_cast<float>(dec) - val)*(static_cast<float>(dec) - val);
                        dq = (static_cast<float>(q) - val)*(static_cast<float>(q) - val);
                        ssum = cq;
                    }
                    ssum += __shfl_xor_sync(0xFFu, ssum, 1); ssum += __shfl_xor_sync(0xFFu, ssum, 2);      // per end point
                    const int osum = __shfl_xor_sync(0xFFu, ssum, 4);
                    // one reduction for both errors: their difference decides
                    float diff = dd - dq;
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) diff += __shfl_xor_sync(0xFFu, diff, o);
                    const bool order_ok = which ? ssum > osum : osum > ssum;   // stored sum of end point 1 strictly above end point 0's
                    use = order_ok && diff < 0.0f;
                    if (use) {
                        q = (ch < 3u || (has_alpha && smode != 2u)) ? dec : q;
                        ws.epv[p*8u + 2u*ch + (which ? 0u : 1u)] = cq;      // end point 1 first: v0 v2 v4 (v6), then end point 0
                    }
                }
                if (lane == 0) ws.contr = use ? (ws.contr | (1u << p)) : (ws.contr & ~(1u << p));
            } else if (lane == 0 && !hdr) ws.contr &= ~(1u << p);
            if (!hdr) ws.ep[p*8u + lane] = q;
        }
    }
    __syncwarp();
    if (hdr) {
        // end point mode 11: lane = (subset, sub-mode) packs the least-squares end points, snaps the six values to the
        // colour level, decodes them again; the sub-mode whose decoded end points are closest wins (astc_hdr.cuh)
        const uint32_t hp = lane/9u, hm = lane - hp*9u;
        int vq[6], e0[3], e1[3];
        float herr = 3.0e38f;
        if (hp < pc) {
            float t0[3], t1[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                t0[k] = fminf(fmaxf(ws.epf[hp*8u + k]*16.0f, 0.0f), 4095.0f);
                t1[k] = fminf(fmaxf(ws.epf[hp*8u + 4u + k]*16.0f, 0.0f), 4095.0f);
            }
            herr = hdr_rgb_try(c, cl, static_cast<int>(hm), t0, t1, vq, e0, e1);
            ws.hkey[lane] = (__float_as_uint(herr) & ~15u) | hm;
        }
        __syncwarp();
        if (hp < pc) {
            uint32_t bestk = 0xFFFFFFFFu;
#pragma unroll
            for (uint32_t k = 0; k < 9; ++k) bestk = min(bestk, ws.hkey[hp*9u + k]);
            if ((bestk & 15u) == hm) {
#pragma unroll
                for (int k = 0; k < 3; ++k) { ws.ep[hp*8u + k] = e0[k]; ws.ep[hp*8u + 4u + k] = e1[k]; }
                // alpha (end point mode 14: HDR RGB + LDR alpha) stays an 8-bit pair at the same colour level
                int a0 = 255, a1 = 255;
                if (has_alpha) {
                    a0 = static_cast<int>(quant_color(c, cl, ws.epf[hp*8u + 3u]*(1.0f/kHdrAlphaScale)));
                    a1 = static_cast<int>(quant_color(c, cl, ws.epf[hp*8u + 7u]*(1.0f/kHdrAlphaScale)));
                }
                ws.ep[hp*8u + 3u] = a0; ws.ep[hp*8u + 7u] = a1;
#pragma unroll
                for (int k = 0; k < 6; ++k) ws.epv[hp*8u + k] = vq[k];
                ws.epv[hp*8u + 6u] = a0; ws.epv[hp*8u + 7u] = a1;
            }
        }
        __syncwarp();
    }
    // keep sum(e1.rgb) >= sum(e0.rgb) (otherwise the decoder would blue-contract): swap the end points
    if (lane < pc && !lum && !hdr && sub_mode(lane) != 1u && !((ws.contr >> lane) & 1u)) {
        int* e = ws.ep + lane*8u;
        if (e[4] + e[5] + e[6] < e[0] + e[1] + e[2]) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { const int tmp = e[k]; e[k] = e[4 + k]; e[4 + k] = tmp; }
        }
    }
    __syncwarp();
    // exact decoded error (lane = texel)
    if (hdr) {
        // 12-bit end points against LNS texels (FX units are LNS/32 = half a 12-bit step); result in FX^2 units
        uint32_t err12 = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (!live[k]) continue;
            const int* e = ws.ep + part[k]*8u;
            const int4 x = ws.v[lane + 32u*k];
            const int w0 = w[k][0], w1 = w[k][1];
            int d = ((e[0]*(64 - (dc == 0 ? w1 : w0)) + e[4]*(dc == 0 ? w1 : w0) + 32) >> 6) - 2*x.x; err12 += static_cast<uint32_t>(d*d) >> 2;
            d = ((e[1]*(64 - (dc == 1 ? w1 : w0)) + e[5]*(dc == 1 ? w1 : w0) + 32) >> 6) - 2*x.y; err12 += static_cast<uint32_t>(d*d) >> 2;
            d = ((e[2]*(64 - (dc == 2 ? w1 : w0)) + e[6]*(dc == 2 ? w1 : w0) + 32) >> 6) - 2*x.z; err12 += static_cast<uint32_t>(d*d) >> 2;
            if (has_alpha) {
                // alpha end points are real 8-bit values, the texel's alpha is scaled by kHdrAlphaScale
                d = (((e[3]*(64 - (dc == 3 ? w1 : w0)) + e[7]*(dc == 3 ? w1 : w0))*FX + (32 << kHdrAlphaShift)) >> (6 + kHdrAlphaShift)) - x.w;
                err12 += static_cast<uint32_t>(d*d);
            }
        }
        return static_cast<float>(redux_addu(err12));
    }
    uint32_t err = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (!live[k]) continue;
        const int* e = ws.ep + part[k]*8u;
        const int4 x = ws.v[lane + 32u*k];
        const int w0 = w[k][0], w1 = w[k][1];
        // (channel weights in sixteenths: 16 = the plain squared error, bit for bit)
        int d = ((e[0]*FX*(64 - (dc == 0 ? w1 : w0)) + e[4]*FX*(dc == 0 ? w1 : w0) + 32) >> 6) - x.x; err += (static_cast<uint32_t>(d*d)*(cw & 0xFFu)) >> 4;
        d = ((e[1]*FX*(64 - (dc == 1 ? w1 : w0)) + e[5]*FX*(dc == 1 ? w1 : w0) + 32) >> 6) - x.y; err += (static_cast<uint32_t>(d*d)*((cw >> 8) & 0xFFu)) >> 4;
        d = ((e[2]*FX*(64 - (dc == 2 ? w1 : w0)) + e[6]*FX*(dc == 2 ? w1 : w0) + 32) >> 6) - x.z; err += (static_cast<uint32_t>(d*d)*((cw >> 16) & 0xFFu)) >> 4;
        if (has_alpha) {
            d = ((e[3]*FX*(64 - (dc == 3 ? w1 : w0)) + e[7]*FX*(dc == 3 ? w1 : w0) + 32) >> 6) - x.w; err += (static_cast<uint32_t>(d*d)*(cw >> 24)) >> 4;
        }
    }
    err = redux_addu(err);
    return static_cast<float>(err);
}

// One pass of DISCRETE weight realignment on the best candidate so far (the role of astcenc's realign_weights_decimated,
// lib/astc-encoder/Source/astcenc_compress_symbolic.cpp:188): every grid weight tries the quantisation level below
// and above its own and keeps whichever lowers the exact decoded error of the texels it feeds, end points held fixed.
// Rounding the least-squares grid weights one by one is far from the best joint choice at the coarse levels (2..6
// values) that real images mostly end up with; this descent is what closes most of that gap. Lane = grid weight.
// The weights a texel blends are a 2x2 patch of the grid, so weights of equal (x, y) parity never share a texel: the
// four parity classes are swept one after the other, each class in parallel. LDR only. Leaves the result in
// ws.su / ws.sk; the caller re-solves the end points and measures (evaluate3 with quantise = false).
template <typename WS>
__device__ __forceinline__ void realign_weights(const Ctx& c, WS& ws, uint32_t s, const ModeInfo& m, bool has_alpha, uint32_t lane,
    uint32_t cw = 0x10101010u)
{
    constexpr uint32_t TP = WS::TP;
    const uint32_t T = c.tab.texels;
    const Slot3& slot = ws.slots[slot_base(s)];
    const uint32_t pc = slot.pc;
    const int dc = slot.dual_ch;
    const uint32_t planes = dc >= 0 ? 2u : 1u;
    const uint32_t L = m.level, nw = m.nw, grid = m.grid;
    const int nq = static_cast<int>(kWqN[L]);
    const uint8_t* parts = ws.part[slot_part(s)];
    const uint32_t inf_off = c.tab.off_infill + grid*T*8u;
    const uint32_t start_off = c.tab.off_csr_start + grid*(kMaxTexels + 2)*2u;
    const uint32_t ent_off = c.tab.off_csr_ent + grid*4u*T*2u;
    const uint32_t gw = tab_u8(c, c.tab.off_grids + grid*4u);           // GridInfo::w
    int* S = reinterpret_cast<int*>(&ws.g[0][0]);                       // [plane][TP]: infill sums (16 x texel weight + 8)
    for (uint32_t j = lane; j < nw*planes; j += 32) { ws.su[j] = ws.best_su[j]; ws.sk[j] = ws.best_sk[j]; }
    __syncwarp();
    for (uint32_t i = lane; i < T; i += 32) {
        const uint2 inf = tab_u32x2(c, inf_off + i*8u);
        for (uint32_t pl = 0; pl < planes; ++pl) {
            uint32_t acc = 8;
#pragma unroll
            for (int q = 0; q < 4; ++q) acc += ((inf.y >> (8*q)) & 0xFFu)*ws.su[((inf.x >> (8*q)) & 0xFFu)*planes + pl];
            S[pl*TP + i] = static_cast<int>(acc);
        }
    }
    __syncwarp();
    const int nchan = has_alpha ? 4 : 3;
#pragma unroll 1
    for (uint32_t pl = 0; pl < planes; ++pl) {
#pragma unroll 1
        // (a full-resolution grid has one weight per texel: nothing is shared, one sweep does them all)
        const uint32_t ncol = nw == T ? 1u : 4u;
        for (uint32_t colour = 0; colour < ncol; ++colour) {
            for (uint32_t j = lane; j < nw; j += 32) {
                const uint32_t jy = j/gw, jx = j - jy*gw;
                if (ncol == 4u && ((jx & 1u) | ((jy & 1u) << 1)) != colour) continue;
                const int k = ws.sk[j*planes + pl];
                const int u = ws.su[j*planes + pl];
                const int du_dn = k > 0 ? static_cast<int>(tab_u8(c, c.tab.off_wq_val + L*32u + static_cast<uint32_t>(k - 1))) - u : 0;
                const int du_up = k + 1 < nq ? static_cast<int>(tab_u8(c, c.tab.off_wq_val + L*32u + static_cast<uint32_t>(k + 1))) - u : 0;
                const uint32_t e0 = tab_u16(c, start_off + j*2u), e1 = tab_u16(c, start_off + (j + 1u)*2u);
                int d_dn = 0, d_up = 0;
                for (uint32_t e = e0; e < e1; ++e) {
                    const uint32_t ent = tab_u16(c, ent_off + e*2u);
                    const uint32_t i = ent & 0xFFu;
                    const int f = static_cast<int>(ent >> 8);
                    const int* ep = ws.best_ep + (pc > 1 ? parts[i] : 0u)*8u;
                    const int4 x = ws.v[i];
                    const int xs[4] = {x.x, x.y, x.z, x.w};
                    const int sum = S[pl*TP + i];
                    const int w0 = sum >> 4, wd = (sum + f*du_dn) >> 4, wu = (sum + f*du_up) >> 4;
                    int c0 = 0, cd = 0, cu = 0;
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        if (ch >= nchan) continue;
                        // this plane's channels only: the dual channel rides the second plane
                        if (dc >= 0 && (ch == dc) != (pl == 1u)) continue;
                        const int a = ep[ch]*FX, b = ep[4 + ch]*FX, xv = xs[ch];
                        const int wc = static_cast<int>((cw >> (8*ch)) & 0xFFu);
                        int d = ((a*(64 - w0) + b*w0 + 32) >> 6) - xv; c0 += (d*d*wc) >> 4;
                        d = ((a*(64 - wd) + b*wd + 32) >> 6) - xv; cd += (d*d*wc) >> 4;
                        d = ((a*(64 - wu) + b*wu + 32) >> 6) - xv; cu += (d*d*wc) >> 4;
                    }
                    d_dn += cd - c0; d_up += cu - c0;
                }
                int du = 0, dk = 0;
                if (du_dn != 0 && d_dn < 0 && d_dn <= d_up) { du = du_dn; dk = -1; }
                else if (du_up != 0 && d_up < 0) { du = du_up; dk = 1; }
                if (dk != 0) {
                    ws.su[j*planes + pl] = static_cast<uint8_t>(u + du);
                    ws.sk[j*planes + pl] = static_cast<uint8_t>(k + dk);
                    for (uint32_t e = e0; e < e1; ++e) {
                        const uint32_t ent = tab_u16(c, ent_off + e*2u);
                        S[pl*TP + (ent & 0xFFu)] += static_cast<int>(ent >> 8)*du;
                    }
                }
            }
            __syncwarp();
        }
    }
}

template <typename WS>
__device__ __forceinline__ void keep_best3(WS& ws, uint32_t nw, uint32_t planes, uint32_t pc, uint32_t lane)
{
    for (uint32_t j = lane; j < nw*planes; j += 32) { ws.best_su[j] = ws.su[j]; ws.best_sk[j] = ws.sk[j]; }
    if (lane < pc*8u) ws.best_ep[lane] = ws.ep[lane];
    if (lane < pc*8u) ws.best_epv[lane] = ws.epv[lane];
    if (lane < 3u) ws.best_sc[lane] = ws.sc[lane];
    if (lane == 0u) ws.best_contr = ws.contr;
    __syncwarp();
}


// ---- warp-cooperative setup ----------------------------------------------------------------------
// Moments of a set of texels about the block centre `ctr` (FX units): count, sums, upper triangle of products.

__device__ __forceinline__ float warp_min_f(float f)
{
    int o = __float_as_int(f); o ^= (o >> 31) & 0x7FFFFFFF;
    o = __reduce_min_sync(0xFFFFFFFFu, o);
    o ^= (o >> 31) & 0x7FFFFFFF;
    return __int_as_float(o);
}
__device__ __forceinline__ float warp_max_f(float f)
{
    int o = __float_as_int(f); o ^= (o >> 31) & 0x7FFFFFFF;
    o = __reduce_max_sync(0xFFFFFFFFu, o);
    o ^= (o >> 31) & 0x7FFFFFFF;
    return __int_as_float(o);
}

// Principal line of a texel set from its moments; zero_ch (>= 0) is left out (it gets its own weight plane).
__device__ __forceinline__ float line_core(const Mom& mo, int zero_ch, int iters, float (&m)[4], float (&v)[4])
{
    const float inv = mo.n > 0.0f ? 1.0f/mo.n : 0.0f;
    m[0] = mo.s[0]*inv; m[1] = mo.s[1]*inv; m[2] = mo.s[2]*inv; m[3] = mo.s[3]*inv;
    float cv[10];
    cv[0] = mo.p[0] - mo.s[0]*m[0]; cv[1] = mo.p[1] - mo.s[0]*m[1]; cv[2] = mo.p[2] - mo.s[0]*m[2]; cv[3] = mo.p[3] - mo.s[0]*m[3];
    cv[4] = mo.p[4] - mo.s[1]*m[1]; cv[5] = mo.p[5] - mo.s[1]*m[2]; cv[6] = mo.p[6] - mo.s[1]*m[3];
    cv[7] = mo.p[7] - mo.s[2]*m[2]; cv[8] = mo.p[8] - mo.s[2]*m[3]; cv[9] = mo.p[9] - mo.s[3]*m[3];
    if (zero_ch == 0) { cv[0] = cv[1] = cv[2] = cv[3] = 0.0f; }
    if (zero_ch == 1) { cv[1] = cv[4] = cv[5] = cv[6] = 0.0f; }
    if (zero_ch == 2) { cv[2] = cv[5] = cv[7] = cv[8] = 0.0f; }
    if (zero_ch == 3) { cv[3] = cv[6] = cv[8] = cv[9] = 0.0f; }
    v[0] = cv[0]; v[1] = cv[1]; v[2] = cv[2]; v[3] = cv[3];
    float best = cv[0];
    if (cv[4] > best) { best = cv[4]; v[0] = cv[1]; v[1] = cv[4]; v[2] = cv[5]; v[3] = cv[6]; }
    if (cv[7] > best) { best = cv[7]; v[0] = cv[2]; v[1] = cv[5]; v[2] = cv[7]; v[3] = cv[8]; }
    if (cv[9] > best) { best = cv[9]; v[0] = cv[3]; v[1] = cv[6]; v[2] = cv[8]; v[3] = cv[9]; }
    float lam = 0.0f;
    for (int it = 0; it < iters; ++it) {
        const float n2 = v[0]*v[0] + v[1]*v[1] + v[2]*v[2] + v[3]*v[3];
        const float s2 = n2 > 1e-20f ? rsqrtf(n2) : 0.0f;
        const float a0 = v[0]*s2, a1 = v[1]*s2, a2 = v[2]*s2, a3 = v[3]*s2;
        v[0] = cv[0]*a0 + cv[1]*a1 + cv[2]*a2 + cv[3]*a3;
        v[1] = cv[1]*a0 + cv[4]*a1 + cv[5]*a2 + cv[6]*a3;
        v[2] = cv[2]*a0 + cv[5]*a1 + cv[7]*a2 + cv[8]*a3;
        v[3] = cv[3]*a0 + cv[6]*a1 + cv[8]*a2 + cv[9]*a3;
        lam = a0*v[0] + a1*v[1] + a2*v[2] + a3*v[3];
    }
    const float n2 = v[0]*v[0] + v[1]*v[1] + v[2]*v[2] + v[3]*v[3];
    const float s2 = n2 > 1e-20f ? rsqrtf(n2) : 0.0f;
    v[0] *= s2; v[1] *= s2; v[2] *= s2; v[3] *= s2;
    if (n2 <= 1e-20f) { v[0] = v[1] = v[2] = 0.57735f; v[3] = 0.0f; }
    if (v[0] + v[1] + v[2] < 0.0f) { v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2]; v[3] = -v[3]; }
    return fmaxf(cv[0] + cv[4] + cv[7] + cv[9] - lam, 0.0f);
}

// Full line fit; mo and out live in SHARED memory (no local-memory traffic).
__device__ __noinline__ void subset_line(const Mom& mo, int zero_ch, int iters, LineFit& out)
{
    float m[4], v[4];
    const float resid = line_core(mo, zero_ch, iters, m, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) { out.m[k] = m[k]; out.v[k] = v[k]; }
    out.resid = resid;
}

// Squared distance of a texel set from the best line THROUGH BLACK (RGB only): trace - lambda_max of the uncentred second
// moments, rebuilt from the moments about the block centre c.  mo lives in shared memory.
__device__ __noinline__ float origin_residual(const Mom& mo, float c0, float c1, float c2)
{
    const float n = mo.n, s0 = mo.s[0], s1 = mo.s[1], s2 = mo.s[2];
    const float m00 = mo.p[0] + 2.0f*c0*s0 + n*c0*c0, m01 = mo.p[1] + c0*s1 + c1*s0 + n*c0*c1, m02 = mo.p[2] + c0*s2 + c2*s0 + n*c0*c2;
    const float m11 = mo.p[4] + 2.0f*c1*s1 + n*c1*c1, m12 = mo.p[5] + c1*s2 + c2*s1 + n*c1*c2, m22 = mo.p[7] + 2.0f*c2*s2 + n*c2*c2;
    float v0 = s0 + n*c0, v1 = s1 + n*c1, v2 = s2 + n*c2, lam = 0.0f;      // start: the mean colour
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        const float n2 = v0*v0 + v1*v1 + v2*v2;
        const float is = n2 > 1e-20f ? rsqrtf(n2) : 0.0f;
        const float a0 = v0*is, a1 = v1*is, a2 = v2*is;
        v0 = m00*a0 + m01*a1 + m02*a2; v1 = m01*a0 + m11*a1 + m12*a2; v2 = m02*a0 + m12*a1 + m22*a2;
        lam = a0*v0 + a1*v1 + a2*v2;
    }
    return fmaxf(m00 + m11 + m22 - lam, 0.0f);
}

// Squared distance of a texel set from its best FREE line in RGB alone (alpha left out): trace - lambda_max of the RGB
// scatter about the mean.
__device__ __noinline__ float rgb_free_residual(const Mom& mo)
{
    const float inv = mo.n > 0.0f ? 1.0f/mo.n : 0.0f;
    const float c00 = mo.p[0] - mo.s[0]*mo.s[0]*inv, c01 = mo.p[1] - mo.s[0]*mo.s[1]*inv, c02 = mo.p[2] - mo.s[0]*mo.s[2]*inv;
    const float c11 = mo.p[4] - mo.s[1]*mo.s[1]*inv, c12 = mo.p[5] - mo.s[1]*mo.s[2]*inv, c22 = mo.p[7] - mo.s[2]*mo.s[2]*inv;
    float v0 = c00, v1 = c01, v2 = c02, best = c00, lam = 0.0f;
    if (c11 > best) { best = c11; v0 = c01; v1 = c11; v2 = c12; }
    if (c22 > best) { v0 = c02; v1 = c12; v2 = c22; }
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        const float n2 = v0*v0 + v1*v1 + v2*v2;
        const float is = n2 > 1e-20f ? rsqrtf(n2) : 0.0f;
        const float a0 = v0*is, a1 = v1*is, a2 = v2*is;
        v0 = c00*a0 + c01*a1 + c02*a2; v1 = c01*a0 + c11*a1 + c12*a2; v2 = c02*a0 + c12*a1 + c22*a2;
        lam = a0*v0 + a1*v1 + a2*v2;
    }
    return fmaxf(c00 + c11 + c22 - lam, 0.0f);
}

// Residual only, moments passed in registers (integer sums about the block centre).
// Channel error weights (sRGB textures, see launch_astc3): the hypotheses are built in the WEIGHTED colour space
// x'_c = sw_c x_c (sw = square roots of the weights; all 1.0f -- exact -- for linear textures).  The integer moments stay
// plain; they are scaled where they become floats: entry k of (n, s0..s3, p00 p01 p02 p03 p11 p12 p13 p22 p23 p33).
__device__ __forceinline__ float mom_scale(int k, float4 sw)
{
    const float a[4] = {sw.x, sw.y, sw.z, sw.w};
    if (k == 0) return 1.0f;
    if (k < 5) return a[k - 1];
    const int i = k < 9 ? 0 : (k < 12 ? 1 : (k < 14 ? 2 : 3));
    const int j = k < 9 ? k - 5 : (k < 12 ? k - 8 : (k < 14 ? k - 10 : 3));
    return a[i]*a[j];
}

__device__ __noinline__ float subset_resid(int n, int s0, int s1, int s2, int s3, int p0, int p1, int p2, int p3, int p4, int p5,
    int p6, int p7, int p8, int p9, float4 sw)
{
    Mom mo;
    mo.n = static_cast<float>(n);
    mo.s[0] = static_cast<float>(s0)*sw.x; mo.s[1] = static_cast<float>(s1)*sw.y; mo.s[2] = static_cast<float>(s2)*sw.z; mo.s[3] = static_cast<float>(s3)*sw.w;
    mo.p[0] = static_cast<float>(p0)*(sw.x*sw.x); mo.p[1] = static_cast<float>(p1)*(sw.x*sw.y); mo.p[2] = static_cast<float>(p2)*(sw.x*sw.z); mo.p[3] = static_cast<float>(p3)*(sw.x*sw.w);
    mo.p[4] = static_cast<float>(p4)*(sw.y*sw.y); mo.p[5] = static_cast<float>(p5)*(sw.y*sw.z); mo.p[6] = static_cast<float>(p6)*(sw.y*sw.w); mo.p[7] = static_cast<float>(p7)*(sw.z*sw.z);
    mo.p[8] = static_cast<float>(p8)*(sw.z*sw.w); mo.p[9] = static_cast<float>(p9)*(sw.w*sw.w);
    float m[4], v[4];
    return line_core(mo, -1, 4, m, v);
}
#define CFX_RESID15(a) subset_resid(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14], sw)

// A set of texels: one bit per texel, MW 64-bit words (footprints above 64 texels need 2 or 3).
template <int MW> struct TexelMask { uint64_t w[MW]; };

template <int MW>
__device__ __forceinline__ TexelMask<MW> load_mask(const Ctx& c, uint32_t off)
{
    TexelMask<MW> m;
#pragma unroll
    for (int k = 0; k < MW; ++k) m.w[k] = tab_u64(c, off + 8u*k);
    return m;
}
template <int MW>
__device__ __forceinline__ bool mask_empty(const TexelMask<MW>& m)
{
    uint64_t o = 0;
#pragma unroll
    for (int k = 0; k < MW; ++k) o |= m.w[k];
    return o == 0;
}
template <int MW>
__device__ __forceinline__ uint32_t mask_bit(const TexelMask<MW>& m, uint32_t i)
{
    uint64_t word = m.w[0];
#pragma unroll
    for (int k = 1; k < MW; ++k) if ((i >> 6) == static_cast<uint32_t>(k)) word = m.w[k];
    return static_cast<uint32_t>((word >> (i & 63u)) & 1ull);
}
// Number of texels a 2-subset seed assigns differently from the clustering (up to relabelling).
template <int MW>
__device__ __forceinline__ uint32_t mask_mismatch2(const TexelMask<MW>& km, const TexelMask<MW>& pm, uint32_t T)
{
    uint32_t d = 0;
#pragma unroll
    for (int k = 0; k < MW; ++k) d += popc64(km.w[k] ^ pm.w[k]);
    return min(d, T - d);
}
// Same for three subsets: best of the 6 label permutations.
template <int MW>
__device__ __forceinline__ uint32_t mask_mismatch3(const TexelMask<MW>& k1, const TexelMask<MW>& k2, const TexelMask<MW>& p1,
    const TexelMask<MW>& p2, const TexelMask<MW>& full, uint32_t T)
{
    uint32_t m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int w = 0; w < MW; ++w) {
        const uint64_t kk[3] = {full.w[w] ^ k1.w[w] ^ k2.w[w], k1.w[w], k2.w[w]};
        const uint64_t pp[3] = {full.w[w] ^ p1.w[w] ^ p2.w[w], p1.w[w], p2.w[w]};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) m[i][j] += popc64(kk[i] & pp[j]);
    }
    uint32_t a = m[0][0] + m[1][1] + m[2][2];
    a = max(a, m[0][0] + m[1][2] + m[2][1]);
    a = max(a, m[0][1] + m[1][0] + m[2][2]);
    a = max(a, m[0][1] + m[1][2] + m[2][0]);
    a = max(a, m[0][2] + m[1][0] + m[2][1]);
    a = max(a, m[0][2] + m[1][1] + m[2][0]);
    return T - a;
}

// Lane-local masked moments: texels whose bit is set in `mask`, about `ctr`.
template <int MW>
__device__ __forceinline__ void masked_moments(const int4* v, uint32_t T, int4 ctr, TexelMask<MW> mask, int (&acc)[15])
{
    int a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0, a8 = 0, a9 = 0, a10 = 0, a11 = 0, a12 = 0, a13 = 0, a14 = 0;
#pragma unroll
    for (int wd = 0; wd < MW; ++wd) {
    const uint64_t word = mask.w[wd];
    const uint32_t i1 = min(T, 64u*(wd + 1));
#pragma unroll 4
    for (uint32_t i = 64u*wd; i < i1; ++i) {
        const int4 x = v[i];
        const int f = static_cast<int>((word >> (i & 63u)) & 1ull);
        const int x0 = f*(x.x - ctr.x), x1 = f*(x.y - ctr.y), x2 = f*(x.z - ctr.z), x3 = f*(x.w - ctr.w);
        a0 += f; a1 += x0; a2 += x1; a3 += x2; a4 += x3;
        a5 += x0*x0; a6 += x0*x1; a7 += x0*x2; a8 += x0*x3; a9 += x1*x1;
        a10 += x1*x2; a11 += x1*x3; a12 += x2*x2; a13 += x2*x3; a14 += x3*x3;
    }
    }
    acc[0] = a0; acc[1] = a1; acc[2] = a2; acc[3] = a3; acc[4] = a4; acc[5] = a5; acc[6] = a6; acc[7] = a7; acc[8] = a8;
    acc[9] = a9; acc[10] = a10; acc[11] = a11; acc[12] = a12; acc[13] = a13; acc[14] = a14;
}

__device__ __forceinline__ void mom_from(const int (&a)[15], Mom& m)
{
    m.n = static_cast<float>(a[0]);
#pragma unroll
    for (int k = 0; k < 4; ++k) m.s[k] = static_cast<float>(a[1 + k]);
#pragma unroll
    for (int k = 0; k < 10; ++k) m.p[k] = static_cast<float>(a[5 + k]);
}

// k-means clustering of the block's texels (lane = texel) into k = 2 or 3 groups: farthest-point seeds, three
// Lloyd iterations; returns the texel masks of clusters 1 and 2.
template <int K, int MW>
__device__ __noinline__ void kmeans_warp(const int4* v, uint32_t T, uint32_t k, int4 mean, uint32_t lane, TexelMask<MW>& m1,
    TexelMask<MW>& m2)
{
    int4 x[K];
    bool live[K];
#pragma unroll
    for (int r = 0; r < K; ++r) { const uint32_t i = lane + 32u*r; live[r] = i < T; x[r] = v[live[r] ? i : 0u]; }
    float4 ctr[3];
    ctr[0] = make_float4(static_cast<float>(mean.x), static_cast<float>(mean.y), static_cast<float>(mean.z), static_cast<float>(mean.w));
    ctr[1] = ctr[0]; ctr[2] = ctr[0];
    for (uint32_t c = 0; c < k; ++c) {
        uint32_t key = 0;
#pragma unroll
        for (int r = 0; r < K; ++r) {
            if (!live[r]) continue;
            float d = 3.0e38f;
            for (uint32_t q = 0; q < (c == 0 ? 1u : c); ++q) {
                const float dx = x[r].x - ctr[q].x, dy = x[r].y - ctr[q].y, dz = x[r].z - ctr[q].z, dw = x[r].w - ctr[q].w;
                d = fminf(d, dx*dx + dy*dy + dz*dz + dw*dw);
            }
            // farthest texel, lowest index on ties
            const uint32_t kk = (__float_as_uint(d) & ~255u) | (255u - (lane + 32u*r));
            key = max(key, kk);
        }
        key = __reduce_max_sync(0xFFFFFFFFu, key);
        const int4 far = v[255u - (key & 255u)];
        ctr[c] = make_float4(static_cast<float>(far.x), static_cast<float>(far.y), static_cast<float>(far.z), static_cast<float>(far.w));
    }
    uint32_t b1[2*MW], b2[2*MW];
#pragma unroll
    for (int r = 0; r < 2*MW; ++r) b1[r] = b2[r] = 0;
    for (int it = 0; it < 3; ++it) {
        int cnt[3] = {0, 0, 0}, sx[3] = {0, 0, 0}, sy[3] = {0, 0, 0}, sz[3] = {0, 0, 0}, sw[3] = {0, 0, 0};
        uint32_t lab[K];
#pragma unroll
        for (int r = 0; r < K; ++r) {
            uint32_t bi = 0;
            float bd = 3.0e38f;
            for (uint32_t q = 0; q < k; ++q) {
                const float dx = x[r].x - ctr[q].x, dy = x[r].y - ctr[q].y, dz = x[r].z - ctr[q].z, dw = x[r].w - ctr[q].w;
                const float d = dx*dx + dy*dy + dz*dz + dw*dw;
                if (d < bd) { bd = d; bi = q; }
            }
            lab[r] = live[r] ? bi : 3u;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int f = lab[r] == static_cast<uint32_t>(q) ? 1 : 0;
                cnt[q] += f; sx[q] += f*x[r].x; sy[q] += f*x[r].y; sz[q] += f*x[r].z; sw[q] += f*x[r].w;
            }
        }
#pragma unroll
        for (int r = 0; r < K; ++r) { b1[r] = __ballot_sync(0xFFFFFFFFu, lab[r] == 1u); b2[r] = __ballot_sync(0xFFFFFFFFu, lab[r] == 2u); }
        for (uint32_t q = 0; q < k; ++q) {
            const int n = redux_add(cnt[q]);
            const int ax = redux_add(sx[q]), ay = redux_add(sy[q]), az = redux_add(sz[q]), aw = redux_add(sw[q]);
            if (n > 0) {
                const float ic = 1.0f/static_cast<float>(n);
                ctr[q] = make_float4(ax*ic, ay*ic, az*ic, aw*ic);
            }
        }
    }
#pragma unroll
    for (int w = 0; w < MW; ++w) {
        m1.w[w] = static_cast<uint64_t>(b1[2*w]) | (static_cast<uint64_t>(b1[2*w + 1]) << 32);
        m2.w[w] = static_cast<uint64_t>(b2[2*w]) | (static_cast<uint64_t>(b2[2*w + 1]) << 32);
    }
}

struct SlotView {           // what pack_block needs from a slot
    uint32_t pc, seed;
    int32_t dual_ch;
};

// ---- warp-cooperative block packing ----------------------------------------------------------------
// The same 128 bits astc_core.cuh's pack_block() writes, but lane = value: every BISE element's position in the stream is
// closed form (a trit group of five takes 5 b + 8 bits, a quint group of three 3 b + 7), so each lane places its own
// element -- and its share of the group's trit / quint word -- into a private 128-bit word and four redux.or's merge
// them.  The serial version was 2600 single-lane instructions per block behind a chain of dependent table loads.
__device__ __forceinline__ void put128(uint64_t& lo, uint64_t& hi, uint32_t pos, uint32_t v, uint32_t n)
{
    if (n == 0) return;
    const uint64_t vv = static_cast<uint64_t>(v) & ((1ull << n) - 1ull);
    if (pos < 64) {
        lo |= vv << pos;
        if (pos + n > 64) hi |= vv >> (64u - pos);
    } else hi |= vv << (pos - 64u);
}

// Element i of a BISE sequence of n values starting at bit pos0; E[] holds the encoded values (low `bits` bits = the
// plain part, the rest = the trit / quint).
__device__ __forceinline__ void ise_piece(const Ctx& c, const uint8_t* E, uint32_t i, uint32_t n, uint32_t pos0, uint32_t bits,
    bool trits, bool quints, uint64_t& lo, uint64_t& hi)
{
    const uint32_t mask = (1u << bits) - 1u;
    if (trits) {
        const uint32_t g = i/5u, k = i - 5u*g;
        uint32_t v[5];
#pragma unroll
        for (uint32_t q = 0; q < 5; ++q) v[q] = 5u*g + q < n ? E[5u*g + q] : 0u;
        const uint32_t tw = tab_u8(c, c.tab.off_trit_enc + (v[0] >> bits) + 3u*(v[1] >> bits) + 9u*(v[2] >> bits) + 27u*(v[3] >> bits) +
            81u*(v[4] >> bits));
        const uint32_t sh = (0x75420u >> (4u*k)) & 15u, nb = (0x12122u >> (4u*k)) & 15u;
        const uint32_t pos = pos0 + g*(5u*bits + 8u) + k*bits + sh;
        put128(lo, hi, pos, E[i] & mask, bits);
        put128(lo, hi, pos + bits, tw >> sh, nb);
    } else if (quints) {
        const uint32_t g = i/3u, k = i - 3u*g;
        uint32_t v[3];
#pragma unroll
        for (uint32_t q = 0; q < 3; ++q) v[q] = 3u*g + q < n ? E[3u*g + q] : 0u;
        const uint32_t qw = tab_u8(c, c.tab.off_quint_enc + (v[0] >> bits) + 5u*(v[1] >> bits) + 25u*(v[2] >> bits));
        const uint32_t sh = (0x530u >> (4u*k)) & 15u, nb = (0x223u >> (4u*k)) & 15u;
        const uint32_t pos = pos0 + g*(3u*bits + 7u) + k*bits + sh;
        put128(lo, hi, pos, E[i] & mask, bits);
        put128(lo, hi, pos + bits, qw >> sh, nb);
    } else put128(lo, hi, pos0 + i*bits, E[i], bits);
}

// ep: [subset][e0 rgba, e1 rgba] (8 ints per subset); sk: the weights' ranks in bit-stream order; ecol / ewgt: scratch
// for the encoded colour values (>= 18 bytes) and weights (>= nw*planes bytes).  The other arguments as pack_block().
__device__ __forceinline__ uint4 pack_block_warp(const Ctx& c, const SlotView& slot, const ModeInfo& m, uint32_t cl, const int* ep,
    bool has_alpha, const uint8_t* sk, bool lum, const int* hdr_vals, const uint8_t* cems, const int* scales, uint32_t contracted,
    const int* raw_vals, uint8_t* ecol, uint8_t* ewgt, uint32_t lane)
{
    uint64_t lo = 0, hi = 0;
    const uint32_t pc = slot.pc;
    uint32_t cem[4];
    bool same = true;
    for (uint32_t s = 0; s < pc; ++s) {
        cem[s] = cems ? cems[s] : (hdr_vals ? (has_alpha ? 14u : 11u) : (lum ? (has_alpha ? 4u : 0u) : (has_alpha ? 12u : 8u)));
        same = same && cem[s] == cem[0];
    }
    // header (the same bits in every lane: or is idempotent)
    put128(lo, hi, 0, m.mode_bits, 11);
    put128(lo, hi, 11, pc - 1, 2);
    uint32_t pos;
    uint32_t below = 128u - m.wbits;                  // first free bit below the weights
    if (pc == 1) { put128(lo, hi, 13, cem[0], 4); pos = 17; }
    else {
        put128(lo, hi, 13, slot.seed, 10);
        if (same) put128(lo, hi, 25, cem[0], 4);
        else {
            uint32_t low = 4;
            for (uint32_t s = 0; s < pc; ++s) low = low < (cem[s] >> 2) ? low : (cem[s] >> 2);
            if (low == 3) low = 2;
            uint32_t enc = low + 1u, bp = 2;
            for (uint32_t s = 0; s < pc; ++s) enc |= ((cem[s] >> 2) - low) << bp++;
            for (uint32_t s = 0; s < pc; ++s) { enc |= (cem[s] & 3u) << bp; bp += 2; }
            const uint32_t hi_bits = 3u*pc - 4u;
            put128(lo, hi, 23, enc & 0x3Fu, 6);
            below -= hi_bits;
            put128(lo, hi, below, enc >> 6, hi_bits);
        }
        pos = 29;
    }
    uint32_t start[5];                                  // first value of every subset
    start[0] = 0;
    for (uint32_t s = 0; s < pc; ++s) start[s + 1] = start[s] + ((cem[s] >> 2) + 1u)*2u;
    const uint32_t ncol = start[pc];
    const uint32_t planes = slot.dual_ch >= 0 ? 2u : 1u;
    if (planes == 2) put128(lo, hi, below - 2u, static_cast<uint32_t>(slot.dual_ch), 2);
    const uint32_t L = m.level, nwt = m.nw*planes;
    // encoded values: lane = colour value, lane (+32) = weight
    if (lane < ncol) {
        uint32_t s = 0;
        while (s + 1u < pc && lane >= start[s + 1]) ++s;
        const uint32_t k = lane - start[s];
        const int* e = ep + s*8u;
        uint32_t val;
        if (hdr_vals) val = static_cast<uint32_t>(hdr_vals[s*8u + k]) & 0xFFu;
        else if ((contracted >> s) & 1u) val = static_cast<uint32_t>(raw_vals[s*8u + k]) & 0xFFu;
        else if (cem[s] == 6u || cem[s] == 10u) val = static_cast<uint32_t>(k < 3u ? e[4u + k] : (k == 3u ? scales[s] : e[4u*(k & 1u) + 3u])) & 0xFFu;
        else if (cem[s] == 4u) val = static_cast<uint32_t>(k < 2u ? e[4u*k] : e[4u*(k & 1u) + 3u]) & 0xFFu;
        else val = static_cast<uint32_t>(e[4u*(k & 1u) + (k >> 1)]) & 0xFFu;
        const uint32_t rank = tab_u8(c, c.tab.off_cq_near + cl*256u + val);
        ecol[lane] = static_cast<uint8_t>(tab_u8(c, c.tab.off_cq_enc + cl*256u + rank));
    }
    for (uint32_t j = lane; j < nwt; j += 32) ewgt[j] = static_cast<uint8_t>(tab_u8(c, c.tab.off_wq_enc + L*32u + sk[j]));
    __syncwarp();
    if (lane < ncol) ise_piece(c, ecol, lane, ncol, pos, kCqBits[cl], kCqTrits[cl] != 0, kCqQuints[cl] != 0, lo, hi);
    // weights: BISE from bit 0 of a scratch word, mirrored into the top of the block
    uint64_t wlo = 0, whi = 0;
    for (uint32_t j = lane; j < nwt; j += 32) ise_piece(c, ewgt, j, nwt, 0, kWqBits[L], kWqTrits[L] != 0, kWqQuints[L] != 0, wlo, whi);
    hi |= (static_cast<uint64_t>(__brev(static_cast<uint32_t>(wlo))) << 32) | __brev(static_cast<uint32_t>(wlo >> 32));
    lo |= (static_cast<uint64_t>(__brev(static_cast<uint32_t>(whi))) << 32) | __brev(static_cast<uint32_t>(whi >> 32));
    uint4 out;
    out.x = __reduce_or_sync(0xFFFFFFFFu, static_cast<uint32_t>(lo));
    out.y = __reduce_or_sync(0xFFFFFFFFu, static_cast<uint32_t>(lo >> 32));
    out.z = __reduce_or_sync(0xFFFFFFFFu, static_cast<uint32_t>(hi));
    out.w = __reduce_or_sync(0xFFFFFFFFu, static_cast<uint32_t>(hi >> 32));
    return out;
}

} // namespace

template <int NT, int KS, int W, int CTAS, bool LOCK, bool HDR, bool STAGE>
__global__ void __launch_bounds__(W*32, CTAS) astc3_kernel(const __grid_constant__ EncodeParams p, const __grid_constant__ Tab3 tb, uint32_t n_exact, uint32_t refine)
{
    constexpr int K = (NT*8 + 31)/32;            // texels per lane
    constexpr int MW = (NT*8 + 63)/64;           // words per texel mask
    static_assert(2*MW >= K, "ballot words");
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = lane_id(), warp = warp_id();
    const uint32_t gq = lane >> 2, tq = lane & 3u;
    using WS = Warp3T<NT>;
    constexpr size_t kWsBytes = (sizeof(WS) + 15)/16*16;
    constexpr uint32_t TP = WS::TP;
    // (keeping lane, warp and the working-set offset in registers behind an opaque asm -- the compiler re-reads the thread
    // id and rebuilds smem + warp*kWsBytes + field in ~170 places, 8 % of the executed instructions -- was measured:
    // 2 - 5 % slower at 20 warps / 96 registers, the three registers cost more in spills than the recomputation; at 16
    // warps / 128 registers it gains 3 %, which still leaves 16 warps 4 % behind 20)
    WS& ws = *reinterpret_cast<WS*>(smem + warp*kWsBytes);
    // STAGE: behind the warps' working sets, two buffers of [R tiles | M tiles] of one weight grid and their mbarriers
    constexpr uint32_t kTileBytes = NT*KS*256;                    // the R (or at most the M) fragments of one grid
    uint8_t* const stage = smem + W*kWsBytes;
    uint64_t* const stage_bar = reinterpret_cast<uint64_t*>(stage + 4*kTileBytes);
    uint32_t staged = 0;                                          // grids staged so far by this CTA: buffer and parity
    if (STAGE && threadIdx.x == 0) {
        mbar_init(&stage_bar[0], 1); mbar_init(&stage_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const Ctx& ctx = tb.ctx;
    const uint32_t T = ctx.tab.texels, bw = ctx.tab.bw, bh = ctx.tab.bh;
    const uint32_t G = ctx.tab.n_grids;
    const bool alpha_off = p.alpha_type == 0;
    // square roots of the channel error weights and their reciprocals (1.0f for linear textures): read from the kernel
    // parameters wherever they are used -- as locals they cost 8 registers the kernel does not have (-8 % throughput)
    const float4& sw = tb.sw;
    const float4& isw = tb.isw;
    const float fx2 = static_cast<float>(FX*FX);
    const float ifx = 1.0f/static_cast<float>(FX);
    // setup scratch lives where phase 1 later writes D
    LineFit* lines = ws.u.setup.lines;
    Mom* moms = ws.u.setup.moms;
    // the A operand's padding columns must be finite (they meet zero B entries)
    for (uint32_t i = lane; i < kRows3*WS::TS; i += 32) (&ws.ta[0][0])[i] = __float2half_rn(0.0f);

    for (uint32_t base = blockIdx.x*W; base < p.total_blocks; base += gridDim.x*W) {
        // every warp of the CTA runs every iteration (the barriers below need that): warps past the end redo the last
        // block without storing it
        const uint32_t blk = min(base + warp, p.total_blocks - 1u);
        bool active = base + warp < p.total_blocks;
        const uint32_t by = blk / p.blocks_x, bx = blk - by*p.blocks_x;
        PHASE_SYNC();
        bool differs = false, alpha = false;
        const float opaque = HDR ? 255.0f*kHdrAlphaScale : 255.0f;
        float4 first = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll 1
        for (uint32_t i0 = 0; i0 < T; i0 += 32) {
            const uint32_t i = i0 + lane;
            float4 v = make_float4(0.0f, 0.0f, 0.0f, opaque);
            if (i < T) {
                const uint32_t ty = i / bw, tx = i - ty*bw;
                const uint32_t x = min(bx*bw + tx, p.width - 1), y = min(by*bh + ty, p.height - 1);
                if (p.src_format == SRC_RGBA8) {
                    const uint32_t q = __ldg(reinterpret_cast<const uint32_t*>(p.src + static_cast<uint64_t>(y)*p.pitch) + x);
                    v = make_float4(static_cast<float>(q & 0xFF), static_cast<float>((q >> 8) & 0xFF), static_cast<float>((q >> 16) & 0xFF),
                        static_cast<float>(q >> 24));
                } else {
                    const float4 f = load_texel_f32(p, x, y);
                    v = make_float4(fminf(fmaxf(f.x, 0.0f), 1.0f)*255.0f, fminf(fmaxf(f.y, 0.0f), 1.0f)*255.0f,
                        fminf(fmaxf(f.z, 0.0f), 1.0f)*255.0f, fminf(fmaxf(f.w, 0.0f), 1.0f)*255.0f);
                }
                if (HDR) {
                    // HDR: colour as 16-bit LNS scaled to the 8-bit-like range the search works in (LNS/256); alpha stays
                    // an LDR channel (end point mode 14 when the block is not opaque), scaled down so that its error
                    // weighs about what the reference's HDR-alpha metric gives it
                    const float4 f = load_texel_f32(p, x, y);
                    v = make_float4(lns_from_float(f.x)*(1.0f/256.0f), lns_from_float(f.y)*(1.0f/256.0f), lns_from_float(f.z)*(1.0f/256.0f),
                        fminf(fmaxf(f.w, 0.0f), 1.0f)*(255.0f*kHdrAlphaScale));
                }
                if (!(p.color_mask & 1u)) v.x = 0.0f;
                if (!(p.color_mask & 2u)) v.y = 0.0f;
                if (!(p.color_mask & 4u)) v.z = 0.0f;
                if (!(p.color_mask & 8u)) v.w = 0.0f; else if (alpha_off) v.w = opaque;
                ws.v[i] = make_int4(__float2int_rn(v.x*FX), __float2int_rn(v.y*FX), __float2int_rn(v.z*FX), __float2int_rn(v.w*FX));
            }
            if (i0 == 0) {
                first.x = __shfl_sync(0xFFFFFFFFu, v.x, 0); first.y = __shfl_sync(0xFFFFFFFFu, v.y, 0);
                first.z = __shfl_sync(0xFFFFFFFFu, v.z, 0); first.w = __shfl_sync(0xFFFFFFFFu, v.w, 0);
            }
            if (i < T) {
                differs |= v.x != first.x || v.y != first.y || v.z != first.z || v.w != first.w;
                alpha |= v.w != opaque;
            }
        }
        __syncwarp();
        const bool constant = !__any_sync(0xFFFFFFFFu, differs);
        const bool has_alpha = __any_sync(0xFFFFFFFFu, alpha);
        uint4* dst = reinterpret_cast<uint4*>(p.dst) + blk;
        if (constant) {
            if (active && lane == 0) {
                if (HDR) {
                    const float4 f = load_texel_f32(p, min(bx*bw, p.width - 1), min(by*bh, p.height - 1));
                    auto hb = [](float q) -> uint32_t { return __half_as_ushort(__float2half_rn(fminf(fmaxf(q, 0.0f), 65504.0f))); };
                    *dst = pack_void_extent_hdr((p.color_mask & 1u) ? hb(f.x) : 0u, (p.color_mask & 2u) ? hb(f.y) : 0u,
                        (p.color_mask & 4u) ? hb(f.z) : 0u, !(p.color_mask & 8u) ? 0u : (alpha_off ? 0x3C00u : hb(fminf(f.w, 1.0f))));
                } else *dst = pack_void_extent(first);
            }
            active = false;
        }
        const uint32_t nch = has_alpha ? 4u : 3u;

        // ---- setup 1: block moments about the (integer) mean, channel ranges
        int4 ctr;
        int tot[15];
        int lo4[4], hi4[4];
        if (active) {
            int sx = 0, sy = 0, sz = 0, sw = 0;
            int mn[4] = {1 << 30, 1 << 30, 1 << 30, 1 << 30}, mx[4] = {-(1 << 30), -(1 << 30), -(1 << 30), -(1 << 30)};
            for (uint32_t i = lane; i < T; i += 32) {
                const int4 x = ws.v[i];
                sx += x.x; sy += x.y; sz += x.z; sw += x.w;
                mn[0] = min(mn[0], x.x); mn[1] = min(mn[1], x.y); mn[2] = min(mn[2], x.z); mn[3] = min(mn[3], x.w);
                mx[0] = max(mx[0], x.x); mx[1] = max(mx[1], x.y); mx[2] = max(mx[2], x.z); mx[3] = max(mx[3], x.w);
            }
            const int it = static_cast<int>(T);
            ctr = make_int4((redux_add(sx) + it/2)/it, (redux_add(sy) + it/2)/it, (redux_add(sz) + it/2)/it, (redux_add(sw) + it/2)/it);
#pragma unroll
            for (int k = 0; k < 4; ++k) { lo4[k] = __reduce_min_sync(0xFFFFFFFFu, mn[k]); hi4[k] = __reduce_max_sync(0xFFFFFFFFu, mx[k]); }
            int acc[15];
#pragma unroll
            for (int k = 0; k < 15; ++k) acc[k] = 0;
            for (uint32_t i = lane; i < T; i += 32) {
                const int4 x = ws.v[i];
                const int x0 = x.x - ctr.x, x1 = x.y - ctr.y, x2 = x.z - ctr.z, x3 = x.w - ctr.w;
                acc[0] += 1; acc[1] += x0; acc[2] += x1; acc[3] += x2; acc[4] += x3;
                acc[5] += x0*x0; acc[6] += x0*x1; acc[7] += x0*x2; acc[8] += x0*x3; acc[9] += x1*x1;
                acc[10] += x1*x2; acc[11] += x1*x3; acc[12] += x2*x2; acc[13] += x2*x3; acc[14] += x3*x3;
            }
#pragma unroll
            for (int k = 0; k < 15; ++k) tot[k] = redux_add(acc[k]);
        }
        // ---- setup 2: lines of the single-subset slot (lane 0) and of the four dual-plane slots (lanes 1..4)
        if (active) {
#pragma unroll
            for (int k = 0; k < 15; ++k) if (lane == static_cast<uint32_t>(k)) (&moms[10].n)[k] = static_cast<float>(tot[k])*mom_scale(k, sw);
            __syncwarp();
            if (lane < 5) subset_line(moms[10], static_cast<int>(lane) - 1, 6, lines[lane]);
        }
        if (active && lane < kSlots3) {
            Slot3& sl = ws.slots[lane];
            sl.pc = 1; sl.seed = 0; sl.dual_ch = (lane >= 5 && lane < 9) ? static_cast<int32_t>(lane - 5) : -1;
            sl.valid = lane == 0 || (lane >= 5 && lane < 9 && lane - 5 < nch) || (lane == kLumSlot && !(tb.flags & 1u) && !HDR) ? 1u : 0u;
            // (slots 10..13 are validated in setup 8, after the partitionings are known)
        }
        // ---- setup 3: cluster the texels into 2 and 3 groups, match the partition seeds against the clusters
        TexelMask<MW> km0, km1, km2;
#pragma unroll
        for (int w = 0; w < MW; ++w) km0.w[w] = km1.w[w] = km2.w[w] = 0;
#pragma unroll 1
        for (uint32_t kk = 2; active && kk <= 3; ++kk) {
            TexelMask<MW> ka, kb;
            kmeans_warp<K, MW>(ws.v, T, kk, ctr, lane, ka, kb);
            if (kk == 2) km0 = ka; else { km1 = ka; km2 = kb; }
        }
        uint32_t b2 = 0xFFFFFFFFu, b3 = 0xFFFFFFFFu;
        if (active) {
            TexelMask<MW> full;
#pragma unroll
            for (int w = 0; w < MW; ++w) {
                const uint32_t left = T > 64u*w ? T - 64u*w : 0u;
                full.w[w] = left >= 64u ? ~0ull : ((1ull << left) - 1ull);
            }
            // compacted seed lists: every lane scores a usable seed in every iteration
#pragma unroll 4
            for (uint32_t i = lane; i < tb.t3.n_seed2; i += 32) {
                const uint32_t seed = tab_u16(ctx, tb.t3.off_seed2 + i*2u);
                const TexelMask<MW> q = load_mask<MW>(ctx, tb.t3.off_part2c + i*(8u*MW));
                b2 = min(b2, (mask_mismatch2<MW>(km0, q, T) << 10) | seed);
            }
#pragma unroll 2
            for (uint32_t i = lane; i < tb.t3.n_seed3; i += 32) {
                const uint32_t seed = tab_u16(ctx, tb.t3.off_seed3 + i*2u);
                const TexelMask<MW> q1 = load_mask<MW>(ctx, tb.t3.off_part3c + i*(16u*MW));
                const TexelMask<MW> q2 = load_mask<MW>(ctx, tb.t3.off_part3c + i*(16u*MW) + 8u*MW);
                b3 = min(b3, (mask_mismatch3<MW>(km1, km2, q1, q2, full, T) << 10) | seed);
            }
        }
        PHASE_SYNC();
        // A texel the seed assigns differently from the clustering costs a fraction of the block's mean squared spread:
        // on gray (one-dimensional) content every partitioning has a zero line-fit residual and only this term tells
        // the seed that follows the clusters from an arbitrary one.
        const float mis_unit = tb.mis_w*static_cast<float>(tot[5] + tot[9] + tot[12] + tot[14])/static_cast<float>(T);
        // ---- setup 4: exact line-fit residual of every lane's two-subset seed; the two best become slots 1, 2.
        //      (Moments stay in registers; only the residual survives: the winners' moments are re-derived in setup 6.)
        if (active) {
            float sc = 3.0e38f;
            if (b2 != 0xFFFFFFFFu) {
                int a1[15];
                masked_moments<MW>(ws.v, T, ctr, load_mask<MW>(ctx, tb.t3.off_part2w + (b2 & 1023u)*(8u*MW)), a1);
                if (a1[0] >= 1 && tot[0] - a1[0] >= 1) {
                    const float r1 = CFX_RESID15(a1);
#pragma unroll
                    for (int k = 0; k < 15; ++k) a1[k] = tot[k] - a1[k];
                    sc = r1 + CFX_RESID15(a1) + mis_unit*static_cast<float>(b2 >> 10);
                }
            }
            for (uint32_t rank = 0; rank < 2; ++rank) {
                const uint32_t kmin = __reduce_min_sync(0xFFFFFFFFu, (__float_as_uint(sc) & ~31u) | lane);
                const uint32_t wl = kmin & 31u;
                const bool ok = __shfl_sync(0xFFFFFFFFu, sc, wl) < 3.0e38f;
                const uint32_t seed = __shfl_sync(0xFFFFFFFFu, b2, wl) & 1023u;
                if (lane == wl) sc = 3.0e38f;
                if (lane == 0) { Slot3& sl = ws.slots[1 + rank]; sl.pc = 2; sl.seed = seed; sl.dual_ch = -1; sl.valid = ok ? 1u : 0u; }
                const TexelMask<MW> m1 = load_mask<MW>(ctx, tb.t3.off_part2w + seed*(8u*MW));
                for (uint32_t i = lane; i < TP; i += 32) ws.part[rank][i] = static_cast<uint8_t>(mask_bit<MW>(m1, i));
            }
        }
        PHASE_SYNC();
        // ---- setup 5: same for the three-subset seeds -> slots 3, 4
        if (active) {
            float sc = 3.0e38f;
            if (b3 != 0xFFFFFFFFu) {
                const uint32_t seed = b3 & 1023u;
                int a1[15], a2[15];
                masked_moments<MW>(ws.v, T, ctr, load_mask<MW>(ctx, tb.t3.off_part3w + seed*(16u*MW)), a1);
                const float r1 = a1[0] >= 1 ? CFX_RESID15(a1) : 3.0e38f;
                masked_moments<MW>(ws.v, T, ctr, load_mask<MW>(ctx, tb.t3.off_part3w + seed*(16u*MW) + 8u*MW), a2);
                const float r2 = a2[0] >= 1 ? CFX_RESID15(a2) : 3.0e38f;
#pragma unroll
                for (int k = 0; k < 15; ++k) a1[k] = tot[k] - a1[k] - a2[k];
                if (a1[0] >= 1 && r1 < 3.0e38f && r2 < 3.0e38f) sc = r1 + r2 + CFX_RESID15(a1) + mis_unit*static_cast<float>(b3 >> 10);
            }
            for (uint32_t rank = 0; rank < 2; ++rank) {
                const uint32_t kmin = __reduce_min_sync(0xFFFFFFFFu, (__float_as_uint(sc) & ~31u) | lane);
                const uint32_t wl = kmin & 31u;
                const bool ok = __shfl_sync(0xFFFFFFFFu, sc, wl) < 3.0e38f;
                const uint32_t seed = __shfl_sync(0xFFFFFFFFu, b3, wl) & 1023u;
                if (lane == wl) sc = 3.0e38f;
                if (lane == 0) { Slot3& sl = ws.slots[3 + rank]; sl.pc = 3; sl.seed = seed; sl.dual_ch = -1; sl.valid = ok ? 1u : 0u; }
                const TexelMask<MW> m1 = load_mask<MW>(ctx, tb.t3.off_part3w + seed*(16u*MW));
                const TexelMask<MW> m2 = load_mask<MW>(ctx, tb.t3.off_part3w + seed*(16u*MW) + 8u*MW);
                for (uint32_t i = lane; i < TP; i += 32)
                    ws.part[2 + rank][i] = static_cast<uint8_t>(mask_bit<MW>(m1, i) ? 1u : (mask_bit<MW>(m2, i) ? 2u : 0u));
            }
        }
        PHASE_SYNC();
        // ---- setup 6: moments of the ten subsets of slots 1..4 (lane = texel, redux per field), then their lines
        //      (lane = subset)
        if (active) {
#pragma unroll 1
            for (uint32_t job = 0; job < 10; ++job) {
                const uint32_t k = job < 2 ? 0u : (job < 4 ? 1u : (job < 7 ? 2u : 3u));          // part index = slot - 1
                const uint32_t q = job < 4 ? (job & 1u) : (job < 7 ? job - 4u : job - 7u);       // subset
                if (!ws.slots[1 + k].valid) continue;
                int acc[15];
#pragma unroll
                for (int f = 0; f < 15; ++f) acc[f] = 0;
                for (uint32_t i = lane; i < T; i += 32) {
                    if (ws.part[k][i] != q) continue;
                    const int4 x = ws.v[i];
                    const int x0 = x.x - ctr.x, x1 = x.y - ctr.y, x2 = x.z - ctr.z, x3 = x.w - ctr.w;
                    acc[0] += 1; acc[1] += x0; acc[2] += x1; acc[3] += x2; acc[4] += x3;
                    acc[5] += x0*x0; acc[6] += x0*x1; acc[7] += x0*x2; acc[8] += x0*x3; acc[9] += x1*x1;
                    acc[10] += x1*x2; acc[11] += x1*x3; acc[12] += x2*x2; acc[13] += x2*x3; acc[14] += x3*x3;
                }
#pragma unroll
                for (int f = 0; f < 15; ++f) {
                    const int t = redux_add(acc[f]);
                    if (lane == static_cast<uint32_t>(f)) (&moms[job].n)[f] = static_cast<float>(t)*mom_scale(f, sw);
                }
            }
            __syncwarp();
            uint32_t opq2[2] = {0u, 0u};
            if (has_alpha) {
                const int full = __float2int_rn(opaque*static_cast<float>(FX));
#pragma unroll
                for (uint32_t k = 0; k < 2; ++k) {
                    bool t0 = false, t1 = false;                        // "has a translucent texel"
                    for (uint32_t i = lane; i < T; i += 32) {
                        const bool tr = ws.v[i].w != full;
                        if (ws.part[k][i]) t1 |= tr; else t0 |= tr;
                    }
                    opq2[k] = (__any_sync(0xFFFFFFFFu, t0) ? 0u : 1u) | (__any_sync(0xFFFFFFFFu, t1) ? 0u : 2u);
                }
            }
            if (lane < 10) {
                const uint32_t sl = lane < 2 ? 1u : (lane < 4 ? 2u : (lane < 7 ? 3u : 4u));
                if (ws.slots[sl].valid) {
                    subset_line(moms[lane], -1, 6, lines[5 + lane]);
                    // for the base + scale siblings (slots 14..17): the subset's distance from its line through black
                    // (blocks with alpha, CEM 10: what the RGB part loses by going through black, on top of the RGBA line's floor)
                    const float org = origin_residual(moms[lane], sw.x*static_cast<float>(ctr.x), sw.y*static_cast<float>(ctr.y), sw.z*static_cast<float>(ctr.z));
                    lines[5 + lane].pad[0] = has_alpha ? fmaxf(org - rgb_free_residual(moms[lane]), 0.0f) : org;
                }
            }
            __syncwarp();
            if (lane < 4) {
                const uint32_t first = lane == 0 ? 5u : (lane == 1 ? 7u : (lane == 2 ? 9u : 12u)), cnt = lane < 2 ? 2u : 3u;
                float e = 0.0f;
                for (uint32_t q = 0; q < cnt; ++q) e += lines[first + q].pad[0];
                ws.scale_eline[1 + lane] = e*ifx*ifx;
                ws.scale_valid[1 + lane] = ws.slots[1 + lane].valid && !HDR && !(tb.flags & 4u) ? 1u : 0u;
            } else if (lane == 4) {
                // one subset: CEM 6, or CEM 10 in a block with alpha (the alpha pair stays direct: what the RGB part
                // loses by going through black comes on top of the RGBA line's floor)
                const float org = origin_residual(moms[10], sw.x*static_cast<float>(ctr.x), sw.y*static_cast<float>(ctr.y), sw.z*static_cast<float>(ctr.z));
                ws.scale_eline[0] = org*ifx*ifx;
                ws.scale_valid[0] = !HDR && !(tb.flags & 4u) ? 1u : 0u;
                if (has_alpha) ws.scale_eline[0] = fmaxf(org - rgb_free_residual(moms[10]), 0.0f)*ifx*ifx;     // + slot 0's floor, added in phase 1c
            } else if (lane == 5 || lane == 6) {
                // mixed modes on the two-subset partitionings (slots 19, 20)
                const uint32_t k = lane - 5u, first = k == 0 ? 5u : 7u;
                uint32_t ok = ws.slots[1 + k].valid && !HDR && !(tb.flags & 4u) ? 1u : 0u, mask = 0u;
                float e = 0.0f;
                if (!has_alpha) {
                    const float d0 = lines[first].pad[0] - lines[first].resid, d1 = lines[first + 1].pad[0] - lines[first + 1].resid;
                    mask = d0 <= d1 ? 1u : 2u;
                    e = (mask == 1u ? lines[first].pad[0] + lines[first + 1].resid : lines[first].resid + lines[first + 1].pad[0])*ifx*ifx;
                } else {
                    mask = opq2[k];                                     // subsets whose alpha is 255 throughout
                    uint32_t kind = 0u;
                    e = lines[first].resid + lines[first + 1].resid;
                    if (mask == 0u) {
                        // no opaque subset: the one that loses least by it goes through black (CEM 10 + 12)
                        kind = 1u;
                        mask = lines[first].pad[0] <= lines[first + 1].pad[0] ? 1u : 2u;
                        e += mask == 1u ? lines[first].pad[0] : lines[first + 1].pad[0];
                    } else if (mask == 3u) ok = 0u;
                    e *= ifx*ifx;
                    ws.mix_kind[k] = kind;
                }
                ws.scale_eline[5 + k] = e; ws.scale_valid[5 + k] = ok; ws.mix_mask[k] = mask;
            }
        }
        __syncwarp();
        // ---- setup 7: per slot, project the texels on their subset's line -> ideal weights (fp16 A operand rows),
        //      end points, line lengths
        for (uint32_t s = 0; s < kSlots; ++s) {
            if (CFX_ASTC3_SLOTSYNC) PHASE_SYNC();
            if (!active) continue;
            Slot3& slot = ws.slots[s];
            if (!slot.valid) continue;
            const uint32_t pc = slot.pc;
            const uint32_t base = s == 0 ? 0u : (s >= 5 ? s - 4u : (s == 1 ? 5u : (s == 2 ? 7u : (s == 3 ? 9u : 12u))));
            float tl[K];
            uint32_t ql[K];
#pragma unroll
            for (int r = 0; r < K; ++r) {
                const uint32_t i = lane + 32u*r;
                tl[r] = 0.0f; ql[r] = 3u;
                if (i < T) {
                    const uint32_t q = pc > 1 ? ws.part[s - 1][i] : 0u;
                    const LineFit& lf = lines[base + q];
                    const int4 x = ws.v[i];
                    tl[r] = (sw.x*static_cast<float>(x.x - ctr.x) - lf.m[0])*lf.v[0] + (sw.y*static_cast<float>(x.y - ctr.y) - lf.m[1])*lf.v[1] +
                        (sw.z*static_cast<float>(x.z - ctr.z) - lf.m[2])*lf.v[2] + (sw.w*static_cast<float>(x.w - ctr.w) - lf.m[3])*lf.v[3];
                    ql[r] = q;
                }
            }
            float eline = 0.0f;
            float rg0 = 0.0f, rg1 = 0.0f, rg2 = 0.0f, dom_rg = -1.0f, dom_mn = 0.0f;
            uint32_t qd = 0;
            for (uint32_t q = 0; q < pc; ++q) {
                float mn = 3.0e38f, mx = -3.0e38f;
#pragma unroll
                for (int r = 0; r < K; ++r) if (ql[r] == q) { mn = fminf(mn, tl[r]); mx = fmaxf(mx, tl[r]); }
                mn = warp_min_f(mn); mx = warp_max_f(mx);
                if (!(mx > mn)) { mn = 0.0f; mx = 0.0f; }
                const float range = mx - mn;
                const float ir = range > 1e-6f*FX ? 1.0f/range : 0.0f;
#pragma unroll
                for (int r = 0; r < K; ++r) if (ql[r] == q) tl[r] = (tl[r] - mn)*ir;
                if (q == 0) rg0 = range; else if (q == 1) rg1 = range; else rg2 = range;
                if (range > dom_rg) { dom_rg = range; dom_mn = mn; qd = q; }
                const LineFit& lf = lines[base + q];
                eline += lf.resid;
                if (lane == 0) {
                    // (back from the weighted space: isw = 1/sw, exactly 1.0f for linear textures)
                    const float c0 = static_cast<float>(ctr.x), c1 = static_cast<float>(ctr.y), c2 = static_cast<float>(ctr.z), c3 = static_cast<float>(ctr.w);
                    slot.e0[q] = make_float4((c0 + (lf.m[0] + mn*lf.v[0])*isw.x)*ifx, (c1 + (lf.m[1] + mn*lf.v[1])*isw.y)*ifx,
                        (c2 + (lf.m[2] + mn*lf.v[2])*isw.z)*ifx, (c3 + (lf.m[3] + mn*lf.v[3])*isw.w)*ifx);
                    slot.e1[q] = make_float4((c0 + (lf.m[0] + mx*lf.v[0])*isw.x)*ifx, (c1 + (lf.m[1] + mx*lf.v[1])*isw.y)*ifx,
                        (c2 + (lf.m[2] + mx*lf.v[2])*isw.z)*ifx, (c3 + (lf.m[3] + mx*lf.v[3])*isw.w)*ifx);
                    slot.len2[q] = range*range*ifx*ifx;
                }
            }
            if (pc > 1 && dom_rg > 1e-6f*FX && !(tb.flags & 64u)) {
                // a subset whose line is negligible next to the longest one cannot lose much whatever its weights
                // are: its texels take the weights the longest line would give them, so a shared (decimated) weight
                // grid is not pulled about by the rounding noise of a flat background
                const LineFit& ld = lines[base + qd];
                const float dir = 1.0f/dom_rg;
#pragma unroll
                for (int r = 0; r < K; ++r) {
                    const uint32_t i = lane + 32u*r;
                    const float rq = ql[r] == 0 ? rg0 : (ql[r] == 1 ? rg1 : rg2);
                    if (i < T && ql[r] != qd && rq*8.0f < dom_rg) {
                        const int4 x = ws.v[i];
                        const float t = (sw.x*static_cast<float>(x.x - ctr.x) - ld.m[0])*ld.v[0] + (sw.y*static_cast<float>(x.y - ctr.y) - ld.m[1])*ld.v[1] +
                            (sw.z*static_cast<float>(x.z - ctr.z) - ld.m[2])*ld.v[2] + (sw.w*static_cast<float>(x.w - ctr.w) - ld.m[3])*ld.v[3];
                        tl[r] = fminf(fmaxf((t - dom_mn)*dir, 0.0f), 1.0f);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < K; ++r) {
                const uint32_t i = lane + 32u*r;
                if (i < T) ws.ta[s][i] = __float2half_rn(tl[r]);
            }
            if (lane == 0) { slot.e_line = eline*ifx*ifx; slot.len2b = 0.0f; }
            if (s >= 5) {
                // the dual channel runs on its own plane from its minimum to its maximum
                const int dc = static_cast<int>(s) - 5;
                const int lo = lo4[dc], hi = hi4[dc];
                const float ir2 = hi > lo ? 1.0f/static_cast<float>(hi - lo) : 0.0f;
                for (uint32_t i = lane; i < T; i += 32) {
                    const int4 x = ws.v[i];
                    const int xc = dc == 0 ? x.x : (dc == 1 ? x.y : (dc == 2 ? x.z : x.w));
                    ws.ta[s + 4][i] = __float2half_rn(static_cast<float>(xc - lo)*ir2);
                }
                if (lane == 0) {
                    float a4[4] = {slot.e0[0].x, slot.e0[0].y, slot.e0[0].z, slot.e0[0].w};
                    float b4[4] = {slot.e1[0].x, slot.e1[0].y, slot.e1[0].z, slot.e1[0].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (k == dc) { a4[k] = static_cast<float>(lo)*ifx; b4[k] = static_cast<float>(hi)*ifx; }
                    slot.e0[0] = make_float4(a4[0], a4[1], a4[2], a4[3]);
                    slot.e1[0] = make_float4(b4[0], b4[1], b4[2], b4[3]);
                    const float d = static_cast<float>(hi - lo)*ifx*(dc == 0 ? sw.x : (dc == 1 ? sw.y : (dc == 2 ? sw.z : sw.w)));
                    slot.len2b = d*d;
                }
            }
        }
        // ---- setup 8: the luminance slot (opaque blocks): weights along the gray axis, the chroma is its error floor
        if (active && !has_alpha) {
            int lmin = 1 << 30, lmax = -(1 << 30);
            float chroma = 0.0f;
            for (uint32_t i = lane; i < T; i += 32) {
                const int4 x = ws.v[i];
                const int l3 = x.x + x.y + x.z;                       // 3 * luminance, FX units
                lmin = min(lmin, l3); lmax = max(lmax, l3);
                const float l = static_cast<float>(l3)*(1.0f/3.0f);
                const float d0 = static_cast<float>(x.x) - l, d1 = static_cast<float>(x.y) - l, d2 = static_cast<float>(x.z) - l;
                chroma += d0*d0 + d1*d1 + d2*d2;
            }
            lmin = __reduce_min_sync(0xFFFFFFFFu, lmin); lmax = __reduce_max_sync(0xFFFFFFFFu, lmax);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) chroma += __shfl_xor_sync(0xFFFFFFFFu, chroma, o);
            const float ir = lmax > lmin ? 1.0f/static_cast<float>(lmax - lmin) : 0.0f;
            for (uint32_t i = lane; i < T; i += 32) {
                const int4 x = ws.v[i];
                ws.ta[kLumRow][i] = __float2half_rn(static_cast<float>(x.x + x.y + x.z - lmin)*ir);
            }
            if (lane == 0) {
                Slot3& sl = ws.slots[kLumSlot];
                const float l0 = static_cast<float>(lmin)*(ifx/3.0f), l1 = static_cast<float>(lmax)*(ifx/3.0f);
                sl.e0[0] = make_float4(l0, l0, l0, 255.0f); sl.e1[0] = make_float4(l1, l1, l1, 255.0f);
                sl.len2[0] = 3.0f*(l1 - l0)*(l1 - l0); sl.len2b = 0.0f;
                sl.e_line = chroma*ifx*ifx;
            }
            // the same with the two-subset (k = 0, 1) and three-subset (k = 2, 3) partitionings: a gray range per subset
            for (uint32_t k = 0; k < 4; ++k) {
                Slot3& sl = ws.slots[10 + k];
                const uint32_t npc = k < 2 ? 2u : 3u;
                const bool ok = ws.slots[1 + k].valid != 0 && !(tb.flags & 1u) && !HDR;
                if (lane == 0) { sl.valid = ok ? 1u : 0u; sl.pc = npc; sl.seed = ws.slots[1 + k].seed; sl.dual_ch = -1; }
                if (!ok) continue;
                int mn[3] = {1 << 30, 1 << 30, 1 << 30}, mx[3] = {-(1 << 30), -(1 << 30), -(1 << 30)};
                for (uint32_t i = lane; i < T; i += 32) {
                    const int4 x = ws.v[i];
                    const int l3 = x.x + x.y + x.z;
                    const uint32_t q = ws.part[k][i];
#pragma unroll
                    for (int qq = 0; qq < 3; ++qq) if (q == static_cast<uint32_t>(qq)) { mn[qq] = min(mn[qq], l3); mx[qq] = max(mx[qq], l3); }
                }
                float ir[3];
#pragma unroll
                for (int qq = 0; qq < 3; ++qq) {
                    mn[qq] = __reduce_min_sync(0xFFFFFFFFu, mn[qq]); mx[qq] = __reduce_max_sync(0xFFFFFFFFu, mx[qq]);
                    if (mx[qq] < mn[qq]) { mn[qq] = 0; mx[qq] = 0; }           // empty subset
                    ir[qq] = mx[qq] > mn[qq] ? 1.0f/static_cast<float>(mx[qq] - mn[qq]) : 0.0f;
                }
                // a subset whose range is negligible next to the widest one cannot lose much whatever its weights are:
                // give its texels the weights the widest subset's range would, so that a shared (decimated) weight
                // grid is not pulled about by the rounding noise of a flat background
                int rg[3] = {mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]};
                const int qd = rg[0] >= rg[1] && rg[0] >= rg[2] ? 0 : (rg[1] >= rg[2] ? 1 : 2);
                const int dom_lo = qd == 0 ? mn[0] : (qd == 1 ? mn[1] : mn[2]);
                const int dom_rg = qd == 0 ? rg[0] : (qd == 1 ? rg[1] : rg[2]);
                const float dom_ir = qd == 0 ? ir[0] : (qd == 1 ? ir[1] : ir[2]);
                const bool borrow = !(tb.flags & 64u);
                const uint32_t row = slot_row(10 + k);
                for (uint32_t i = lane; i < T; i += 32) {
                    const int4 x = ws.v[i];
                    const int l3 = x.x + x.y + x.z;
                    const uint32_t q = ws.part[k][i];
                    const int lo = q == 0 ? mn[0] : (q == 1 ? mn[1] : mn[2]);
                    const float r = q == 0 ? ir[0] : (q == 1 ? ir[1] : ir[2]);
                    const int rq = q == 0 ? rg[0] : (q == 1 ? rg[1] : rg[2]);
                    float t = static_cast<float>(l3 - lo)*r;
                    if (borrow && rq*8 < dom_rg) t = fminf(fmaxf(static_cast<float>(l3 - dom_lo)*dom_ir, 0.0f), 1.0f);
                    ws.ta[row][i] = __float2half_rn(t);
                }
                if (lane < npc) {
                    const float a0 = static_cast<float>(lane == 0 ? mn[0] : (lane == 1 ? mn[1] : mn[2]))*(ifx/3.0f);
                    const float b0 = static_cast<float>(lane == 0 ? mx[0] : (lane == 1 ? mx[1] : mx[2]))*(ifx/3.0f);
                    sl.e0[lane] = make_float4(a0, a0, a0, 255.0f); sl.e1[lane] = make_float4(b0, b0, b0, 255.0f);
                    sl.len2[lane] = 3.0f*(b0 - a0)*(b0 - a0);
                }
                if (lane == 0) { sl.len2b = 0.0f; sl.e_line = chroma*ifx*ifx; }
            }
        } else if (active) {
            // ---- setup 8a: blocks with alpha: LUMINANCE + ALPHA end points (CEM 4: L0 L1 A0 A1) for one subset (slot 9) and
            //      the two-subset partitionings (slots 10, 11; the three-subset rows belong to the alpha dual-plane slot
            //      here).  Gray content under a varying alpha -- smoke, glyphs, shadows -- is what astcenc spends this mode
            //      on.  The weights run along the principal axis of the subset in the plane (sqrt(3) L, A): a unit of
            //      luminance error is paid on three channels.
            if (lane < 4) ws.slots[10 + lane].valid = 0;
            float chroma = 0.0f;
            for (uint32_t i = lane; i < T; i += 32) {
                const int4 x = ws.v[i];
                const float l = static_cast<float>(x.x + x.y + x.z)*(1.0f/3.0f);
                const float d0 = static_cast<float>(x.x) - l, d1 = static_cast<float>(x.y) - l, d2 = static_cast<float>(x.z) - l;
                chroma += d0*d0 + d1*d1 + d2*d2;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) chroma += __shfl_xor_sync(0xFFFFFFFFu, chroma, o);
            // lane j < 5: axis of set j (0 = the whole block, 1..4 = the subsets of slots 1, 2) from its RGBA moments
            float* ax = &ws.g[0][0];            // [set][8]: mean u, mean a, cos, sin, residual      (phase-2 scratch, free here)
            __syncwarp();
            if (lane < 5) {
                Mom mo = moms[lane == 0 ? 10u : lane - 1u];
                // (this axis lives in the plain (sqrt(3) L, A) plane: undo the channel scaling of the moments)
#pragma unroll
                for (int k = 1; k < 15; ++k) (&mo.n)[k] *= 1.0f/mom_scale(k, sw);
                const float n = fmaxf(mo.n, 1.0f), in = 1.0f/n;
                const float su = mo.s[0] + mo.s[1] + mo.s[2], sa = mo.s[3];
                const float suu = mo.p[0] + mo.p[4] + mo.p[7] + 2.0f*(mo.p[1] + mo.p[2] + mo.p[5]) - su*su*in;
                const float sua = mo.p[3] + mo.p[6] + mo.p[8] - su*sa*in;
                const float saa = mo.p[9] - sa*sa*in;
                const float xx = suu*(1.0f/3.0f), xa = sua*0.57735027f;
                const float half_d = 0.5f*(xx - saa), r = sqrtf(half_d*half_d + xa*xa);
                // eigenvector of the larger eigenvalue of [[xx, xa], [xa, saa]]
                float cx = half_d + r, cy = xa;
                if (cx*cx + cy*cy < 1e-12f*(xx + saa)*(xx + saa) + 1e-20f) { cx = xa; cy = r - half_d; }
                float nn = cx*cx + cy*cy;
                if (nn > 0.0f) { nn = rsqrtf(nn); cx *= nn; cy *= nn; } else { cx = 1.0f; cy = 0.0f; }
                ax[lane*8u + 0u] = su*in + static_cast<float>(ctr.x + ctr.y + ctr.z);
                ax[lane*8u + 1u] = sa*in + static_cast<float>(ctr.w);
                ax[lane*8u + 2u] = cx; ax[lane*8u + 3u] = cy;
                ax[lane*8u + 4u] = fmaxf(0.5f*(xx + saa) - r, 0.0f);
            }
            __syncwarp();
            for (uint32_t job = 0; job < 3; ++job) {           // slot 9, 10, 11
                const uint32_t sidx = job == 0 ? static_cast<uint32_t>(kLumSlot) : 9u + job;
                Slot3& sl = ws.slots[sidx];
                const bool ok = (job == 0 ? sl.valid != 0 : ws.slots[job].valid != 0) && !(tb.flags & 1u) && !HDR;
                if (job > 0 && lane == 0) { sl.valid = ok ? 1u : 0u; sl.pc = 2u; sl.seed = ws.slots[job].seed; sl.dual_ch = -1; }
                if (!ok) continue;
                const uint32_t row = job == 0 ? static_cast<uint32_t>(kLumRow) : slot_row(sidx);
                const uint32_t set0 = job == 0 ? 0u : 1u + 2u*(job - 1u);
                float tl[K]; uint32_t ql[K];
                float mn[2] = {3.0e38f, 3.0e38f}, mx[2] = {-3.0e38f, -3.0e38f};
#pragma unroll
                for (int r = 0; r < K; ++r) {
                    const uint32_t i = lane + 32u*r;
                    tl[r] = 0.0f; ql[r] = 0u;
                    if (i < T) {
                        const uint32_t q = job == 0 ? 0u : ws.part[job - 1u][i];
                        const float* a5 = ax + (set0 + q)*8u;
                        const int4 x = ws.v[i];
                        tl[r] = (static_cast<float>(x.x + x.y + x.z) - a5[0])*0.57735027f*a5[2] + (static_cast<float>(x.w) - a5[1])*a5[3];
                        ql[r] = q;
                        if (q == 0u) { mn[0] = fminf(mn[0], tl[r]); mx[0] = fmaxf(mx[0], tl[r]); }
                        else { mn[1] = fminf(mn[1], tl[r]); mx[1] = fmaxf(mx[1], tl[r]); }
                    }
                }
                float eline = chroma;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (job == 0 && q == 1) continue;
                    float lo = warp_min_f(mn[q]), hi = warp_max_f(mx[q]);
                    if (!(hi > lo)) { lo = 0.0f; hi = 0.0f; }
                    const float range = hi - lo, irg = range > 1e-6f*FX ? 1.0f/range : 0.0f;
#pragma unroll
                    for (int r = 0; r < K; ++r) if (ql[r] == static_cast<uint32_t>(q) && lane + 32u*r < T) tl[r] = (tl[r] - lo)*irg;
                    const float* a5 = ax + (set0 + static_cast<uint32_t>(q))*8u;
                    eline += a5[4];
                    if (lane == 0) {
                        // back from (sqrt(3) L, A) to luminance and alpha, 0..255
                        const float l0 = (a5[0] + lo*a5[2]*1.7320508f)*(ifx/3.0f), l1 = (a5[0] + hi*a5[2]*1.7320508f)*(ifx/3.0f);
                        const float a0 = (a5[1] + lo*a5[3])*ifx, a1 = (a5[1] + hi*a5[3])*ifx;
                        sl.e0[q] = make_float4(l0, l0, l0, a0); sl.e1[q] = make_float4(l1, l1, l1, a1);
                        sl.len2[q] = range*range*ifx*ifx;
                    }
                }
#pragma unroll
                for (int r = 0; r < K; ++r) {
                    const uint32_t i = lane + 32u*r;
                    if (i < T) ws.ta[row][i] = __float2half_rn(fminf(fmaxf(tl[r], 0.0f), 1.0f));
                }
                if (lane == 0) { sl.len2b = 0.0f; sl.e_line = eline*ifx*ifx; }
            }
            __syncwarp();
        }
        PHASE_SYNC();

        // ---- phase 1a: decimation loss D[slot plane][grid] on the tensor cores
        uint32_t a[KS][4];
        load_a<KS>(ws, a, lane);
        float lw[NT][2], lw2[NT][2];
        float scale0 = 0.0f, scale1 = 0.0f, pscale0 = 0.0f, pscale1 = 0.0f;
        const float* const colen = reinterpret_cast<const float*>(ctx.blob + tb.t3.off_colenergy);
        const uint8_t* const dec = ctx.blob + tb.t3.off_dec_list;
        if (active) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const uint32_t i = nt*8 + 2*tq + e;
                    float wgt = i < T ? 1.0f : 0.0f;
                    if (gq >= 1 && gq <= 4 && i < T) wgt = ws.slots[gq].len2[ws.part[gq - 1][i]];
                    lw[nt][e] = wgt;
                    // upper fragment half: rows 14, 15 are the two-subset luminance slots 10, 11 (partitionings 0, 1)
                    float wgt2 = i < T ? 1.0f : 0.0f;
                    if (gq >= 6 && i < T) wgt2 = ws.slots[gq + 4].len2[ws.part[gq - 6][i]];
                    // rows 8 / 12 belong to the three-subset luminance slots 12 / 13 when the block is opaque
                    if (gq == 0 && i < T && ws.slots[12].valid) wgt2 = ws.slots[12].len2[ws.part[2][i]];
                    if (gq == 4 && i < T && ws.slots[13].valid) wgt2 = ws.slots[13].len2[ws.part[3][i]];
                    lw2[nt][e] = wgt2;
                }
            scale0 = gq == 0 ? ws.slots[0].len2[0] : (gq <= 4 ? 1.0f : ws.slots[gq].len2[0]);
            scale1 = gq == 0 ? (ws.slots[12].valid ? 1.0f : ws.slots[8].len2[0]) :
                (gq == 4 ? (ws.slots[13].valid ? 1.0f : ws.slots[8].len2b) :
                (gq < 4 ? ws.slots[gq + 4].len2b : (gq == 5 ? ws.slots[kLumSlot].len2[0] : 1.0f)));
            // what one unit of clamped overshoot of a grid weight costs this row: its line length (multi-subset rows:
            // the mean over the subsets, the grid is shared)
            auto mean_len2 = [&](uint32_t sl) {
                const Slot3& q = ws.slots[sl];
                return q.pc > 2 ? (q.len2[0] + q.len2[1] + q.len2[2])*(1.0f/3.0f) : (q.len2[0] + q.len2[1])*0.5f;
            };
            pscale0 = gq >= 1 && gq <= 4 ? mean_len2(gq) : scale0;
            pscale1 = gq == 0 && ws.slots[12].valid ? mean_len2(12) : (gq == 4 && ws.slots[13].valid ? mean_len2(13) :
                (gq >= 6 ? mean_len2(gq + 4) : scale1));
            // full-resolution grids lose nothing
            for (uint32_t g = lane; g < G; g += 32)
                if (__ldg(reinterpret_cast<const uint32_t*>(ctx.blob + tb.t3.off_rfrag_idx) + g) == 0u)
                    for (uint32_t r = 0; r < 16; ++r) ws.u.est.D[r][g] = 0.0f;
        }
        if constexpr (STAGE) {
            // The R and M fragments of a weight grid are the same for every block: ONE thread of the CTA fetches them into
            // shared memory with two bulk copies (cp.async.bulk, the 1-D TMA path) per grid, double buffered on two mbarriers,
            // and all the warps -- which walk the grids in step anyway -- feed their MMAs from there, instead of every
            // warp pulling its own copy of the fragments through L1 (long-scoreboard stalls on the MMAs were 6 % of the
            // kernel's samples with a one-tile register prefetch).
            auto issue = [&](uint32_t k, uint32_t buf) {
                const uint32_t g = __ldg(dec + k);
                const uint32_t mbytes = ((tab_u8(ctx, ctx.tab.off_grids + g*4u + 2u) + 7u) >> 3)*(KS*256u);
                uint8_t* dstb = stage + buf*2u*kTileBytes;
                mbar_expect_tx(&stage_bar[buf], kTileBytes + mbytes);
                bulk_load(dstb, ctx.blob + tb.t3.off_rstream + k*kTileBytes, kTileBytes, &stage_bar[buf]);
                bulk_load(dstb + kTileBytes, ctx.blob + __ldg(reinterpret_cast<const uint32_t*>(ctx.blob + tb.t3.off_mfrag_idx) + g), mbytes, &stage_bar[buf]);
            };
            const uint32_t n_dec = tb.t3.n_dec;
            if (threadIdx.x == 0 && n_dec) issue(0, staged & 1u);
#pragma unroll 1
            for (uint32_t k = 0; k < n_dec; ++k) {
                const uint32_t buf = staged & 1u;
                // the other buffer was last read one trip ago: every warp has passed that trip's closing barrier
                if (threadIdx.x == 0 && k + 1u < n_dec) issue(k + 1u, buf ^ 1u);
                mbar_wait(&stage_bar[buf], (staged >> 1) & 1u);
                if (active) {
                    const uint32_t g = __ldg(dec + k);
                    const uint2* rf = reinterpret_cast<const uint2*>(stage + buf*2u*kTileBytes) + lane;
                    const uint2* mfs = rf + kTileBytes/8u;
                    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) mma16816(c, a[ks], rf[(nt*KS + ks)*32]);
                        acc0 += lw[nt][0]*c[0]*c[0] + lw[nt][1]*c[1]*c[1];
                        acc1 += lw2[nt][0]*c[2]*c[2] + lw2[nt][1]*c[3]*c[3];
                    }
                    float pen0 = 0.0f, pen1 = 0.0f;
                    if (!(tb.flags & 16u)) {
                        const uint32_t ntw = (tab_u8(ctx, ctx.tab.off_grids + g*4u + 2u) + 7u) >> 3;      // GridInfo::nw in tiles of 8
#pragma unroll 1
                        for (uint32_t nt = 0; nt < ntw; ++nt) {
                            float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                            for (int ks = 0; ks < KS; ++ks) mma16816(c, a[ks], mfs[(nt*KS + ks)*32u]);
                            const float2 ce = __ldg(reinterpret_cast<const float2*>(colen + g*64u + nt*8u + 2u*tq));
                            float o = c[0] - fminf(fmaxf(c[0], 0.0f), 1.0f); pen0 += ce.x*o*o;
                            o = c[1] - fminf(fmaxf(c[1], 0.0f), 1.0f); pen0 += ce.y*o*o;
                            o = c[2] - fminf(fmaxf(c[2], 0.0f), 1.0f); pen1 += ce.x*o*o;
                            o = c[3] - fminf(fmaxf(c[3], 0.0f), 1.0f); pen1 += ce.y*o*o;
                        }
                    }
                    acc0 = acc0*scale0 + pen0*pscale0; acc1 = acc1*scale1 + pen1*pscale1;
                    acc0 += __shfl_xor_sync(0xFFFFFFFFu, acc0, 1); acc0 += __shfl_xor_sync(0xFFFFFFFFu, acc0, 2);
                    acc1 += __shfl_xor_sync(0xFFFFFFFFu, acc1, 1); acc1 += __shfl_xor_sync(0xFFFFFFFFu, acc1, 2);
                    if (tq == 0) {
                        ws.u.est.D[gq][g] = acc0;
                        ws.u.est.D[gq + 8][g] = acc1;
                    }
                }
                __syncthreads();
                ++staged;
            }
        } else if (active) {
            // the decimated grids' R fragments are one contiguous stream: walk it with a one-tile prefetch
            const uint2* frag = reinterpret_cast<const uint2*>(ctx.blob + tb.t3.off_rstream) + lane;
            uint2 bcur[KS];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) bcur[ks] = __ldg(frag + ks*32);
#pragma unroll 1
            for (uint32_t k = 0; k < tb.t3.n_dec; ++k) {
                const uint32_t g = __ldg(dec + k);
                float acc0 = 0.0f, acc1 = 0.0f;
                // the M fragments of this grid (second GEMM below): the dependent loads g -> fragment offset -> first tile
                // are issued here, ahead of the R tiles' MMAs
                const bool with_pen = !(tb.flags & 16u);
                uint32_t ntw = 0;
                const uint2* mf = nullptr;
                uint2 mcur[KS];
                if (with_pen) {
                    ntw = (tab_u8(ctx, ctx.tab.off_grids + g*4u + 2u) + 7u) >> 3;          // GridInfo::nw in tiles of 8
                    mf = reinterpret_cast<const uint2*>(ctx.blob + __ldg(reinterpret_cast<const uint32_t*>(ctx.blob + tb.t3.off_mfrag_idx) + g)) + lane;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) mcur[ks] = __ldg(mf + ks*32u);
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    frag += KS*32;
                    uint2 bnext[KS];
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) bnext[ks] = __ldg(frag + ks*32);     // the table ends with a pad tile
                    float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) mma16816(c, a[ks], bcur[ks]);
                    acc0 += lw[nt][0]*c[0]*c[0] + lw[nt][1]*c[1]*c[1];
                    acc1 += lw2[nt][0]*c[2]*c[2] + lw2[nt][1]*c[3]*c[3];
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) bcur[ks] = bnext[ks];
                }
                // + what clamping the least-squares grid weights into [0, 1] costs (see astc3_tables.hpp): M t on the
                // tensor cores, overshoot^2 weighted by the weight's column energy
                float pen0 = 0.0f, pen1 = 0.0f;
                if (with_pen) {
                    // (one-tile prefetch, as for R above)
#pragma unroll 1
                    for (uint32_t nt = 0; nt < ntw; ++nt) {
                        uint2 mnext[KS];
                        const uint32_t ntn = nt + 1u < ntw ? nt + 1u : nt;
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) mnext[ks] = __ldg(mf + (ntn*KS + ks)*32u);
                        float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) mma16816(c, a[ks], mcur[ks]);
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) mcur[ks] = mnext[ks];
                        const float2 ce = __ldg(reinterpret_cast<const float2*>(colen + g*64u + nt*8u + 2u*tq));
                        float o = c[0] - fminf(fmaxf(c[0], 0.0f), 1.0f); pen0 += ce.x*o*o;
                        o = c[1] - fminf(fmaxf(c[1], 0.0f), 1.0f); pen0 += ce.y*o*o;
                        o = c[2] - fminf(fmaxf(c[2], 0.0f), 1.0f); pen1 += ce.x*o*o;
                        o = c[3] - fminf(fmaxf(c[3], 0.0f), 1.0f); pen1 += ce.y*o*o;
                    }
                }
                acc0 = acc0*scale0 + pen0*pscale0; acc1 = acc1*scale1 + pen1*pscale1;
                acc0 += __shfl_xor_sync(0xFFFFFFFFu, acc0, 1); acc0 += __shfl_xor_sync(0xFFFFFFFFu, acc0, 2);
                acc1 += __shfl_xor_sync(0xFFFFFFFFu, acc1, 1); acc1 += __shfl_xor_sync(0xFFFFFFFFu, acc1, 2);
                if (tq == 0) {
                    ws.u.est.D[gq][g] = acc0;
                    ws.u.est.D[gq + 8][g] = acc1;
                }
            }
        }
        // ---- phase 1b: S of the multi-subset slots (lane = grid)
        for (uint32_t g = lane; active && g < G; g += 32) {
            const float* kap = reinterpret_cast<const float*>(ctx.blob + tb.t3.off_kappa) + g*kMaxTexels3;
            float s1 = 0.0f, s2 = 0.0f, s3 = 0.0f, s4 = 0.0f, s5 = 0.0f, s6 = 0.0f, s7 = 0.0f, s8 = 0.0f;
            for (uint32_t i = 0; i < T; ++i) {
                const float k = __ldg(kap + i);
                const uint32_t p0 = ws.part[0][i], p1 = ws.part[1][i], p2 = ws.part[2][i], p3 = ws.part[3][i];
                s1 += k*ws.slots[1].len2[p0]; s2 += k*ws.slots[2].len2[p1];
                s3 += k*ws.slots[3].len2[p2]; s4 += k*ws.slots[4].len2[p3];
                s5 += k*ws.slots[10].len2[p0]; s6 += k*ws.slots[11].len2[p1];
                s7 += k*ws.slots[12].len2[p2]; s8 += k*ws.slots[13].len2[p3];
            }
            ws.u.est.Sm[0][g] = s1; ws.u.est.Sm[1][g] = s2; ws.u.est.Sm[2][g] = s3; ws.u.est.Sm[3][g] = s4;
            ws.u.est.Sm[4][g] = s5; ws.u.est.Sm[5][g] = s6; ws.u.est.Sm[6][g] = s7; ws.u.est.Sm[7][g] = s8;
        }
        // ---- phase 1b': measured weight-quantisation loss of every slot at the coarse levels (lane = texel):
        //      qn[s][L] = sum_i len2_i (t_i - Q_L(t_i))^2 / sum_i len2_i, with the end point refit gain folded in
        {
            for (uint32_t s = 0; s < kSlots3; ++s) {
                if (CFX_ASTC3_SLOTSYNC) PHASE_SYNC();
                if (!active) continue;
                const Slot3& slot = ws.slots[s];
                if (!slot.valid) continue;
                const uint32_t row0 = slot_row(s);
                const bool dual = slot.dual_ch >= 0;
                const float l2b = dual ? slot.len2b : 0.0f;
                float e[kDataLevels];
#pragma unroll
                for (int L = 0; L < kDataLevels; ++L) e[L] = 0.0f;
                float wsum = 0.0f;
#pragma unroll
                for (int r = 0; r < K; ++r) {
                    const uint32_t i = lane + 32u*r;
                    if (i >= T) continue;
                    const float t = __half2float(ws.ta[row0][i]);
                    const float w = slot.pc > 1 ? slot.len2[ws.part[slot_part(s)][i]] : slot.len2[0];
                    const float t2 = dual ? __half2float(ws.ta[s + 4][i]) : 0.0f;
                    wsum += w + l2b;
#pragma unroll
                    for (int L = 0; L < kDataLevels; ++L) {
                        constexpr int kLv[kDataLevels] = {2, 3, 4, 5, 6, 8};
                        const float nm1 = static_cast<float>(kLv[L] - 1), inm1 = 1.0f/nm1;   // compile-time constants
                        const float d = fmaf(-rintf(t*nm1), inm1, t);
                        if (dual) {
                            const float d2 = fmaf(-rintf(t2*nm1), inm1, t2);
                            e[L] += w*d*d + l2b*d2*d2;
                        } else e[L] = __fadd_rn(e[L], __fmul_rn(__fmul_rn(w, d), d));      // (the same roundings as the line above with l2b = 0)
                    }
                }
                // one integer reduction per level (the sums are scaled into 2^26)
                const float wtot = static_cast<float>(redux_add(__float2int_rn(wsum*64.0f)))*(1.0f/64.0f);
                const float sc = wtot > 0.0f ? 67108864.0f/wtot : 0.0f;
#pragma unroll
                for (int L = 0; L < kDataLevels; ++L) {
                    constexpr int kLv[kDataLevels] = {2, 3, 4, 5, 6, 8};
                    const float inm1 = 1.0f/static_cast<float>(kLv[L] - 1);
                    const int tot = redux_add(__float2int_rn(e[L]*sc));
                    if (lane == static_cast<uint32_t>(L)) ws.u.est.qn[s][L] = static_cast<float>(tot)*(1.0f/67108864.0f)*(1.0f - 0.75f*inm1);
                }
            }
        }
        PHASE_SYNC();
        // ---- phase 1c: estimate every (slot, mode); each lane keeps its three best
        float be0 = 3.0e38f, be1 = 3.0e38f, be2 = 3.0e38f;
        uint32_t bc0 = 0, bc1 = 0, bc2 = 0;
        {
            const float tn = kColor*static_cast<float>(T*(has_alpha ? 4u : 3u));
            const float* ksum = reinterpret_cast<const float*>(ctx.blob + tb.t3.off_ksum);
            float* gb = &ws.g[0][0];           // per grid: floor + decimation loss   (phase 2 scratch, free until then)
            float* gs = &ws.g[1][0];           // per grid: weight-quantisation scale
            for (uint32_t s = 0; s < kSlotsAll; ++s) {
                if (CFX_ASTC3_SLOTSYNC) PHASE_SYNC();
                if (!active) continue;
                const Slot3& slot = ws.slots[slot_base(s)];
                const bool virt = slot_is_scale(s);
                if (virt ? !(ws.scale_valid[s - kScaleSlot] && slot.valid) : !slot.valid) continue;
                const uint32_t type = slot_kind(s, has_alpha);
                const uint32_t drow = slot_row(s);
                const uint4* list = reinterpret_cast<const uint4*>(ctx.blob + tb.t3.off_est[has_alpha ? 1 : 0][type]);
                const uint32_t count = tb.t3.n_est[has_alpha ? 1 : 0][type];
                const float base = kLine*(virt ? ws.scale_eline[s - kScaleSlot] + (s < static_cast<uint32_t>(kMixSlot) && has_alpha ? slot.e_line : 0.0f) : slot.e_line);
                // every estimate of this slot is at least its floor: when n_exact lanes already hold something better, none
                // of them can be among the n_exact candidates phase 2 looks at
                if (static_cast<uint32_t>(__popc(__ballot_sync(0xFFFFFFFFu, be0 < base))) >= n_exact CFX_ASTC3_TUNE_KEEP_ALL) continue;
                const float l2sum = slot.len2[0] + slot.len2b;
                // per-grid terms of this slot (lane = grid)
                __syncwarp();
                for (uint32_t g = lane; g < G; g += 32) {
                    float dsum = ws.u.est.D[drow][g], ssum;
                    if (type == 3) { dsum += ws.u.est.D[s + 4][g]; ssum = l2sum*__ldg(ksum + g); }
                    else if (type == 0 || type == 4 || type == 7) ssum = l2sum*__ldg(ksum + g);
                    else ssum = ws.u.est.Sm[type >= 10 ? s - 19 : (type >= 8 ? s - 15 : (type >= 5 ? s - 6 : s - 1))][g];
                    gb[g] = base + kDec*dsum; gs[g] = kQuant*ssum;
                }
                __syncwarp();
                const float* qn = ws.u.est.qn[slot_base(s)];
#pragma unroll 2
                for (uint32_t e = lane; e < count; e += 32) {
                    const uint4 q = __ldg(list + e);
                    const uint32_t g = (q.z >> 16) & 0xFFu;
#ifdef CFX_ASTC3_TUNE
                    {   // developer knobs: only the given slot / weight level / number of grid weights
                        const ModeInfo dm = tab_mode(ctx, q.z & 0xFFFFu);
                        if ((tb.dbg_slot >= 0 && static_cast<int>(s) != tb.dbg_slot) || (tb.dbg_level >= 0 && static_cast<int>(kWqN[dm.level]) != tb.dbg_level) ||
                            (tb.dbg_nw >= 0 && static_cast<int>(dm.nw) != tb.dbg_nw)) continue;
                    }
#endif
                    const float a = (tb.flags & 2u) ? 0.0f : __half2float(__ushort_as_half(static_cast<unsigned short>(q.w >> 16)));
                    const float qv = fmaf(a, qn[q.w & 0xFFu], __uint_as_float(q.x));
                    const float est = fmaf(gs[g], qv, fmaf(tn, __uint_as_float(q.y), gb[g]));
                    const uint32_t code = (s << 16) | (q.z & 0xFFFFu);
#ifdef CFX_ASTC3_TUNE
                    if (tb.flags & 512u)     // developer: the terms of every estimate (floor, decimation, weight quantisation, colour)
                        printf("EST blk %u slot %u mode %u base %.1f dec %.1f quant %.1f colour %.1f\n", blk, s, q.z & 0xFFFFu, base, gb[g] - base,
                            gs[g]*qv, tn*__uint_as_float(q.y));
#endif
                    if (est < be2) {
                        if (est < be1) {
                            be2 = be1; bc2 = bc1;
                            if (est < be0) { be1 = be0; bc1 = bc0; be0 = est; bc0 = code; }
                            else { be1 = est; bc1 = code; }
                        } else { be2 = est; bc2 = code; }
                    }
                }
            }
        }
        PHASE_SYNC();
        // ---- phase 2: exact evaluation of the n_exact best estimates, then refinement rounds on the winner
        //      (re-project the texels on its quantised end points, re-decimate, re-solve) -- one loop, so that the
        //      evaluation code exists once
        float best_err = 3.0e38f;
        uint32_t best_code = 0, best_cl = 0;
        const float stop_db = fmaxf(95.0f - 35.0f*log10f(static_cast<float>(T)), 70.0f - 19.0f*log10f(static_cast<float>(T))) + 12.0f;
        const float stop_err = 65025.0f*exp10f(-0.1f*stop_db)*static_cast<float>(T*nch)*fx2;
        bool refining = false;
        uint32_t n = 0, rounds = 0, fails = 0;
#ifdef CFX_ASTC3_TUNE
        float dbg_est = 0.0f;
#endif
        // every trip of the loop starts at a CTA-wide barrier as well (LOCK): the body is ~2500 instructions, the warps
        // take different branches of it and run different numbers of trips, and the instruction cache only keeps up while
        // they stay together; a warp that is done keeps arriving until the whole CTA is
        bool running = active;
#pragma unroll 1
        while (LOCK && CFX_ASTC3_LOOPSYNC ? __syncthreads_or(running ? 1 : 0) != 0 : running) {
            if (!running) continue;
            uint32_t code = 0, cl = 0;
            int row0 = 0, row1 = -1;
            bool realign = false;
            if (!refining) {
                bool done = n >= n_exact || best_err <= stop_err;
#ifdef CFX_ASTC3_TUNE
                if (tb.flags & 256u) done = false;
#endif
                if (!done) {
                    const uint32_t kmin = __reduce_min_sync(0xFFFFFFFFu, (__float_as_uint(be0) & ~31u) | lane);
                    const uint32_t wl = kmin & 31u;
                    const float est = __shfl_sync(0xFFFFFFFFu, be0, wl);
                    code = __shfl_sync(0xFFFFFFFFu, bc0, wl);
                    if (lane == wl) { be0 = be1; bc0 = bc1; be1 = be2; bc1 = bc2; be2 = 3.0e38f; }
                    // estimates come out sorted: stop when nothing later can win
                    done = est >= 3.0e38f || (n_exact > 8u ? 0.6f : 0.8f)*est*fx2 > best_err;
#ifdef CFX_ASTC3_TUNE
                    if (tb.flags & 256u) done = est >= 3.0e38f;      // developer: evaluate everything that was kept
                    dbg_est = est;
#endif
                    ++n;
                }
                if (done) refining = true;
                else {
                    const uint32_t s = code >> 16;
                    cl = __ldg(ctx.blob + tb.t3.off_modecl + ((has_alpha ? static_cast<uint32_t>(kSlotTypes3) : 0u) + slot_kind(s, has_alpha))*tb.t3.n_modes + (code & 0xFFFFu));
                    row0 = static_cast<int>(slot_row(s));
                    row1 = ws.slots[slot_base(s)].dual_ch >= 0 ? static_cast<int>(s) + 4 : -1;
                }
            }
            if (refining) {
                // refinement alternates a continuous step (re-project on the quantised end points, re-decimate, re-solve)
                // with a discrete one (realign_weights); it ends after `refine` pairs or two failures in a row
                if (rounds >= 2u*refine || fails >= 2u || best_err <= 0.0f || best_err >= 3.0e38f) { running = false; continue; }
                realign = !HDR && (rounds & 1u) != 0u && !(tb.flags & 8u);
                const bool skip = (rounds & 1u) != 0u && !realign;
                ++rounds;
                if (skip) { ++fails; continue; }
                code = best_code; cl = best_cl;
                const uint32_t bs = code >> 16;
                const Slot3& bslot = ws.slots[slot_base(bs)];
                const int dc = bslot.dual_ch;
                const float escale = HDR ? 1.0f/16.0f : 1.0f;     // HDR end points are 12-bit, texels 8-bit-like
                const uint8_t* parts = ws.part[slot_part(bs)];
                for (uint32_t i = lane; !realign && i < T; i += 32) {
                    const int* e = ws.best_ep + (bslot.pc > 1 ? parts[i] : 0u)*8u;
                    const int4 xi = ws.v[i];
                    const float xs[4] = {static_cast<float>(xi.x)*ifx, static_cast<float>(xi.y)*ifx, static_cast<float>(xi.z)*ifx,
                        static_cast<float>(xi.w)*ifx};
                    float num0 = 0, den0 = 0, num1 = 0, den1 = 0;
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        if (c4 == 3 && !has_alpha) continue;
                        const float es = c4 < 3 ? escale : (HDR ? kHdrAlphaScale : 1.0f);
                        const float a0 = static_cast<float>(e[c4])*es, d = static_cast<float>(e[4 + c4])*es - a0;
                        const float w4 = c4 == 0 ? sw.x*sw.x : (c4 == 1 ? sw.y*sw.y : (c4 == 2 ? sw.z*sw.z : sw.w*sw.w));
                        if (c4 == dc) { num1 += w4*((xs[c4] - a0)*d); den1 += w4*(d*d); } else { num0 += w4*((xs[c4] - a0)*d); den0 += w4*(d*d); }
                    }
                    ws.ta[0][i] = __float2half_rn(den0 > 0.0f ? fminf(fmaxf(num0/den0, 0.0f), 1.0f) : 0.0f);
                    ws.ta[1][i] = __float2half_rn(den1 > 0.0f ? fminf(fmaxf(num1/den1, 0.0f), 1.0f) : 0.0f);
                }
                __syncwarp();
                if (!realign) load_a<KS>(ws, a, lane);             // the candidates' rows are not needed any more
                row0 = 0; row1 = dc >= 0 ? 1 : -1;
            }
            const uint32_t s = code >> 16;
            const ModeInfo m = tab_mode(ctx, code & 0xFFFFu);
            if (realign) realign_weights(ctx, ws, s, m, has_alpha, lane, tb.cw);
            else decimate_mma<KS>(tb, ws, a, m.grid, m.nw, row0, row1, lane);
            if (!realign && NT > 8 && ws.slots[slot_base(s)].pc > 1 && m.nw < T && !(tb.flags & 32u)) reweight_decimation(ctx, ws, s, m.grid, m.nw, static_cast<uint32_t>(row0), lane);
            if (realign) {
                // the realigned weights were chosen against the winner's end points: they are measured with exactly those
                // (each accepted move lowered the exact error, so this can only improve on the winner); the next
                // continuous round re-solves the end points for them
                const uint32_t rpc = ws.slots[slot_base(s)].pc;
                if (lane < rpc*8u) { ws.ep[lane] = ws.best_ep[lane]; ws.epv[lane] = ws.best_epv[lane]; }
                if (lane < 3u) ws.sc[lane] = ws.best_sc[lane];
                if (lane == 0u) ws.contr = ws.best_contr;
                __syncwarp();
            }
            const float err = evaluate3<K, HDR>(ctx, ws, s, m, cl, has_alpha, lane, !realign, !realign, tb.cw);
#ifdef CFX_ASTC3_TUNE
            if ((tb.flags & 256u) && lane == 0 && !refining)
                printf("CAND blk %u slot %u mode %u nw %u level %u cl %u est %.1f exact %.1f\n", blk, s, code & 0xFFFFu, static_cast<uint32_t>(m.nw),
                    kWqN[m.level], cl, dbg_est, err/fx2);
#endif
#ifdef CFX_ASTC3_TUNE
            if ((tb.flags & 256u) && lane == 0 && refining)
                printf("REFINE blk %u round %u realign %d err %.1f best %.1f contr %u\n", blk, rounds, realign ? 1 : 0, err/fx2, best_err/fx2, ws.contr);
#endif
            if (err < best_err) {
                best_err = err; best_code = code; best_cl = cl; fails = 0;
                keep_best3(ws, m.nw, ws.slots[slot_base(s)].dual_ch >= 0 ? 2u : 1u, ws.slots[slot_base(s)].pc, lane);
            } else if (refining) ++fails;
            __syncwarp();
        }
        const uint32_t bs = best_code >> 16;
        const Slot3& bslot = ws.slots[slot_base(bs)];
        const ModeInfo bm = tab_mode(ctx, best_code & 0xFFFFu);
        PHASE_SYNC();       // (measured: without this barrier the warps drift apart and instruction fetch stalls cost 25 %)
        if (active) {
            SlotView sv; sv.pc = bslot.pc; sv.seed = bslot.seed; sv.dual_ch = bslot.dual_ch;
            uint8_t cems[4] = {0, 0, 0, 0};
            const bool virt = slot_is_scale(bs);
            for (uint32_t q = 0; q < bslot.pc; ++q) {
                const bool cheap = bs >= static_cast<uint32_t>(kMixSlot) ? ((ws.mix_mask[bs - kMixSlot] >> q) & 1u) != 0u : virt;
                cems[q] = static_cast<uint8_t>(!cheap ? (has_alpha ? 12 : 8) : (bs >= static_cast<uint32_t>(kMixSlot) && has_alpha && ws.mix_kind[bs - kMixSlot] == 0u ? 8 : (has_alpha ? 10 : 6)));
            }
            // (the candidate scratch ws.sk / ws.su is dead by now: it takes the encoded colour values and weights)
            const uint4 packed = pack_block_warp(ctx, sv, bm, best_cl, ws.best_ep, has_alpha, ws.best_sk, slot_is_lum(bs), HDR ? ws.best_epv : nullptr,
                virt ? cems : nullptr, ws.best_sc, HDR ? 0u : ws.best_contr, ws.best_epv, ws.sk, ws.su, lane);
            if (lane == 0) *dst = packed;
        }
    }
}

namespace {

template <int NT, int KS, int W, int CTAS, bool LOCK>
int launch_cfg(const EncodeParams& p, const Tab3& tb, uint32_t n_exact, uint32_t refine, cudaStream_t stream)
{
    // the TMA-staged phase 1a needs one CTA per SM in lockstep and room for its two [R | M] buffers behind the working sets
    constexpr size_t kWs = (sizeof(Warp3T<NT>) + 15)/16*16;
    constexpr size_t kStageBytes = 4u*NT*KS*256u + 16u;
    constexpr bool STAGE = CFX_ASTC3_STAGE && LOCK && CTAS == 1 && W*kWs + kStageBytes <= 232448u - 1024u;
    // the HDR variant is its own instantiation, so that the LDR kernel carries none of its code
    const void* k = tb.hdr ? reinterpret_cast<const void*>(&astc3_kernel<NT, KS, W, CTAS, LOCK, true, STAGE>)
                           : reinterpret_cast<const void*>(&astc3_kernel<NT, KS, W, CTAS, LOCK, false, STAGE>);
    const size_t smem = W*kWs + (STAGE ? kStageBytes : 0u);
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) return -4;
    const uint32_t ctas_needed = (p.total_blocks + W - 1)/W;
    const uint32_t grid = min(ctas_needed, persistent_ctas(k, W*32, smem));
    void* args[] = {const_cast<EncodeParams*>(&p), const_cast<Tab3*>(&tb), &n_exact, &refine};
    if (cudaLaunchKernel(k, dim3(grid), dim3(W*32), args, smem, stream) != cudaSuccess) return -4;
    return 1;
}

template <int NT, int KS>
int launch_one(const EncodeParams& p, const Tab3& tb, uint32_t n_exact, uint32_t refine, cudaStream_t stream)
{
#ifdef CFX_ASTC3_TUNE
    if constexpr (NT <= 8) {
    // developer knob: compare launch shapes without rebuilding (CFX_ASTC3_CFG = warps*100 + ctas*10 + lockstep)
    static const int cfg = getenv("CFX_ASTC3_CFG") ? atoi(getenv("CFX_ASTC3_CFG")) : 0;
    if (cfg == 830) return launch_cfg<NT, KS, 8, 3, false>(p, tb, n_exact, refine, stream);
    if (cfg == 831) return launch_cfg<NT, KS, 8, 3, true>(p, tb, n_exact, refine, stream);
    if (cfg == 1221) return launch_cfg<NT, KS, 12, 2, true>(p, tb, n_exact, refine, stream);
    if (cfg == 1611) return launch_cfg<NT, KS, 16, 1, true>(p, tb, n_exact, refine, stream);
    if (cfg == 2410) return launch_cfg<NT, KS, 24, 1, false>(p, tb, n_exact, refine, stream);
    if (cfg == 2411) return launch_cfg<NT, KS, 24, 1, true>(p, tb, n_exact, refine, stream);
    if (cfg == 2011) return launch_cfg<NT, KS, 20, 1, true>(p, tb, n_exact, refine, stream);
    if (cfg == 1211) return launch_cfg<NT, KS, 12, 1, true>(p, tb, n_exact, refine, stream);
    if (cfg == 1411) return launch_cfg<NT, KS, 14, 1, true>(p, tb, n_exact, refine, stream);
    if (cfg == 1811) return launch_cfg<NT, KS, 18, 1, true>(p, tb, n_exact, refine, stream);
    if (cfg == 1610) return launch_cfg<NT, KS, 16, 1, false>(p, tb, n_exact, refine, stream);
    }
#endif
    // ONE CTA per SM.  The kernel is ~15k SASS instructions and only runs well while all the warps of an SM walk the same
    // part of it (measured at 6x6: 16 warps without the barriers 81 MTexel/s, with them 412): two CTAs per SM are two
    // instruction streams competing for one instruction cache (12 warps x 2 CTAs: 359), one CTA of 20 warps is one stream,
    // with 96 registers per thread instead of 80 (no spills) and 75 KB more L1 for the tables (428; 16 warps 412, 22
    // warps 413, 24 warps 424).  The larger footprints take as many warps as their working set leaves room for, up to the
    // count that measured best.
    if constexpr (NT <= 6) return launch_cfg<NT, KS, 20, 1, true>(p, tb, n_exact, refine, stream);        // 4x4 ... 8x6: 5.5 - 10 KB per warp
    else if constexpr (NT <= 8) return launch_cfg<NT, KS, 16, 1, true>(p, tb, n_exact, refine, stream);   // ... 8x8: 12.7 KB (18 warps: -9 %)
    else if constexpr (NT <= 10) return launch_cfg<NT, KS, 12, 1, true>(p, tb, n_exact, refine, stream);  // 10x8: 18 KB (8 warps: -20 %)
    else return launch_cfg<NT, KS, 8, 1, true>(p, tb, n_exact, refine, stream);                           // 10x10 (11 warps: the same) ... 12x12 (10 warps: -23 %)
}


} // namespace

// t3 must describe tables appended to ctx.blob by build_tables3() (astc_host.cu owns the per-device table cache).
int launch_astc3(const EncodeParams& p, const Ctx& ctx, const Astc3Tab& t3, cudaStream_t stream)
{
    // exact evaluations / refinement rounds per Texture::Quality (AstcConverter maps the levels to astcenc's fastest,
    // fast, medium, thorough, exhaustive presets, lib/src/AstcConverter.cpp:174-195)
    static const uint32_t kExact[5] = {4, 6, 8, 16, 32};
    static const uint32_t kRefine[5] = {1, 2, 2, 3, 4};
    const uint32_t q5 = p.quality < 5 ? p.quality : 2;
    const uint32_t n_exact = kExact[q5];
    const uint32_t refine = kRefine[q5];
    if (ctx.tab.n_grids > t3.NT*8u) return -2;          // phase 1c keeps two per-grid arrays in the texel-sized scratch
    Tab3 tb; tb.ctx = ctx; tb.t3 = t3;
#ifdef CFX_ASTC3_TUNE
    // developer knobs, only in builds made with CFX_ASTC3_TUNE=1 (tools/astc_quality.py); never in the shipped library
    static const uint32_t dev_flags = getenv("CFX_ASTC3_FLAGS") ? static_cast<uint32_t>(atoi(getenv("CFX_ASTC3_FLAGS"))) : 0u;
    static const float mis_w = getenv("CFX_ASTC3_MISW") ? static_cast<float>(atof(getenv("CFX_ASTC3_MISW"))) : kMismatchWeight;
    tb.flags = dev_flags;
    tb.mis_w = mis_w;
    tb.dbg_slot = getenv("CFX_ASTC3_SLOT") ? atoi(getenv("CFX_ASTC3_SLOT")) : -1;
    tb.dbg_level = getenv("CFX_ASTC3_LEVEL") ? atoi(getenv("CFX_ASTC3_LEVEL")) : -1;
    tb.dbg_nw = getenv("CFX_ASTC3_NW") ? atoi(getenv("CFX_ASTC3_NW")) : -1;
#else
    tb.flags = 0u;
    tb.mis_w = kMismatchWeight;
    tb.dbg_slot = tb.dbg_level = tb.dbg_nw = -1;
#endif
    tb.hdr = p.type == 4u ? 1u : 0u;                      // Texture::Type::UFloat
    // sRGB textures: astcenc runs with ASTCENC_FLG_USE_PERCEPTUAL (lib/src/AstcConverter.cpp:171-172), i.e. channel error
    // weights 0.30 / 0.59 / 0.11 x 2.25 against 1 for alpha (astcenc_entry.cpp:644-649): 10.8 / 21.2 / 4.0 / 16 sixteenths.
    // They weigh the exact error of phase 2 (which candidate wins, which refinement step is kept), and the hypotheses are
    // built in the weighted colour space (moments, principal lines, ideal weights, line lengths: `sw` in the kernel); the
    // luminance slots, k-means and the colour quantisation term of the estimates stay in plain RGB.
    tb.cw = p.color_space == 1u && !tb.hdr ? 0x1004150Bu : 0x10101010u;
    {
        const float w[4] = {static_cast<float>(tb.cw & 0xFFu)/16.0f, static_cast<float>((tb.cw >> 8) & 0xFFu)/16.0f,
            static_cast<float>((tb.cw >> 16) & 0xFFu)/16.0f, static_cast<float>(tb.cw >> 24)/16.0f};
        tb.sw = make_float4(sqrtf(w[0]), sqrtf(w[1]), sqrtf(w[2]), sqrtf(w[3]));
        tb.isw = make_float4(1.0f/tb.sw.x, 1.0f/tb.sw.y, 1.0f/tb.sw.z, 1.0f/tb.sw.w);
    }
    const uint32_t NT = t3.NT, KS = t3.KS;
    if (NT == 2 && KS == 1) return launch_one<2, 1>(p, tb, n_exact, refine, stream);
    if (NT == 3 && KS == 2) return launch_one<3, 2>(p, tb, n_exact, refine, stream);
    if (NT == 4 && KS == 2) return launch_one<4, 2>(p, tb, n_exact, refine, stream);
    if (NT == 5 && KS == 3) return launch_one<5, 3>(p, tb, n_exact, refine, stream);
    if (NT == 6 && KS == 3) return launch_one<6, 3>(p, tb, n_exact, refine, stream);
    if (NT == 7 && KS == 4) return launch_one<7, 4>(p, tb, n_exact, refine, stream);
    if (NT == 8 && KS == 4) return launch_one<8, 4>(p, tb, n_exact, refine, stream);
    if (NT == 10 && KS == 5) return launch_one<10, 5>(p, tb, n_exact, refine, stream);
    if (NT == 13 && KS == 7) return launch_one<13, 7>(p, tb, n_exact, refine, stream);
    if (NT == 15 && KS == 8) return launch_one<15, 8>(p, tb, n_exact, refine, stream);
    if (NT == 18 && KS == 9) return launch_one<18, 9>(p, tb, n_exact, refine, stream);
    return -2;
}

} // namespace cfx
