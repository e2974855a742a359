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
#include "common.cuh"
#include "kernels.h"

#include <cuda_fp16.h>

namespace cfx {

using namespace astc;

namespace {

constexpr int kWarps3 = 8;
constexpr int FX = 8;                          // texel fixed point scale
constexpr int kTaStride = kMaxTexels + 8;      // halfs; 36 words: conflict-free A fragment loads
constexpr int kMaxCand = 16;

struct Slot3 {
    float4 e0[3], e1[3];       // line end points (0..255 per channel), sum(e1.rgb) >= sum(e0.rgb)
    float len2[3];             // squared length of each subset's line (first plane)
    float len2b;               // squared length of the second plane's line (dual plane slots)
    float e_line;              // squared distance of the texels from their lines: error floor of the slot
    uint32_t pc, seed, valid;
    int32_t dual_ch;
};

struct Warp3 {
    int4 v[kMaxTexels];                     // texels, FX fixed point
    Slot3 slots[kSlots];
    uint8_t part[4][kMaxTexels];            // subset of every texel for slots 1..4
    __half ta[kRows3][kTaStride];           // A operand: ideal weights per slot plane (rows 0..8 first planes,
                                            // 9..12 second planes of slots 5..8, 13/14 refinement scratch)
    float D[13][kMaxGrids3];                // decimation loss per slot plane and grid
    float Sm[4][kMaxGrids3];                // sum_i len2_i kappa_gi for the multi-subset slots 1..4
    float g[2][kMaxTexels];                 // decimated ideal grid weights of the current candidate, per plane
    int ep[24];                             // quantised end points of the candidate: [subset][e0 rgba, e1 rgba]
    int best_ep[24];
    uint8_t su[2*kMaxTexels];               // candidate grid weights (unquantised values 0..64), bit-stream order
    uint8_t best_su[2*kMaxTexels];
};

// model constants (fitted on the host with tools/emu_astc3.py)
constexpr float kLine = 1.1f, kDec = 1.0f, kQuant = 0.9f, kColor = 0.5f;

__constant__ float c_qvar[kWeightLevels];     // expected squared quantisation error of a uniform weight, after end point refit
__constant__ float c_cvar[kColorLevels];      // expected squared error per texel channel from end point quantisation

struct Tab3 { Ctx ctx; Astc3Tab t3; };

__device__ __forceinline__ int redux_add(int v) { return __reduce_add_sync(0xFFFFFFFFu, v); }
__device__ __forceinline__ uint32_t redux_addu(uint32_t v) { return __reduce_add_sync(0xFFFFFFFFu, v); }

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint2 b)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

template <int KS>
__device__ __forceinline__ void load_a(const Warp3& ws, uint32_t (&a)[KS][4], uint32_t lane)
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
template <int KS>
__device__ __forceinline__ void decimate_mma(const Tab3& tb, Warp3& ws, const uint32_t (&a)[KS][4], uint32_t grid, uint32_t nw,
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

// Evaluate block mode m on slot s with the decimated weights in ws.g; on success ws.su / ws.ep hold the candidate and
// its exact decoded error (FX^2 units) is returned.
template <int K>
__device__ __forceinline__ float evaluate3(const Ctx& c, Warp3& ws, uint32_t s, const ModeInfo& m, uint32_t cl, bool has_alpha,
    uint32_t lane)
{
    const uint32_t T = c.tab.texels;
    const Slot3& slot = ws.slots[s];
    const uint32_t pc = slot.pc;
    const int dc = slot.dual_ch;
    const uint32_t planes = dc >= 0 ? 2u : 1u;
    const uint32_t L = m.level, nw = m.nw;
    const float nm1 = static_cast<float>(kWqN[L] - 1);
    // quantise (lane = grid weight)
    for (uint32_t pl = 0; pl < planes; ++pl)
        for (uint32_t j = lane; j < nw; j += 32) {
            const int k = min(max(__float2int_rn(ws.g[pl][j]*nm1), 0), static_cast<int>(kWqN[L]) - 1);
            ws.su[j*planes + pl] = static_cast<uint8_t>(tab_u8(c, c.tab.off_wq_val + L*32u + static_cast<uint32_t>(k)));
        }
    __syncwarp();
    // infill (lane = texel)
    int w[K][2];
    uint32_t part[K];
    bool live[K];
    const uint32_t inf_off = c.tab.off_infill + static_cast<uint32_t>(m.grid)*T*8u;
    const uint8_t* parts = ws.part[(s - 1u) & 3u];
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
    // least-squares end points per subset from integer moment sums
    for (uint32_t p = 0; p < pc; ++p) {
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
            int q = 255;
            if (ch < 3 || has_alpha) {
                const int iv = min(max(__float2int_rn(val), 0), 255);
                const uint32_t rank = tab_u8(c, c.tab.off_cq_near + cl*256u + static_cast<uint32_t>(iv));
                q = static_cast<int>(tab_u8(c, c.tab.off_cq_val + cl*256u + rank));
            }
            ws.ep[p*8u + lane] = q;
        }
    }
    __syncwarp();
    // keep sum(e1.rgb) >= sum(e0.rgb) (otherwise the decoder would blue-contract): swap the end points
    if (lane < pc) {
        int* e = ws.ep + lane*8u;
        if (e[4] + e[5] + e[6] < e[0] + e[1] + e[2]) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { const int tmp = e[k]; e[k] = e[4 + k]; e[4 + k] = tmp; }
        }
    }
    __syncwarp();
    // exact decoded error (lane = texel)
    uint32_t err = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (!live[k]) continue;
        const int* e = ws.ep + part[k]*8u;
        const int4 x = ws.v[lane + 32u*k];
        const int w0 = w[k][0], w1 = w[k][1];
        int d = ((e[0]*FX*(64 - (dc == 0 ? w1 : w0)) + e[4]*FX*(dc == 0 ? w1 : w0) + 32) >> 6) - x.x; err += static_cast<uint32_t>(d*d);
        d = ((e[1]*FX*(64 - (dc == 1 ? w1 : w0)) + e[5]*FX*(dc == 1 ? w1 : w0) + 32) >> 6) - x.y; err += static_cast<uint32_t>(d*d);
        d = ((e[2]*FX*(64 - (dc == 2 ? w1 : w0)) + e[6]*FX*(dc == 2 ? w1 : w0) + 32) >> 6) - x.z; err += static_cast<uint32_t>(d*d);
        if (has_alpha) {
            d = ((e[3]*FX*(64 - (dc == 3 ? w1 : w0)) + e[7]*FX*(dc == 3 ? w1 : w0) + 32) >> 6) - x.w; err += static_cast<uint32_t>(d*d);
        }
    }
    err = redux_addu(err);
    return static_cast<float>(err);
}

__device__ __forceinline__ void keep_best3(Warp3& ws, uint32_t nw, uint32_t planes, uint32_t pc, uint32_t lane)
{
    for (uint32_t j = lane; j < nw*planes; j += 32) ws.best_su[j] = ws.su[j];
    if (lane < pc*8u) ws.best_ep[lane] = ws.ep[lane];
    __syncwarp();
}

struct SlotView {           // what pack_block needs from a slot
    uint32_t pc, seed;
    int32_t dual_ch;
};

} // namespace

template <int NT, int KS>
__global__ void __launch_bounds__(kWarps3*32) astc3_kernel(const EncodeParams p, const Tab3 tb, uint32_t n_exact, uint32_t refine)
{
    constexpr int K = (NT*8 > 32) ? 2 : 1;
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = lane_id(), warp = warp_id();
    const uint32_t gq = lane >> 2, tq = lane & 3u;
    constexpr size_t kWsBytes = (sizeof(Warp3) + 15)/16*16, kStBytes = (sizeof(BlockState) + 15)/16*16;
    Warp3& ws = *reinterpret_cast<Warp3*>(smem + warp*(kWsBytes + kStBytes));
    BlockState& st = *reinterpret_cast<BlockState*>(smem + warp*(kWsBytes + kStBytes) + kWsBytes);
    const Ctx& ctx = tb.ctx;
    const uint32_t T = ctx.tab.texels, bw = ctx.tab.bw, bh = ctx.tab.bh;
    const uint32_t G = ctx.tab.n_grids;
    const bool alpha_off = p.alpha_type == 0;
    const float fx2 = static_cast<float>(FX*FX);

    for (uint32_t blk = blockIdx.x*kWarps3 + warp; blk < p.total_blocks; blk += gridDim.x*kWarps3) {
        const uint32_t by = blk / p.blocks_x, bx = blk - by*p.blocks_x;
        __syncwarp();
        bool differs = false, alpha = false;
        for (uint32_t i = lane; i < T; i += 32) {
            const uint32_t ty = i / bw, tx = i - ty*bw;
            const uint32_t x = min(bx*bw + tx, p.width - 1), y = min(by*bh + ty, p.height - 1);
            float4 v;
            if (p.src_format == SRC_RGBA8) {
                const uint32_t q = __ldg(reinterpret_cast<const uint32_t*>(p.src + static_cast<uint64_t>(y)*p.pitch) + x);
                v = make_float4(static_cast<float>(q & 0xFF), static_cast<float>((q >> 8) & 0xFF), static_cast<float>((q >> 16) & 0xFF),
                    static_cast<float>(q >> 24));
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
            ws.v[i] = make_int4(__float2int_rn(v.x*FX), __float2int_rn(v.y*FX), __float2int_rn(v.z*FX), __float2int_rn(v.w*FX));
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
        // ---- setup: partition hypotheses (lane-local steps shared with astc.cu)
        if (lane == 0) st.has_alpha = has_alpha ? 1u : 0u;
        if (lane < kSlots) st.slots[lane].valid = 0;
        __syncwarp();
        step_init(ctx, st, lane);
        __syncwarp();
        step_rank(ctx, st, lane);
        __syncwarp();
        step_score(ctx, st, lane);
        __syncwarp();
        step_slots(ctx, st, lane);
        __syncwarp();
        // slots -> compact form + fp16 A operand rows
        if (lane < kSlots) {
            const Slot& o = st.slots[lane];
            Slot3& n = ws.slots[lane];
            n.valid = o.valid; n.pc = o.pc; n.seed = o.seed; n.dual_ch = o.dual_ch; n.e_line = o.e_line;
            for (int k = 0; k < 3; ++k) { n.e0[k] = o.e0[k]; n.e1[k] = o.e1[k]; n.len2[k] = o.len2[k]; }
            float lb = 0.0f;
            if (o.valid && o.dual_ch >= 0) { const float d = ch(o.e1[0], o.dual_ch) - ch(o.e0[0], o.dual_ch); lb = d*d; }
            n.len2b = lb;
        }
        for (uint32_t r = 0; r < kRows3; ++r) {
            const uint32_t s = r < 9 ? r : r - 4;
            const bool ok = r < 13 && st.slots[s].valid;
            const float* src = r < 9 ? st.slots[s].t : st.slots[s].t2;
            for (uint32_t i = lane; i < kMaxTexels; i += 32)
                ws.ta[r][i] = __float2half_rn(ok && i < T ? src[i] : 0.0f);
        }
        for (uint32_t s = 1; s <= 4; ++s)
            for (uint32_t i = lane; i < kMaxTexels; i += 32) ws.part[s - 1][i] = i < T ? st.slots[s].part[i] : 0;
        __syncwarp();

        // ---- phase 1a: decimation loss D[slot plane][grid] on the tensor cores
        uint32_t a[KS][4];
        load_a<KS>(ws, a, lane);
        {
            float lw[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const uint32_t i = nt*8 + 2*tq + e;
                    float wgt = i < T ? 1.0f : 0.0f;
                    if (gq >= 1 && gq <= 4 && i < T) wgt = ws.slots[gq].len2[ws.part[gq - 1][i]];
                    lw[nt][e] = wgt;
                }
            const float scale0 = gq == 0 ? ws.slots[0].len2[0] : (gq <= 4 ? 1.0f : ws.slots[gq].len2[0]);
            const float scale1 = gq == 0 ? ws.slots[8].len2[0] : (gq <= 4 ? ws.slots[gq + 4].len2b : 0.0f);
            const uint32_t* ridx = reinterpret_cast<const uint32_t*>(ctx.blob + tb.t3.off_rfrag_idx);
            for (uint32_t g = 0; g < G; ++g) {
                const uint32_t off = __ldg(ridx + g);
                float acc0 = 0.0f, acc1 = 0.0f;
                if (off) {
                    const uint2* frag = reinterpret_cast<const uint2*>(ctx.blob + off) + lane;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) mma16816(c, a[ks], __ldg(frag + (nt*KS + ks)*32));
                        acc0 += lw[nt][0]*c[0]*c[0] + lw[nt][1]*c[1]*c[1];
                        acc1 += c[2]*c[2] + c[3]*c[3];
                    }
                    acc0 += __shfl_xor_sync(0xFFFFFFFFu, acc0, 1); acc0 += __shfl_xor_sync(0xFFFFFFFFu, acc0, 2);
                    acc1 += __shfl_xor_sync(0xFFFFFFFFu, acc1, 1); acc1 += __shfl_xor_sync(0xFFFFFFFFu, acc1, 2);
                }
                if (tq == 0) {
                    ws.D[gq][g] = acc0*scale0;
                    if (gq + 8 < 13) ws.D[gq + 8][g] = acc1*scale1;
                }
            }
        }
        // ---- phase 1b: S of the multi-subset slots (lane = grid)
        for (uint32_t g = lane; g < G; g += 32) {
            const float* kap = reinterpret_cast<const float*>(ctx.blob + tb.t3.off_kappa) + g*kMaxTexels;
            float s1 = 0.0f, s2 = 0.0f, s3 = 0.0f, s4 = 0.0f;
            for (uint32_t i = 0; i < T; ++i) {
                const float k = __ldg(kap + i);
                s1 += k*ws.slots[1].len2[ws.part[0][i]]; s2 += k*ws.slots[2].len2[ws.part[1][i]];
                s3 += k*ws.slots[3].len2[ws.part[2][i]]; s4 += k*ws.slots[4].len2[ws.part[3][i]];
            }
            ws.Sm[0][g] = s1; ws.Sm[1][g] = s2; ws.Sm[2][g] = s3; ws.Sm[3][g] = s4;
        }
        __syncwarp();
        // ---- phase 1c: estimate every (slot, mode); each lane keeps its three best
        float be0 = 3.0e38f, be1 = 3.0e38f, be2 = 3.0e38f;
        uint32_t bc0 = 0, bc1 = 0, bc2 = 0;
        {
            const float tn = static_cast<float>(T*(has_alpha ? 4u : 3u));
            const float* ksum = reinterpret_cast<const float*>(ctx.blob + tb.t3.off_ksum);
            for (uint32_t s = 0; s < kSlots; ++s) {
                const Slot3& slot = ws.slots[s];
                if (!slot.valid) continue;
                const uint32_t type = slot_type(s);
                const uint32_t first = type == 3 ? ctx.tab.n_modes1 : 0u, count = type == 3 ? ctx.tab.n_modes2 : ctx.tab.n_modes1;
                const uint8_t* mcl = ctx.blob + tb.t3.off_modecl + ((has_alpha ? 4u : 0u) + type)*tb.t3.n_modes;
                const float base = kLine*slot.e_line;
                for (uint32_t mi = first + lane; mi < first + count; mi += 32) {
                    const uint32_t cl = __ldg(mcl + mi);
                    if (cl == 0xFFu) continue;
                    const ModeInfo m = tab_mode(ctx, mi);
                    const uint32_t g = m.grid;
                    float dsum = ws.D[s][g], ssum;
                    if (type == 3) { dsum += ws.D[s + 4][g]; ssum = (slot.len2[0] + slot.len2b)*__ldg(ksum + g); }
                    else if (type == 0) ssum = slot.len2[0]*__ldg(ksum + g);
                    else ssum = ws.Sm[s - 1][g];
                    const float est = base + kDec*dsum + kQuant*ssum*c_qvar[m.level] + kColor*tn*c_cvar[cl];
                    const uint32_t code = (s << 16) | mi;
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
        // ---- phase 2: exact evaluation of the n_exact best estimates
        float best_err = 3.0e38f;
        uint32_t best_code = 0, best_cl = 0;
        const float stop_db = fmaxf(95.0f - 35.0f*log10f(static_cast<float>(T)), 70.0f - 19.0f*log10f(static_cast<float>(T))) + 12.0f;
        const float stop_err = 65025.0f*exp10f(-0.1f*stop_db)*static_cast<float>(T*(has_alpha ? 4u : 3u))*fx2;
        for (uint32_t n = 0; n < n_exact && best_err > stop_err; ++n) {
            const uint32_t key = (__float_as_uint(be0) & ~31u) | lane;
            const uint32_t kmin = __reduce_min_sync(0xFFFFFFFFu, key);
            const uint32_t wl = kmin & 31u;
            const float est = __shfl_sync(0xFFFFFFFFu, be0, wl);
            const uint32_t code = __shfl_sync(0xFFFFFFFFu, bc0, wl);
            if (est >= 3.0e38f) break;
            if (lane == wl) { be0 = be1; bc0 = bc1; be1 = be2; bc1 = bc2; be2 = 3.0e38f; }
            if (0.8f*est*fx2 > best_err) break;              // estimates are sorted: nothing later can win
            const uint32_t s = code >> 16, mi = code & 0xFFFFu;
            const Slot3& slot = ws.slots[s];
            const ModeInfo m = tab_mode(ctx, mi);
            const uint32_t cl = __ldg(ctx.blob + tb.t3.off_modecl + ((has_alpha ? 4u : 0u) + slot_type(s))*tb.t3.n_modes + mi);
            decimate_mma<KS>(tb, ws, a, m.grid, m.nw, static_cast<int>(s), slot.dual_ch >= 0 ? static_cast<int>(s) + 4 : -1, lane);
            const float err = evaluate3<K>(ctx, ws, s, m, cl, has_alpha, lane);
            if (err < best_err) {
                best_err = err; best_code = code; best_cl = cl;
                keep_best3(ws, m.nw, slot.dual_ch >= 0 ? 2u : 1u, slot.pc, lane);
            }
            __syncwarp();
        }
        // ---- refine the winner: re-project on its end points, re-decimate, re-solve
        const uint32_t bs = best_code >> 16;
        const Slot3& bslot = ws.slots[bs];
        const ModeInfo bm = tab_mode(ctx, best_code & 0xFFFFu);
        const uint32_t bplanes = bslot.dual_ch >= 0 ? 2u : 1u;
        for (uint32_t r = 0; r < refine && best_err > 0.0f; ++r) {
            const int dc = bslot.dual_ch;
            const uint8_t* parts = ws.part[(bs - 1u) & 3u];
            for (uint32_t i = lane; i < T; i += 32) {
                const int* e = ws.best_ep + (bslot.pc > 1 ? parts[i] : 0u)*8u;
                const int4 xi = ws.v[i];
                const float xs[4] = {static_cast<float>(xi.x)*(1.0f/FX), static_cast<float>(xi.y)*(1.0f/FX),
                    static_cast<float>(xi.z)*(1.0f/FX), static_cast<float>(xi.w)*(1.0f/FX)};
                float num0 = 0, den0 = 0, num1 = 0, den1 = 0;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    if (c4 == 3 && !has_alpha) continue;
                    const float a0 = static_cast<float>(e[c4]), d = static_cast<float>(e[4 + c4]) - a0;
                    if (c4 == dc) { num1 += (xs[c4] - a0)*d; den1 += d*d; } else { num0 += (xs[c4] - a0)*d; den0 += d*d; }
                }
                ws.ta[13][i] = __float2half_rn(den0 > 0.0f ? fminf(fmaxf(num0/den0, 0.0f), 1.0f) : 0.0f);
                ws.ta[14][i] = __float2half_rn(den1 > 0.0f ? fminf(fmaxf(num1/den1, 0.0f), 1.0f) : 0.0f);
            }
            __syncwarp();
            uint32_t a2[KS][4];
            load_a<KS>(ws, a2, lane);
            decimate_mma<KS>(tb, ws, a2, bm.grid, bm.nw, 13, bplanes == 2 ? 14 : -1, lane);
            const float err = evaluate3<K>(ctx, ws, bs, bm, best_cl, has_alpha, lane);
            if (err < best_err) { best_err = err; keep_best3(ws, bm.nw, bplanes, bslot.pc, lane); }
            else break;
            __syncwarp();
        }
        __syncwarp();
        if (lane == 0) {
            Enc enc;
            enc.clevel = best_cl; enc.err = best_err;
            for (uint32_t s = 0; s < bslot.pc; ++s) {
                const int* e = ws.best_ep + s*8u;
                enc.ep[s][0] = static_cast<uint32_t>(e[0]) | (static_cast<uint32_t>(e[1]) << 8) | (static_cast<uint32_t>(e[2]) << 16) |
                    (static_cast<uint32_t>(e[3]) << 24);
                enc.ep[s][1] = static_cast<uint32_t>(e[4]) | (static_cast<uint32_t>(e[5]) << 8) | (static_cast<uint32_t>(e[6]) << 16) |
                    (static_cast<uint32_t>(e[7]) << 24);
            }
            SlotView sv; sv.pc = bslot.pc; sv.seed = bslot.seed; sv.dual_ch = bslot.dual_ch;
            *dst = pack_block(ctx, sv, bm, enc, has_alpha, ws.best_su, 0, true);
        }
    }
}

namespace {

template <int NT, int KS>
int launch_one(const EncodeParams& p, const Tab3& tb, uint32_t n_exact, uint32_t refine, cudaStream_t stream)
{
    const void* k = reinterpret_cast<const void*>(&astc3_kernel<NT, KS>);
    const size_t smem = kWarps3*((sizeof(Warp3) + 15)/16*16 + (sizeof(BlockState) + 15)/16*16);
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) return -4;
    const uint32_t ctas_needed = (p.total_blocks + kWarps3 - 1)/kWarps3;
    const uint32_t grid = min(ctas_needed, persistent_ctas(k, kWarps3*32, smem));
    void* args[] = {const_cast<EncodeParams*>(&p), const_cast<Tab3*>(&tb), &n_exact, &refine};
    if (cudaLaunchKernel(k, dim3(grid), dim3(kWarps3*32), args, smem, stream) != cudaSuccess) return -4;
    return 1;
}

bool g_consts_set[16] = {};

} // namespace

// t3 must describe tables appended to ctx.blob by build_tables3() (astc.cu owns the per-device table cache).
int launch_astc3(const EncodeParams& p, const Ctx& ctx, const Astc3Tab& t3, cudaStream_t stream)
{
    int device = 0;
    cudaGetDevice(&device);
    if (device >= 0 && device < 16 && !g_consts_set[device]) {
        float qv[kWeightLevels], cv[kColorLevels];
        for (int l = 0; l < kWeightLevels; ++l) {
            const float n1 = static_cast<float>(kWeightQuant[l].n - 1);
            qv[l] = (1.0f/(n1*n1))*(1.0f/12.0f)*(1.0f - 0.75f/n1);
        }
        for (int l = 0; l < kColorLevels; ++l) {
            const float step = 255.0f/static_cast<float>(kColorQuant[l].n - 1);
            cv[l] = step*step*(1.0f/18.0f);
        }
        if (cudaMemcpyToSymbol(c_qvar, qv, sizeof(qv)) != cudaSuccess) return -4;
        if (cudaMemcpyToSymbol(c_cvar, cv, sizeof(cv)) != cudaSuccess) return -4;
        g_consts_set[device] = true;
    }
    static const uint32_t kExact[5] = {2, 4, 8, 12, 16};
    const uint32_t n_exact = kExact[p.quality < 5 ? p.quality : 2];
    const uint32_t refine = p.quality >= 3 ? 3u : 2u;
    Tab3 tb; tb.ctx = ctx; tb.t3 = t3;
    const uint32_t NT = t3.NT, KS = t3.KS;
    if (NT == 2 && KS == 1) return launch_one<2, 1>(p, tb, n_exact, refine, stream);
    if (NT == 3 && KS == 2) return launch_one<3, 2>(p, tb, n_exact, refine, stream);
    if (NT == 4 && KS == 2) return launch_one<4, 2>(p, tb, n_exact, refine, stream);
    if (NT == 5 && KS == 3) return launch_one<5, 3>(p, tb, n_exact, refine, stream);
    if (NT == 6 && KS == 3) return launch_one<6, 3>(p, tb, n_exact, refine, stream);
    if (NT == 7 && KS == 4) return launch_one<7, 4>(p, tb, n_exact, refine, stream);
    if (NT == 8 && KS == 4) return launch_one<8, 4>(p, tb, n_exact, refine, stream);
    return -2;
}

} // namespace cfx
