// ASTC LDR encoder, warp-cooperative search (v2): one warp owns one block and evaluates candidate
// encodings ONE AFTER ANOTHER with LANE = TEXEL / GRID WEIGHT, so every step is warp-uniform:
//   setup    (shared with astc.cu) single-subset line, 2/3-means clustering, partition-seed ranking by
//            popcount mismatch + exact line-fit residual, dual-plane hypotheses -> up to 9 "slots"
//   per slot, per weight grid (candidates are ordered so that modes sharing a grid are adjacent):
//     decimate  lane j: factor-weighted mean of the ideal weights of the texels grid weight j touches,
//               then one least-squares step against the infill residual (lane i: residual of texel i)
//     per quantisation level of that grid:
//       quantise (lane j) -> infill (lane i, 4 taps) -> least-squares end points from INTEGER moment
//       sums reduced with redux.sync (11 per subset) -> end points quantised at the colour level the
//       left-over bits allow (lane c: component c) -> exact decoded error (lane i) -> redux.sync
//   refine   the best candidate: re-project texels on its quantised end points, re-decimate, re-solve
//   pack     lane 0 BISE-packs the 128-bit block
// Texel values are 8x fixed point integers (exact for RGBA8 sources), so all sums are exact integers.
#include "astc_core.cuh"
#include "common.cuh"
#include "kernels.h"

namespace cfx {

using namespace astc;

namespace {

constexpr int kWarps2 = 8;
constexpr int FX = 8;                      // texel fixed point scale

struct WarpState {
    BlockState st;
    float g[2][kMaxTexels];                // decimated ideal grid weights of the current (slot, grid), per plane
    float res[kMaxTexels];                 // per-texel scratch (infill residual / re-projected weights)
    float tproj[2][kMaxTexels];            // refinement: re-projected ideal weights per plane
    int4 v[kMaxTexels];                    // texels, FX fixed point
    int ep[24];                            // quantised end points of the candidate: [subset][e0 rgba, e1 rgba]
    int best_ep[24];
    uint8_t su[2*kMaxTexels];              // candidate grid weights (unquantised values 0..64), bit-stream order
    uint8_t best_su[2*kMaxTexels];
};

__device__ __forceinline__ int redux_add(int v) { return __reduce_add_sync(0xFFFFFFFFu, v); }
__device__ __forceinline__ uint32_t redux_addu(uint32_t v) { return __reduce_add_sync(0xFFFFFFFFu, v); }

// Decimate the ideal weights `tt` (per texel) onto grid `gi`: ws.g[plane][j].  Returns the sum over
// texels of (infill of the decimated weights - ideal weight)^2, i.e. what no quantisation level of this
// grid can undo.
__device__ __forceinline__ float decimate(const Ctx& c, WarpState& ws, uint32_t gi, uint32_t nw, const float* tt, uint32_t plane,
    uint32_t lane, const Slot* bound_slot = nullptr)
{
    const uint32_t T = c.tab.texels;
    const uint32_t start_off = c.tab.off_csr_start + gi*(kMaxTexels + 2)*2u;
    const uint32_t ent_off = c.tab.off_csr_ent + gi*4u*T*2u;
    const uint32_t norm_off = c.tab.off_wnorm + gi*kMaxTexels*4u;
    const uint32_t inf_off = c.tab.off_infill + gi*T*8u;
    for (uint32_t j = lane; j < nw; j += 32) {
        const uint32_t e0 = tab_u16(c, start_off + j*2u), e1 = tab_u16(c, start_off + (j + 1)*2u);
        float s = 0.0f;
        for (uint32_t e = e0; e < e1; ++e) {
            const uint32_t ent = tab_u16(c, ent_off + e*2u);
            s += static_cast<float>(ent >> 8)*tt[ent & 0xFFu];
        }
        ws.g[plane][j] = s*tab_f32(c, norm_off + j*4u);
    }
    __syncwarp();
    if (nw >= T) return 0.0f;
    for (uint32_t i = lane; i < T; i += 32) {
        const uint2 inf = tab_u32x2(c, inf_off + i*8u);
        float r = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) r += static_cast<float>((inf.y >> (8*k)) & 0xFFu)*ws.g[plane][(inf.x >> (8*k)) & 0xFFu];
        ws.res[i] = tt[i] - r*(1.0f/16.0f);
    }
    __syncwarp();
    for (uint32_t j = lane; j < nw; j += 32) {
        const uint32_t e0 = tab_u16(c, start_off + j*2u), e1 = tab_u16(c, start_off + (j + 1)*2u);
        float s = 0.0f, s2 = 0.0f;
        for (uint32_t e = e0; e < e1; ++e) {
            const uint32_t ent = tab_u16(c, ent_off + e*2u);
            const float f = static_cast<float>(ent >> 8);
            s += f*ws.res[ent & 0xFFu]; s2 += f*f;
        }
        ws.g[plane][j] = fminf(fmaxf(ws.g[plane][j] + (s2 > 0.0f ? kDecimationGain*s/s2 : 0.0f), 0.0f), 1.0f);
    }
    __syncwarp();
    if (!bound_slot) return 0.0f;
    float d2 = 0.0f;
    for (uint32_t i = lane; i < T; i += 32) {
        const uint2 inf = tab_u32x2(c, inf_off + i*8u);
        float r = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) r += static_cast<float>((inf.y >> (8*k)) & 0xFFu)*ws.g[plane][(inf.x >> (8*k)) & 0xFFu];
        const float d = tt[i] - r*(1.0f/16.0f);
        const uint32_t q = bound_slot->part[i];
        d2 += d*d*(q == 0 ? bound_slot->len2[0] : (q == 1 ? bound_slot->len2[1] : bound_slot->len2[2]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xFFFFFFFFu, d2, o);
    return d2;
}

struct Best { float err; uint32_t slot, mode, cl; };

// Evaluate block mode `mi` on slot `s` with the decimated weights in ws.g; on success ws.su / ws.ep hold
// the candidate and its exact decoded error is returned.
template <int K>
__device__ __forceinline__ float evaluate2(const Ctx& c, WarpState& ws, const Slot& slot, const ModeInfo& m, bool has_alpha,
    uint32_t lane, uint32_t& cl_out)
{
    const uint32_t T = c.tab.texels;
    const uint32_t pc = slot.pc;
    const int dc = slot.dual_ch;
    const uint32_t planes = dc >= 0 ? 2u : 1u;
    const uint32_t n_ints = pc*(has_alpha ? 8u : 6u);
    const int avail = 128 - static_cast<int>(m.wbits) - (pc == 1 ? 17 : 29) - (dc >= 0 ? 2 : 0);
    if (n_ints > 18 || avail < 0) return 3.0e38f;
    const uint32_t cl = tab_u8(c, c.tab.off_clevel + (n_ints >> 1)*128u + static_cast<uint32_t>(avail));
    if (cl == 0xFF) return 3.0e38f;
    cl_out = cl;
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
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t i = lane + 32u*k;
        live[k] = i < T;
        const uint32_t ii = live[k] ? i : 0u;
        part[k] = slot.part[ii];
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
        // lane c (0..7) of this subset solves and quantises component c: e0 r,g,b,a then e1 r,g,b,a
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

__device__ __forceinline__ void keep_best(WarpState& ws, const ModeInfo& m, uint32_t planes, uint32_t pc, uint32_t lane)
{
    for (uint32_t j = lane; j < static_cast<uint32_t>(m.nw)*planes; j += 32) ws.best_su[j] = ws.su[j];
    if (lane < pc*8u) ws.best_ep[lane] = ws.ep[lane];
    __syncwarp();
}

} // namespace

template <int K>
__global__ void __launch_bounds__(kWarps2*32) astc2_kernel(const EncodeParams p, const Ctx ctx, const Plan plan, uint32_t quality)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = lane_id(), warp = warp_id();
    WarpState& ws = *reinterpret_cast<WarpState*>(smem + warp*((sizeof(WarpState) + 15)/16*16));
    BlockState& st = ws.st;
    const uint32_t T = ctx.tab.texels, bw = ctx.tab.bw, bh = ctx.tab.bh;
    const bool alpha_off = p.alpha_type == 0;

    for (uint32_t blk = blockIdx.x*kWarps2 + warp; blk < p.total_blocks; blk += gridDim.x*kWarps2) {
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

        // ---- search: candidates one after another, every step warp-uniform
        Best best; best.err = 3.0e38f; best.slot = 0; best.mode = 0; best.cl = 0;
        // good enough: astcenc's own medium-preset quality target for this footprint
        // (astcenc_entry.cpp:546-552: max(95 - 35 log10 T, 70 - 19 log10 T) dB) plus a 12 dB margin ends the search
        const float stop_db = fmaxf(95.0f - 35.0f*log10f(static_cast<float>(T)), 70.0f - 19.0f*log10f(static_cast<float>(T))) + 12.0f;
        const float stop_err = 65025.0f*exp10f(-0.1f*stop_db)*static_cast<float>(T*(has_alpha ? 4u : 3u))*static_cast<float>(FX*FX);
        for (uint32_t s = 0; s < plan.slots && best.err > stop_err; ++s) {
            const Slot& slot = st.slots[s];
            if (!slot.valid) continue;
            const uint32_t type = slot_type(s);
            const uint32_t n = ctx.tab.n_cand_q[quality][type];
            const uint32_t list = ctx.tab.off_cand_q[quality][type];
            const uint32_t planes = slot.dual_ch >= 0 ? 2u : 1u;
            // nothing on this slot's lines can beat what we already have
            const float fx2 = static_cast<float>(FX*FX);
            if (0.9f*fx2*slot.e_line > best.err) continue;
            uint32_t cur_grid = 0xFFFFFFFFu;
            bool skip_grid = false;
            for (uint32_t ci = 0; ci < n && best.err > stop_err; ++ci) {
                const uint32_t mi = tab_u16(ctx, list + ci*2u);
                const ModeInfo m = tab_mode(ctx, mi);
                if (m.grid != cur_grid) {
                    cur_grid = m.grid;
                    // lower bound of this grid's error: the slot's line residual plus what decimation alone loses
                    const bool bound = best.err < 3.0e38f && planes == 1 && m.nw < T;
                    const float d2 = decimate(ctx, ws, cur_grid, m.nw, slot.t, 0, lane, bound ? &slot : nullptr);
                    if (planes == 2) decimate(ctx, ws, cur_grid, m.nw, slot.t2, 1, lane);
                    skip_grid = bound && 0.8f*fx2*(slot.e_line + d2) > best.err;
                }
                if (skip_grid) continue;
                uint32_t cl = 0;
                const float err = evaluate2<K>(ctx, ws, slot, m, has_alpha, lane, cl);
                if (err < best.err) {
                    best.err = err; best.slot = s; best.mode = mi; best.cl = cl;
                    keep_best(ws, m, planes, slot.pc, lane);
                }
                __syncwarp();
            }
        }

        // ---- refine the winner: re-project on its end points, re-decimate, re-solve
        const Slot& bslot = st.slots[best.slot];
        const ModeInfo bm = tab_mode(ctx, best.mode);
        const uint32_t bplanes = bslot.dual_ch >= 0 ? 2u : 1u;
        for (uint32_t r = 0; r < plan.refine && best.err > 0.0f; ++r) {
            const int dc = bslot.dual_ch;
            for (uint32_t i = lane; i < T; i += 32) {
                const int* e = ws.best_ep + bslot.part[i]*8u;
                const float4 x = st.cf[i];
                const float xs[4] = {x.x, x.y, x.z, x.w};
                float num0 = 0, den0 = 0, num1 = 0, den1 = 0;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    if (ch == 3 && !has_alpha) continue;
                    const float a = static_cast<float>(e[ch]), d = static_cast<float>(e[4 + ch]) - a;
                    if (ch == dc) { num1 += (xs[ch] - a)*d; den1 += d*d; } else { num0 += (xs[ch] - a)*d; den0 += d*d; }
                }
                ws.tproj[0][i] = den0 > 0.0f ? fminf(fmaxf(num0/den0, 0.0f), 1.0f) : 0.0f;
                ws.tproj[1][i] = den1 > 0.0f ? fminf(fmaxf(num1/den1, 0.0f), 1.0f) : 0.0f;
            }
            __syncwarp();
            decimate(ctx, ws, bm.grid, bm.nw, ws.tproj[0], 0, lane);
            if (bplanes == 2) decimate(ctx, ws, bm.grid, bm.nw, ws.tproj[1], 1, lane);
            uint32_t cl = 0;
            const float err = evaluate2<K>(ctx, ws, bslot, bm, has_alpha, lane, cl);
            if (err < best.err) { best.err = err; keep_best(ws, bm, bplanes, bslot.pc, lane); }
            else break;
            __syncwarp();
        }
        __syncwarp();
        if (lane == 0) {
            Enc enc;
            enc.clevel = best.cl; enc.err = best.err;
            for (uint32_t s = 0; s < bslot.pc; ++s) {
                const int* e = ws.best_ep + s*8u;
                enc.ep[s][0] = static_cast<uint32_t>(e[0]) | (static_cast<uint32_t>(e[1]) << 8) | (static_cast<uint32_t>(e[2]) << 16) |
                    (static_cast<uint32_t>(e[3]) << 24);
                enc.ep[s][1] = static_cast<uint32_t>(e[4]) | (static_cast<uint32_t>(e[5]) << 8) | (static_cast<uint32_t>(e[6]) << 16) |
                    (static_cast<uint32_t>(e[7]) << 24);
            }
            *dst = pack_block(ctx, bslot, bm, enc, has_alpha, ws.best_su, 0, true);
        }
    }
}

int launch_astc2(const EncodeParams& p, const Ctx& ctx, cudaStream_t stream)
{
    const Plan plan = make_plan(p.quality, ctx.tab);
    const size_t smem = kWarps2*((sizeof(WarpState) + 15)/16*16);
    const bool two = ctx.tab.texels > 32;
    const void* k = two ? reinterpret_cast<const void*>(&astc2_kernel<2>) : reinterpret_cast<const void*>(&astc2_kernel<1>);
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) return -4;
    const uint32_t ctas_needed = (p.total_blocks + kWarps2 - 1)/kWarps2;
    const uint32_t grid = min(ctas_needed, persistent_ctas(k, kWarps2*32, smem));
    uint32_t quality = p.quality;
    void* args[] = {const_cast<EncodeParams*>(&p), const_cast<Ctx*>(&ctx), const_cast<Plan*>(&plan), &quality};
    if (cudaLaunchKernel(k, dim3(grid), dim3(kWarps2*32), args, smem, stream) != cudaSuccess) return -4;
    return 1;
}

} // namespace cfx
