// Lane-local part of the ASTC LDR encoder (see astc.cu for how a warp drives it).  Our own search,
// not a port of astcenc: every lane evaluates whole candidate encodings (partitioning x block mode)
// with exact decoded error, keeps its best, refines it by least squares and the warp takes the
// minimum.  Compiles for the device and, through hostdev.h, for tools/emu_astc.cpp.
//
// Replaces AstcConverter::process -> astcenc_compress_image per block
// (lib/src/AstcConverter.cpp:208-230; search driver compress_block,
// lib/astc-encoder/Source/astcenc_compress_symbolic.cpp:1163-1454).
#pragma once
#include "astc_tables.hpp"
#include "hostdev.h"

namespace cfx {
namespace astc {

CFX_CONST uint8_t kWqN[kWeightLevels] = {2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32};
CFX_CONST uint8_t kWqBits[kWeightLevels] = {1, 0, 2, 0, 1, 3, 1, 2, 4, 2, 3, 5};
CFX_CONST uint8_t kWqTrits[kWeightLevels] = {0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0};
CFX_CONST uint8_t kWqQuints[kWeightLevels] = {0, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0};
CFX_CONST uint8_t kCqBits[kColorLevels] = {1, 3, 1, 2, 4, 2, 3, 5, 3, 4, 6, 4, 5, 7, 5, 6, 8};
CFX_CONST uint8_t kCqTrits[kColorLevels] = {1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0};
CFX_CONST uint8_t kCqQuints[kColorLevels] = {0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0};

#ifndef CFX_ASTC_DEC_STEPS
#define CFX_ASTC_DEC_STEPS 1
#define CFX_ASTC_DEC_GAIN 26.0f
#endif
constexpr int kDecimationSteps = CFX_ASTC_DEC_STEPS;
constexpr float kDecimationGain = CFX_ASTC_DEC_GAIN;

#ifdef CFX_COUNT_OPS
static unsigned long long g_ops[8];
#define CFX_OP(k, n) (g_ops[k] += (n))
#else
#define CFX_OP(k, n)
#endif

struct Ctx {
    const uint8_t* blob;
    AstcTab tab;
};

#ifdef __CUDA_ARCH__
#define CFX_LD(ptr) __ldg(ptr)
#else
#define CFX_LD(ptr) (*(ptr))
#endif

CFX_HD uint32_t tab_u8(const Ctx& c, uint32_t off) { return CFX_LD(c.blob + off); }
CFX_HD uint32_t tab_u16(const Ctx& c, uint32_t off) { return CFX_LD(reinterpret_cast<const uint16_t*>(c.blob + off)); }
CFX_HD float tab_f32(const Ctx& c, uint32_t off) { return CFX_LD(reinterpret_cast<const float*>(c.blob + off)); }
CFX_HD uint2 tab_u32x2(const Ctx& c, uint32_t off) { return CFX_LD(reinterpret_cast<const uint2*>(c.blob + off)); }
CFX_HD uint64_t tab_u64(const Ctx& c, uint32_t off)
{
    return CFX_LD(reinterpret_cast<const unsigned long long*>(c.blob + off));
}
CFX_HD ModeInfo tab_mode(const Ctx& c, uint32_t index)
{
    uint2 v = tab_u32x2(c, c.tab.off_modes + index*8u);
    ModeInfo m;
    m.mode_bits = static_cast<uint16_t>(v.x & 0xFFFFu); m.grid = static_cast<uint8_t>((v.x >> 16) & 0xFFu);
    m.level = static_cast<uint8_t>(v.x >> 24); m.wbits = static_cast<uint8_t>(v.y & 0xFFu);
    m.nw = static_cast<uint8_t>((v.y >> 8) & 0xFFu); m.dual = static_cast<uint8_t>((v.y >> 16) & 0xFFu); m.pad = 0;
    return m;
}

// One partitioning hypothesis of the block with its per-subset lines and ideal weights.
struct Slot {
    float t[kMaxTexels];         // ideal weight of every texel in [0,1] along its subset's line
    uint8_t part[kMaxTexels];    // subset of every texel
    float t2[kMaxTexels];        // second-plane ideal weights (dual_ch >= 0 only)
    float4 e0[3], e1[3];         // line end points (0..255 per channel), sum(e1.rgb) >= sum(e0.rgb)
    uint32_t pc, seed, valid;
    int32_t dual_ch;             // channel that gets its own weight plane, or -1
    float e_line;                // squared distance of the texels from their subsets' lines (8-bit units):
                                 // the error left with ideal weights and end points
    float len2[3];               // squared length of each subset's line (first plane)
};
constexpr int kSlots = 9;        // 0: one subset; 1,2: two subsets; 3,4: three subsets; 5..8: dual plane on R,G,B,A

// Per-lane scratch: byte j of lane l lives at word (j/4)*32 + l, so that all lanes touching the same
// j hit different banks.
CFX_HD uint32_t scr_index(uint32_t j, uint32_t lane) { return (((j >> 2)*32u + lane) << 2) | (j & 3u); }

struct Enc {
    uint32_t ep[3][2];           // quantised end points, 8-bit RGBA per byte, [subset][0 = e0, 1 = e1]
    uint32_t clevel;
    float err;
};

CFX_HD float ch(const float4& v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w)); }

CFX_HD uint32_t quant_color(const Ctx& c, uint32_t level, float v)
{
    int iv = min(max(__float2int_rn(v), 0), 255);
    uint32_t rank = tab_u8(c, c.tab.off_cq_near + level*256u + static_cast<uint32_t>(iv));
    return tab_u8(c, c.tab.off_cq_val + level*256u + rank);
}

CFX_HD void quant_endpoints(const Ctx& c, uint32_t level, bool has_alpha, const float4& a, const float4& b, uint32_t& qa,
    uint32_t& qb)
{
    qa = quant_color(c, level, a.x) | (quant_color(c, level, a.y) << 8) | (quant_color(c, level, a.z) << 16);
    qb = quant_color(c, level, b.x) | (quant_color(c, level, b.y) << 8) | (quant_color(c, level, b.z) << 16);
    if (has_alpha) { qa |= quant_color(c, level, a.w) << 24; qb |= quant_color(c, level, b.w) << 24; }
    else { qa |= 0xFF000000u; qb |= 0xFF000000u; }
    // RGB(A) direct end points decode as written only when sum(e1) >= sum(e0) (otherwise the decoder
    // applies blue contraction): keep that ordering
    uint32_t sa = (qa & 0xFF) + ((qa >> 8) & 0xFF) + ((qa >> 16) & 0xFF);
    uint32_t sb = (qb & 0xFF) + ((qb >> 8) & 0xFF) + ((qb >> 16) & 0xFF);
    if (sb < sa) { uint32_t tmp = qa; qa = qb; qb = tmp; }
}

// Squared error of the block decoded from (end points, per-texel weights in w_scr; the second plane's
// texel weights, if any, follow at w_scr[T + i]).
CFX_HD float decoded_error(const float4* cf, uint32_t T, const Slot& slot, const Enc& e, const uint8_t* w_scr, uint32_t lane,
    bool has_alpha)
{
    float err = 0.0f;
    const int dc = slot.dual_ch;
    for (uint32_t i = 0; i < T; ++i) {
        const float w = static_cast<float>(w_scr[scr_index(i, lane)])*(1.0f/64.0f);
        const float w2 = dc >= 0 ? static_cast<float>(w_scr[scr_index(T + i, lane)])*(1.0f/64.0f) : w;
        const uint32_t s = slot.part[i];
        const uint32_t a = s == 0 ? e.ep[0][0] : (s == 1 ? e.ep[1][0] : e.ep[2][0]);
        const uint32_t b = s == 0 ? e.ep[0][1] : (s == 1 ? e.ep[1][1] : e.ep[2][1]);
        const float4 x = cf[i];
        float a0 = static_cast<float>(a & 0xFF), b0 = static_cast<float>(b & 0xFF);
        float d = a0 + (b0 - a0)*(dc == 0 ? w2 : w) - x.x; err += d*d;
        a0 = static_cast<float>((a >> 8) & 0xFF); b0 = static_cast<float>((b >> 8) & 0xFF);
        d = a0 + (b0 - a0)*(dc == 1 ? w2 : w) - x.y; err += d*d;
        a0 = static_cast<float>((a >> 16) & 0xFF); b0 = static_cast<float>((b >> 16) & 0xFF);
        d = a0 + (b0 - a0)*(dc == 2 ? w2 : w) - x.z; err += d*d;
        if (has_alpha) {
            a0 = static_cast<float>(a >> 24); b0 = static_cast<float>(b >> 24);
            d = a0 + (b0 - a0)*(dc == 3 ? w2 : w) - x.w; err += d*d;
        }
    }
    return err;
}

// Grid weights for a block mode: decimate per-texel ideal weights onto the mode's grid (weighted by
// the transposed infill factors), quantise, then infill back to the texel weights the decoder will
// use.  The ideal weights come from the slot's lines (proj == nullptr) or from projecting every texel
// onto the segments between the given quantised end points (proj = ep[subset][2]).  plane 1 is the
// slot's dual channel on its own.  Grid weight j of plane p (unquantised value 0..64) goes to
// u_scr[j*planes + p] -- the order of the bit stream -- and texel weights to w_scr[p*T + i].
CFX_HD void compute_weights(const Ctx& c, const float4* cf, const Slot& slot, const ModeInfo& m, const uint32_t (*proj)[2],
    bool has_alpha, uint32_t plane, uint8_t* u_scr, uint8_t* w_scr, uint32_t lane)
{
    const uint32_t T = c.tab.texels;
    const uint32_t g = m.grid, L = m.level;
    const uint32_t planes = slot.dual_ch >= 0 ? 2u : 1u;
    const int dc = slot.dual_ch;
    const float nm1 = static_cast<float>(kWqN[L] - 1);
    const uint32_t start_off = c.tab.off_csr_start + g*(kMaxTexels + 2)*2u;
    const uint32_t ent_off = c.tab.off_csr_ent + g*4u*T*2u;
    const uint32_t norm_off = c.tab.off_wnorm + g*kMaxTexels*4u;
    // per-subset projection: t = dot(x - a, d)
    float4 pa[3], pd[3];
    if (proj) {
        for (uint32_t s = 0; s < slot.pc; ++s) {
            const uint32_t a = proj[s][0], b = proj[s][1];
            pa[s] = make_float4(static_cast<float>(a & 0xFF), static_cast<float>((a >> 8) & 0xFF),
                static_cast<float>((a >> 16) & 0xFF), static_cast<float>(a >> 24));
            float dd[4] = {static_cast<float>(b & 0xFF) - pa[s].x, static_cast<float>((b >> 8) & 0xFF) - pa[s].y,
                static_cast<float>((b >> 16) & 0xFF) - pa[s].z, has_alpha ? static_cast<float>(b >> 24) - pa[s].w : 0.0f};
#pragma unroll
            for (int k = 0; k < 4; ++k) if (dc >= 0 && ((k == dc) != (plane == 1))) dd[k] = 0.0f;
            const float l2 = dd[0]*dd[0] + dd[1]*dd[1] + dd[2]*dd[2] + dd[3]*dd[3];
            const float inv = l2 > 0.0f ? 1.0f/l2 : 0.0f;
            pd[s] = make_float4(dd[0]*inv, dd[1]*inv, dd[2]*inv, dd[3]*inv);
        }
    }
    CFX_OP(0, 1); CFX_OP(1, tab_u16(c, start_off + m.nw*2u));
    const float* tt = plane ? slot.t2 : slot.t;
    auto ideal = [&](uint32_t i) -> float {
        if (!proj) return tt[i];
        const uint32_t sub = slot.part[i];
        const float4 a = sub == 0 ? pa[0] : (sub == 1 ? pa[1] : pa[2]);
        const float4 d = sub == 0 ? pd[0] : (sub == 1 ? pd[1] : pd[2]);
        const float4 x = cf[i];
        const float t = (x.x - a.x)*d.x + (x.y - a.y)*d.y + (x.z - a.z)*d.z + (x.w - a.w)*d.w;
        return fminf(fmaxf(t, 0.0f), 1.0f);
    };
    const uint32_t inf_off = c.tab.off_infill + g*T*8u;
    const bool decimated = m.nw < T;
    // pass 1: grid weight = factor-weighted mean of the ideal weights of the texels it touches
    uint32_t e = tab_u16(c, start_off);
    for (uint32_t j = 0; j < m.nw; ++j) {
        const uint32_t end = tab_u16(c, start_off + (j + 1)*2u);
        float s = 0.0f;
        for (; e < end; ++e) {
            const uint32_t ent = tab_u16(c, ent_off + e*2u);
            s += static_cast<float>(ent >> 8)*ideal(ent & 0xFFu);
        }
        const float gj = s*tab_f32(c, norm_off + j*4u);
        if (decimated) {
            u_scr[scr_index(j*planes + plane, lane)] = static_cast<uint8_t>(__float2int_rn(fminf(fmaxf(gj, 0.0f), 1.0f)*255.0f));
        } else {
            const int k = min(max(__float2int_rn(gj*nm1), 0), static_cast<int>(kWqN[L]) - 1);
            u_scr[scr_index(j*planes + plane, lane)] =
                static_cast<uint8_t>(tab_u8(c, c.tab.off_wq_val + L*32u + static_cast<uint32_t>(k)));
        }
    }
    for (int it = 0; decimated && it < kDecimationSteps; ++it) {
        // what the texels would get back from the current grid weights (8-bit fixed point)
        for (uint32_t i = 0; i < T; ++i) {
            const uint2 inf = tab_u32x2(c, inf_off + i*8u);
            uint32_t acc = 8;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                acc += ((inf.y >> (8*k)) & 0xFFu)*u_scr[scr_index(((inf.x >> (8*k)) & 0xFFu)*planes + plane, lane)];
            w_scr[scr_index(plane*T + i, lane)] = static_cast<uint8_t>(acc >> 4);
        }
        // a least-squares (Jacobi) step on the grid weights against the infill residual; the last one quantises
        const bool last = it + 1 == kDecimationSteps;
        e = tab_u16(c, start_off);
        for (uint32_t j = 0; j < m.nw; ++j) {
            const uint32_t end = tab_u16(c, start_off + (j + 1)*2u);
            float s = 0.0f, s2 = 0.0f;
            for (; e < end; ++e) {
                const uint32_t ent = tab_u16(c, ent_off + e*2u);
                const uint32_t i = ent & 0xFFu;
                const float f = static_cast<float>(ent >> 8);
                s += f*(ideal(i) - static_cast<float>(w_scr[scr_index(plane*T + i, lane)])*(1.0f/255.0f));
                s2 += f*f;
            }
            const float gj = static_cast<float>(u_scr[scr_index(j*planes + plane, lane)])*(1.0f/255.0f) +
                (s2 > 0.0f ? kDecimationGain*s/s2 : 0.0f);
            if (last) {
                const int k = min(max(__float2int_rn(gj*nm1), 0), static_cast<int>(kWqN[L]) - 1);
                u_scr[scr_index(j*planes + plane, lane)] =
                    static_cast<uint8_t>(tab_u8(c, c.tab.off_wq_val + L*32u + static_cast<uint32_t>(k)));
            } else {
                u_scr[scr_index(j*planes + plane, lane)] = static_cast<uint8_t>(__float2int_rn(fminf(fmaxf(gj, 0.0f), 1.0f)*255.0f));
            }
        }
    }
    for (uint32_t i = 0; i < T; ++i) {
        const uint2 inf = tab_u32x2(c, inf_off + i*8u);
        uint32_t acc = 8;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            acc += ((inf.y >> (8*k)) & 0xFFu)*u_scr[scr_index(((inf.x >> (8*k)) & 0xFFu)*planes + plane, lane)];
        w_scr[scr_index(plane*T + i, lane)] = static_cast<uint8_t>(acc >> 4);
    }
}

CFX_HD void compute_all_weights(const Ctx& c, const float4* cf, const Slot& slot, const ModeInfo& m,
    const uint32_t (*proj)[2], bool has_alpha, uint8_t* u_scr, uint8_t* w_scr, uint32_t lane)
{
    compute_weights(c, cf, slot, m, proj, has_alpha, 0, u_scr, w_scr, lane);
    if (slot.dual_ch >= 0) compute_weights(c, cf, slot, m, proj, has_alpha, 1, u_scr, w_scr, lane);
}

// Least-squares end points of every subset for the texel weights in w_scr, quantised at e.clevel.
CFX_HD void solve_endpoints(const Ctx& c, const float4* cf, const Slot& slot, bool has_alpha, const uint8_t* w_scr,
    uint32_t lane, Enc& e)
{
    const uint32_t T = c.tab.texels;
    const int dc = slot.dual_ch;
    for (uint32_t s = 0; s < slot.pc; ++s) {
        float A = 0, B = 0, C = 0, P[4] = {0, 0, 0, 0}, Q[4] = {0, 0, 0, 0};
        float A2 = 0, B2 = 0, C2 = 0, P2 = 0, Q2 = 0;        // the dual channel's own system
        for (uint32_t i = 0; i < T; ++i) {
            if (slot.part[i] != s) continue;
            const float w = static_cast<float>(w_scr[scr_index(i, lane)])*(1.0f/64.0f), iw = 1.0f - w;
            const float4 x = cf[i];
            A += iw*iw; B += iw*w; C += w*w;
            P[0] += iw*x.x; P[1] += iw*x.y; P[2] += iw*x.z; P[3] += iw*x.w;
            Q[0] += w*x.x; Q[1] += w*x.y; Q[2] += w*x.z; Q[3] += w*x.w;
            if (dc >= 0) {
                const float v = static_cast<float>(w_scr[scr_index(T + i, lane)])*(1.0f/64.0f), iv = 1.0f - v;
                const float xc = ch(x, dc);
                A2 += iv*iv; B2 += iv*v; C2 += v*v; P2 += iv*xc; Q2 += v*xc;
            }
        }
        const float det = A*C - B*B;
        if (fabsf(det) < 1e-4f*(A + C)*(A + C) + 1e-12f) continue;
        const float id = 1.0f/det;
        float a[4] = {(C*P[0] - B*Q[0])*id, (C*P[1] - B*Q[1])*id, (C*P[2] - B*Q[2])*id, (C*P[3] - B*Q[3])*id};
        float b[4] = {(A*Q[0] - B*P[0])*id, (A*Q[1] - B*P[1])*id, (A*Q[2] - B*P[2])*id, (A*Q[3] - B*P[3])*id};
        if (dc >= 0) {
            const float det2 = A2*C2 - B2*B2;
            const uint32_t oa = (e.ep[s][0] >> (8*dc)) & 0xFFu, ob = (e.ep[s][1] >> (8*dc)) & 0xFFu;
            float na = static_cast<float>(oa), nb = static_cast<float>(ob);
            if (fabsf(det2) >= 1e-4f*(A2 + C2)*(A2 + C2) + 1e-12f) {
                na = (C2*P2 - B2*Q2)/det2; nb = (A2*Q2 - B2*P2)/det2;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) if (k == dc) { a[k] = na; b[k] = nb; }
        }
        quant_endpoints(c, e.clevel, has_alpha, make_float4(a[0], a[1], a[2], a[3]), make_float4(b[0], b[1], b[2], b[3]),
            e.ep[s][0], e.ep[s][1]);
    }
}

// Evaluate one (slot, block mode) candidate with exact decoded error.  refine > 0 then alternates
// least-squares end points for the actual (decimated, quantised) weights with re-derived weights
// for the new end points, keeping every step only if the decoded error drops.
CFX_HD void evaluate(const Ctx& c, const float4* cf, const Slot& slot, const ModeInfo& m, bool has_alpha, uint8_t* u_scr,
    uint8_t* w_scr, uint32_t lane, int refine, Enc& out)
{
    CFX_OP(2, 1);
    const uint32_t T = c.tab.texels;
    out.err = 3.0e38f;
    const uint32_t pc = slot.pc;
    const uint32_t n_ints = pc*(has_alpha ? 8u : 6u);
    const int avail = 128 - static_cast<int>(m.wbits) - (pc == 1 ? 17 : 29) - (slot.dual_ch >= 0 ? 2 : 0);
    if (n_ints > 18 || avail < 0 || (m.dual != 0) != (slot.dual_ch >= 0)) return;
    const uint32_t cl = tab_u8(c, c.tab.off_clevel + (n_ints >> 1)*128u + static_cast<uint32_t>(avail));
    if (cl == 0xFF) return;
    out.clevel = cl;
    compute_all_weights(c, cf, slot, m, nullptr, has_alpha, u_scr, w_scr, lane);
    for (uint32_t s = 0; s < pc; ++s) quant_endpoints(c, cl, has_alpha, slot.e0[s], slot.e1[s], out.ep[s][0], out.ep[s][1]);
    out.err = decoded_error(cf, T, slot, out, w_scr, lane, has_alpha);

    if (refine < 0) {
        // ranking pass: one least-squares solve so that candidates are compared at their fitted error
        Enc trial = out;
        solve_endpoints(c, cf, slot, has_alpha, w_scr, lane, trial);
        trial.err = decoded_error(cf, T, slot, trial, w_scr, lane, has_alpha);
        if (trial.err < out.err) out = trial;
        return;
    }
    bool from_slot = true;          // where the weights in scratch came from (to redo them on a failed trial)
    uint32_t src[3][2];
    for (int r = 0; r < refine; ++r) {
        Enc trial = out;
        solve_endpoints(c, cf, slot, has_alpha, w_scr, lane, trial);
        trial.err = decoded_error(cf, T, slot, trial, w_scr, lane, has_alpha);
        const bool better_ep = trial.err < out.err;
        if (better_ep) out = trial;
        // new weights for the (possibly) new end points
        compute_all_weights(c, cf, slot, m, out.ep, has_alpha, u_scr, w_scr, lane);
        const float e2 = decoded_error(cf, T, slot, out, w_scr, lane, has_alpha);
        if (e2 < out.err) {
            out.err = e2; from_slot = false;
            for (uint32_t s = 0; s < pc; ++s) { src[s][0] = out.ep[s][0]; src[s][1] = out.ep[s][1]; }
        } else {
            compute_all_weights(c, cf, slot, m, from_slot ? nullptr : src, has_alpha, u_scr, w_scr, lane);
            if (!better_ep) break;
        }
    }
    if (refine > 0 && m.nw == T && slot.dual_ch < 0) {
        // one weight per texel: pick, per texel, the quantised weight with the smallest decoded error,
        // then re-solve the end points for those weights
        const uint32_t L = m.level;
        for (uint32_t i = 0; i < T; ++i) {
            const uint32_t s = slot.part[i];
            const uint32_t a = s == 0 ? out.ep[0][0] : (s == 1 ? out.ep[1][0] : out.ep[2][0]);
            const uint32_t b = s == 0 ? out.ep[0][1] : (s == 1 ? out.ep[1][1] : out.ep[2][1]);
            const float4 x = cf[i];
            float beste = 3.0e38f;
            uint32_t bestu = 0;
            for (uint32_t k = 0; k < kWqN[L]; ++k) {
                const uint32_t u = tab_u8(c, c.tab.off_wq_val + L*32u + k);
                const float w = static_cast<float>(u)*(1.0f/64.0f);
                float a0 = static_cast<float>(a & 0xFF), b0 = static_cast<float>(b & 0xFF);
                float d = a0 + (b0 - a0)*w - x.x, err = d*d;
                a0 = static_cast<float>((a >> 8) & 0xFF); b0 = static_cast<float>((b >> 8) & 0xFF);
                d = a0 + (b0 - a0)*w - x.y; err += d*d;
                a0 = static_cast<float>((a >> 16) & 0xFF); b0 = static_cast<float>((b >> 16) & 0xFF);
                d = a0 + (b0 - a0)*w - x.z; err += d*d;
                if (has_alpha) { a0 = static_cast<float>(a >> 24); b0 = static_cast<float>(b >> 24); d = a0 + (b0 - a0)*w - x.w; err += d*d; }
                if (err < beste) { beste = err; bestu = u; }
            }
            u_scr[scr_index(i, lane)] = static_cast<uint8_t>(bestu);
            w_scr[scr_index(i, lane)] = static_cast<uint8_t>(bestu);
        }
        out.err = decoded_error(cf, T, slot, out, w_scr, lane, has_alpha);
        Enc trial = out;
        solve_endpoints(c, cf, slot, has_alpha, w_scr, lane, trial);
        trial.err = decoded_error(cf, T, slot, trial, w_scr, lane, has_alpha);
        if (trial.err < out.err) out = trial;
    }
}

// ---- building a slot: per-subset mean, principal axis, projections ----------------------------
CFX_HD void build_slot(const float4* cf, uint32_t T, bool has_alpha, Slot& slot)
{
    slot.e_line = 0.0f;
    for (uint32_t s = 0; s < slot.pc; ++s) {
        float n = 0, m[4] = {0, 0, 0, 0};
        for (uint32_t i = 0; i < T; ++i) {
            if (slot.part[i] != s) continue;
            const float4 x = cf[i];
            n += 1.0f; m[0] += x.x; m[1] += x.y; m[2] += x.z; m[3] += x.w;
        }
        const float inv = n > 0 ? 1.0f/n : 0.0f;
        m[0] *= inv; m[1] *= inv; m[2] *= inv; m[3] *= inv;
        float cv[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (uint32_t i = 0; i < T; ++i) {
            if (slot.part[i] != s) continue;
            const float4 x = cf[i];
            float d0 = x.x - m[0], d1 = x.y - m[1], d2 = x.z - m[2], d3 = has_alpha ? x.w - m[3] : 0.0f;
            if (slot.dual_ch == 0) d0 = 0.0f;
            if (slot.dual_ch == 1) d1 = 0.0f;
            if (slot.dual_ch == 2) d2 = 0.0f;
            if (slot.dual_ch == 3) d3 = 0.0f;
            cv[0] += d0*d0; cv[1] += d0*d1; cv[2] += d0*d2; cv[3] += d0*d3; cv[4] += d1*d1;
            cv[5] += d1*d2; cv[6] += d1*d3; cv[7] += d2*d2; cv[8] += d2*d3; cv[9] += d3*d3;
        }
        // principal axis by power iteration from the strongest row
        float v[4] = {cv[0], cv[1], cv[2], cv[3]};
        float best = cv[0];
        if (cv[4] > best) { best = cv[4]; v[0] = cv[1]; v[1] = cv[4]; v[2] = cv[5]; v[3] = cv[6]; }
        if (cv[7] > best) { best = cv[7]; v[0] = cv[2]; v[1] = cv[5]; v[2] = cv[7]; v[3] = cv[8]; }
        if (cv[9] > best) { best = cv[9]; v[0] = cv[3]; v[1] = cv[6]; v[2] = cv[8]; v[3] = cv[9]; }
        for (int it = 0; it < 6; ++it) {
            const float n2 = v[0]*v[0] + v[1]*v[1] + v[2]*v[2] + v[3]*v[3];
            const float s2 = n2 > 1e-20f ? rsqrtf(n2) : 0.0f;
            const float a0 = v[0]*s2, a1 = v[1]*s2, a2 = v[2]*s2, a3 = v[3]*s2;
            v[0] = cv[0]*a0 + cv[1]*a1 + cv[2]*a2 + cv[3]*a3;
            v[1] = cv[1]*a0 + cv[4]*a1 + cv[5]*a2 + cv[6]*a3;
            v[2] = cv[2]*a0 + cv[5]*a1 + cv[7]*a2 + cv[8]*a3;
            v[3] = cv[3]*a0 + cv[6]*a1 + cv[8]*a2 + cv[9]*a3;
        }
        {
            const float n2 = v[0]*v[0] + v[1]*v[1] + v[2]*v[2] + v[3]*v[3];
            const float s2 = n2 > 1e-20f ? rsqrtf(n2) : 0.0f;
            v[0] *= s2; v[1] *= s2; v[2] *= s2; v[3] *= s2;
            if (n2 <= 1e-20f) { v[0] = v[1] = v[2] = 0.57735f; v[3] = 0.0f; }
            if (v[0] + v[1] + v[2] < 0.0f) { v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2]; v[3] = -v[3]; }
        }
        float tmin = 1e30f, tmax = -1e30f, proj2 = 0.0f;
        for (uint32_t i = 0; i < T; ++i) {
            if (slot.part[i] != s) continue;
            const float4 x = cf[i];
            const float t = (x.x - m[0])*v[0] + (x.y - m[1])*v[1] + (x.z - m[2])*v[2] + (x.w - m[3])*v[3];
            slot.t[i] = t;
            proj2 += t*t;
            tmin = fminf(tmin, t); tmax = fmaxf(tmax, t);
        }
        slot.e_line += fmaxf(cv[0] + cv[4] + cv[7] + cv[9] - proj2, 0.0f);
        if (!(tmax > tmin)) { tmin = 0.0f; tmax = 0.0f; }
        const float range = tmax - tmin;
        slot.len2[s] = range*range;
        const float ir = range > 1e-6f ? 1.0f/range : 0.0f;
        for (uint32_t i = 0; i < T; ++i)
            if (slot.part[i] == s) slot.t[i] = (slot.t[i] - tmin)*ir;
        slot.e0[s] = make_float4(m[0] + tmin*v[0], m[1] + tmin*v[1], m[2] + tmin*v[2], m[3] + tmin*v[3]);
        slot.e1[s] = make_float4(m[0] + tmax*v[0], m[1] + tmax*v[1], m[2] + tmax*v[2], m[3] + tmax*v[3]);
        if (slot.dual_ch >= 0) {
            // the dual channel runs on its own plane from its minimum to its maximum
            float lo = 1e30f, hi = -1e30f;
            for (uint32_t i = 0; i < T; ++i) { const float xc = ch(cf[i], slot.dual_ch); lo = fminf(lo, xc); hi = fmaxf(hi, xc); }
            const float ir2 = hi - lo > 1e-6f ? 1.0f/(hi - lo) : 0.0f;
            for (uint32_t i = 0; i < T; ++i) slot.t2[i] = (ch(cf[i], slot.dual_ch) - lo)*ir2;
            float a[4] = {slot.e0[s].x, slot.e0[s].y, slot.e0[s].z, slot.e0[s].w};
            float b[4] = {slot.e1[s].x, slot.e1[s].y, slot.e1[s].z, slot.e1[s].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) if (k == slot.dual_ch) { a[k] = lo; b[k] = hi; }
            slot.e0[s] = make_float4(a[0], a[1], a[2], a[3]);
            slot.e1[s] = make_float4(b[0], b[1], b[2], b[3]);
        }
    }
    slot.valid = 1;
}

CFX_HD void fill_parts(Slot& slot, uint32_t T, uint32_t pc, uint32_t seed, uint64_t m1, uint64_t m2)
{
    slot.pc = pc; slot.seed = seed; slot.valid = 0; slot.dual_ch = -1;
    for (uint32_t i = 0; i < T; ++i)
        slot.part[i] = static_cast<uint8_t>(((m1 >> i) & 1u) ? 1u : (((m2 >> i) & 1u) ? 2u : 0u));
}

// ---- partition search helpers -----------------------------------------------------------------
// Two-/three-means clustering of the texels (a few Lloyd iterations from spread-out seeds): the
// resulting texel masks are what candidate partition seeds are matched against.
CFX_HD void kmeans(const float4* cf, uint32_t T, uint32_t k, uint64_t* masks /* k-1 masks: cluster 1, cluster 2 */)
{
    float4 ctr[3];
    // seeds: texel 0's farthest texel, then the texel farthest from both, ...
    float4 mean = make_float4(0, 0, 0, 0);
    for (uint32_t i = 0; i < T; ++i) { mean.x += cf[i].x; mean.y += cf[i].y; mean.z += cf[i].z; mean.w += cf[i].w; }
    const float inv = 1.0f/static_cast<float>(T);
    mean.x *= inv; mean.y *= inv; mean.z *= inv; mean.w *= inv;
    ctr[0] = mean;
    for (uint32_t c = 0; c < k; ++c) {
        float bestd = -1.0f;
        uint32_t besti = 0;
        for (uint32_t i = 0; i < T; ++i) {
            float d = 3.0e38f;
            for (uint32_t p = 0; p < (c == 0 ? 1u : c); ++p) {
                const float dx = cf[i].x - ctr[p].x, dy = cf[i].y - ctr[p].y, dz = cf[i].z - ctr[p].z, dw = cf[i].w - ctr[p].w;
                d = fminf(d, dx*dx + dy*dy + dz*dz + dw*dw);
            }
            if (d > bestd) { bestd = d; besti = i; }
        }
        ctr[c] = cf[besti];
    }
    uint64_t m1 = 0, m2 = 0;
    for (int it = 0; it < 3; ++it) {
        float4 sum[3] = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
        float cnt[3] = {0, 0, 0};
        m1 = m2 = 0;
        for (uint32_t i = 0; i < T; ++i) {
            uint32_t bi = 0;
            float bd = 3.0e38f;
            for (uint32_t p = 0; p < k; ++p) {
                const float dx = cf[i].x - ctr[p].x, dy = cf[i].y - ctr[p].y, dz = cf[i].z - ctr[p].z, dw = cf[i].w - ctr[p].w;
                const float d = dx*dx + dy*dy + dz*dz + dw*dw;
                if (d < bd) { bd = d; bi = p; }
            }
            if (bi == 1) m1 |= 1ull << i;
            if (bi == 2) m2 |= 1ull << i;
            // branch-free accumulate into the chosen cluster
            for (uint32_t p = 0; p < k; ++p) {
                const float f = p == bi ? 1.0f : 0.0f;
                sum[p].x += f*cf[i].x; sum[p].y += f*cf[i].y; sum[p].z += f*cf[i].z; sum[p].w += f*cf[i].w; cnt[p] += f;
            }
        }
        for (uint32_t p = 0; p < k; ++p)
            if (cnt[p] > 0) {
                const float ic = 1.0f/cnt[p];
                ctr[p] = make_float4(sum[p].x*ic, sum[p].y*ic, sum[p].z*ic, sum[p].w*ic);
            }
    }
    masks[0] = m1;
    if (k > 2) masks[1] = m2;
}

CFX_HD uint32_t popc64(uint64_t v) { return static_cast<uint32_t>(__popc(static_cast<uint32_t>(v)) + __popc(static_cast<uint32_t>(v >> 32))); }

// Number of texels a 2-subset seed assigns differently from the clustering (up to relabelling).
CFX_HD uint32_t mismatch2(uint64_t km, uint64_t pm, uint32_t T)
{
    const uint32_t d = popc64(km ^ pm);
    return min(d, T - d);
}

// Same for three subsets: best of the 6 label permutations.
CFX_HD uint32_t mismatch3(uint64_t k1, uint64_t k2, uint64_t p1, uint64_t p2, uint64_t full)
{
    const uint64_t k0 = full ^ k1 ^ k2, p0 = full ^ p1 ^ p2;
    const uint64_t kk[3] = {k0, k1, k2}, pp[3] = {p0, p1, p2};
    uint32_t m[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) m[i][j] = popc64(kk[i] & pp[j]);
    // texels that agree under each permutation; mismatch = T - max agreement
    uint32_t a = m[0][0] + m[1][1] + m[2][2];
    a = max(a, m[0][0] + m[1][2] + m[2][1]);
    a = max(a, m[0][1] + m[1][0] + m[2][2]);
    a = max(a, m[0][1] + m[1][2] + m[2][0]);
    a = max(a, m[0][2] + m[1][0] + m[2][1]);
    a = max(a, m[0][2] + m[1][1] + m[2][0]);
    return popc64(full) - a;
}

// Sum over subsets of (trace - lambda_max) of the scatter matrix = squared distance of the texels
// from their subsets' best-fit lines: the exact figure of merit for a partitioning.
CFX_HD float line_fit_residual(const float4* cf, uint32_t T, uint32_t pc, uint64_t m1, uint64_t m2, bool has_alpha)
{
    float total = 0.0f;
    for (uint32_t s = 0; s < pc; ++s) {
        const uint64_t mask = s == 0 ? ~(m1 | m2) : (s == 1 ? m1 : m2);
        float n = 0, sm[4] = {0, 0, 0, 0}, cv[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (uint32_t i = 0; i < T; ++i) {
            const float f = ((mask >> i) & 1u) ? 1.0f : 0.0f;
            const float4 x = cf[i];
            const float xw = has_alpha ? x.w : 0.0f;
            n += f; sm[0] += f*x.x; sm[1] += f*x.y; sm[2] += f*x.z; sm[3] += f*xw;
            const float fx = f*x.x, fy = f*x.y, fz = f*x.z;
            cv[0] += fx*x.x; cv[1] += fx*x.y; cv[2] += fx*x.z; cv[3] += fx*xw; cv[4] += fy*x.y;
            cv[5] += fy*x.z; cv[6] += fy*xw; cv[7] += fz*x.z; cv[8] += fz*xw; cv[9] += f*xw*xw;
        }
        if (n < 1.0f) return 3.0e38f;
        const float inv = 1.0f/n;
        cv[0] -= sm[0]*sm[0]*inv; cv[1] -= sm[0]*sm[1]*inv; cv[2] -= sm[0]*sm[2]*inv; cv[3] -= sm[0]*sm[3]*inv;
        cv[4] -= sm[1]*sm[1]*inv; cv[5] -= sm[1]*sm[2]*inv; cv[6] -= sm[1]*sm[3]*inv;
        cv[7] -= sm[2]*sm[2]*inv; cv[8] -= sm[2]*sm[3]*inv; cv[9] -= sm[3]*sm[3]*inv;
        float v[4] = {cv[0], cv[1], cv[2], cv[3]};
        float best = cv[0];
        if (cv[4] > best) { best = cv[4]; v[0] = cv[1]; v[1] = cv[4]; v[2] = cv[5]; v[3] = cv[6]; }
        if (cv[7] > best) { best = cv[7]; v[0] = cv[2]; v[1] = cv[5]; v[2] = cv[7]; v[3] = cv[8]; }
        if (cv[9] > best) { best = cv[9]; v[0] = cv[3]; v[1] = cv[6]; v[2] = cv[8]; v[3] = cv[9]; }
        float lam = 0.0f;
        for (int it = 0; it < 4; ++it) {
            const float n2 = v[0]*v[0] + v[1]*v[1] + v[2]*v[2] + v[3]*v[3];
            const float s2 = n2 > 1e-20f ? rsqrtf(n2) : 0.0f;
            const float a0 = v[0]*s2, a1 = v[1]*s2, a2 = v[2]*s2, a3 = v[3]*s2;
            v[0] = cv[0]*a0 + cv[1]*a1 + cv[2]*a2 + cv[3]*a3;
            v[1] = cv[1]*a0 + cv[4]*a1 + cv[5]*a2 + cv[6]*a3;
            v[2] = cv[2]*a0 + cv[5]*a1 + cv[7]*a2 + cv[8]*a3;
            v[3] = cv[3]*a0 + cv[6]*a1 + cv[8]*a2 + cv[9]*a3;
            lam = a0*v[0] + a1*v[1] + a2*v[2] + a3*v[3];
        }
        total += fmaxf((cv[0] + cv[4] + cv[7] + cv[9]) - lam, 0.0f);
    }
    return total;
}

// ---- bit packing ------------------------------------------------------------------------------
struct Bits128 {
    uint64_t lo, hi;
    CFX_HD void put(uint32_t pos, uint32_t v, uint32_t n)
    {
        if (n == 0) return;
        const uint64_t vv = static_cast<uint64_t>(v) & ((1ull << n) - 1ull);
        if (pos < 64) {
            lo |= vv << pos;
            if (pos + n > 64) hi |= vv >> (64u - pos);
        } else {
            hi |= vv << (pos - 64u);
        }
    }
};

// BISE-encode n values (read through `get(i)`) at bit position pos; returns the end position.
template <typename Get>
CFX_HD uint32_t ise_encode(const Ctx& c, Bits128& out, uint32_t pos, uint32_t n, uint32_t bits, bool trits, bool quints, Get get)
{
    const uint32_t mask = (1u << bits) - 1u;
    if (trits) {
        for (uint32_t i = 0; i < n; i += 5) {
            uint32_t v[5] = {0, 0, 0, 0, 0};
            for (uint32_t k = 0; k < 5 && i + k < n; ++k) v[k] = get(i + k);
            const uint32_t T = tab_u8(c, c.tab.off_trit_enc + (v[0] >> bits) + 3u*(v[1] >> bits) + 9u*(v[2] >> bits) +
                27u*(v[3] >> bits) + 81u*(v[4] >> bits));
            const uint32_t tshift[5] = {0, 2, 4, 5, 7}, tbits[5] = {2, 2, 1, 2, 1};
            for (uint32_t k = 0; k < 5 && i + k < n; ++k) {
                out.put(pos, v[k] & mask, bits); pos += bits;
                out.put(pos, T >> tshift[k], tbits[k]); pos += tbits[k];
            }
        }
    } else if (quints) {
        for (uint32_t i = 0; i < n; i += 3) {
            uint32_t v[3] = {0, 0, 0};
            for (uint32_t k = 0; k < 3 && i + k < n; ++k) v[k] = get(i + k);
            const uint32_t Q = tab_u8(c, c.tab.off_quint_enc + (v[0] >> bits) + 5u*(v[1] >> bits) + 25u*(v[2] >> bits));
            const uint32_t qshift[3] = {0, 3, 5}, qbits[3] = {3, 2, 2};
            for (uint32_t k = 0; k < 3 && i + k < n; ++k) {
                out.put(pos, v[k] & mask, bits); pos += bits;
                out.put(pos, Q >> qshift[k], qbits[k]); pos += qbits[k];
            }
        }
    } else {
        for (uint32_t i = 0; i < n; ++i) { out.put(pos, get(i), bits); pos += bits; }
    }
    return pos;
}

CFX_HD uint64_t bitrev64(uint64_t v)
{
    v = ((v >> 1) & 0x5555555555555555ull) | ((v & 0x5555555555555555ull) << 1);
    v = ((v >> 2) & 0x3333333333333333ull) | ((v & 0x3333333333333333ull) << 2);
    v = ((v >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((v & 0x0F0F0F0F0F0F0F0Full) << 4);
    v = ((v >> 8) & 0x00FF00FF00FF00FFull) | ((v & 0x00FF00FF00FF00FFull) << 8);
    v = ((v >> 16) & 0x0000FFFF0000FFFFull) | ((v & 0x0000FFFF0000FFFFull) << 16);
    return (v >> 32) | (v << 32);
}

// Physical 128-bit block for (slot, mode, encoding) with grid weights in u_scr.
// u_scr: grid weights in bit-stream order, lane-interleaved scratch (linear == false) or a plain array.
template <typename SlotT>
CFX_HD_NOINLINE uint4 pack_block(const Ctx& c, const SlotT& slot, const ModeInfo& m, const Enc& e, bool has_alpha,
    const uint8_t* u_scr, uint32_t lane, bool linear = false, const uint8_t* k_lin = nullptr, bool lum = false,
    const int* hdr_vals = nullptr, const uint8_t* cems = nullptr, const int* scales = nullptr, uint32_t contracted = 0,
    const int* raw_vals = nullptr)
{
    Bits128 b; b.lo = b.hi = 0;
    const uint32_t pc = slot.pc;
    // colour end point mode: 8 / 12 = LDR RGB / RGBA direct; 0 = LDR luminance direct (lum: opaque)
    // 11 / 14 = HDR RGB direct (+ LDR alpha); hdr_vals: per subset (stride 8) the six packed values of astc_hdr.cuh
    // followed by the two alpha end points.
    // cems (per subset, optional): 6 = RGB base + scale (R G B s), 10 = base + scale + alpha pair, 8, 12; the scale of
    // subset s is scales[s]. Subsets may differ (classes of adjacent numbers only: 6 with 8, 8 with 12).
    // contracted (bit per subset) + raw_vals: the subset's RGB(A) direct end points are stored blue-contracted; raw_vals holds
    // its values v0..v7 (stride 8) as they go into the block.
    uint32_t cem[4];
    bool same = true;
    for (uint32_t s = 0; s < pc; ++s) {
        cem[s] = cems ? cems[s] : (hdr_vals ? (has_alpha ? 14u : 11u) : (lum ? (has_alpha ? 4u : 0u) : (has_alpha ? 12u : 8u)));
        same = same && cem[s] == cem[0];
    }
    b.put(0, m.mode_bits, 11);
    b.put(11, pc - 1, 2);
    uint32_t pos;
    uint32_t below = 128u - m.wbits;                  // first free bit below the weights
    if (pc == 1) { b.put(13, cem[0], 4); pos = 17; }
    else {
        b.put(13, slot.seed, 10);
        if (same) { b.put(23, 0, 2); b.put(25, cem[0], 4); }
        else {
            // ASTC spec, "Color Endpoint Mode" for multi-partition blocks: 2 bits = lowest class + 1, one class bit
            // per partition, two low bits of the mode per partition; what exceeds 6 bits sits just below the weights
            uint32_t low = 4;
            for (uint32_t s = 0; s < pc; ++s) low = low < (cem[s] >> 2) ? low : (cem[s] >> 2);
            if (low == 3) low = 2;
            uint32_t enc = low + 1u, bp = 2;
            for (uint32_t s = 0; s < pc; ++s) enc |= ((cem[s] >> 2) - low) << bp++;
            for (uint32_t s = 0; s < pc; ++s) { enc |= (cem[s] & 3u) << bp; bp += 2; }
            const uint32_t hi_bits = 3u*pc - 4u;
            b.put(23, enc & 0x3Fu, 6);
            below -= hi_bits;
            b.put(below, enc >> 6, hi_bits);
        }
        pos = 29;
    }
    uint32_t start[5];                                  // first value of every subset
    start[0] = 0;
    for (uint32_t s = 0; s < pc; ++s) start[s + 1] = start[s] + ((cem[s] >> 2) + 1u)*2u;
    const uint32_t cl = e.clevel;
    ise_encode(c, b, pos, start[pc], kCqBits[cl], kCqTrits[cl] != 0, kCqQuints[cl] != 0, [&](uint32_t i) {
        uint32_t s = 0;
        while (s + 1u < pc && i >= start[s + 1]) ++s;
        const uint32_t k = i - start[s];
        uint32_t val;
        if (hdr_vals) val = static_cast<uint32_t>(hdr_vals[s*8u + k]) & 0xFFu;
        else if ((contracted >> s) & 1u) val = static_cast<uint32_t>(raw_vals[s*8u + k]) & 0xFFu;      // blue-contracted CEM 8 / 12: v0..v7 as stored
        else if (cem[s] == 6u || cem[s] == 10u)       // R G B s (a0 a1)
            val = k < 3u ? (e.ep[s][1] >> (8u*k)) & 0xFFu : (k == 3u ? static_cast<uint32_t>(scales[s]) & 0xFFu : e.ep[s][k & 1u] >> 24);
        else if (cem[s] == 4u) val = k < 2u ? e.ep[s][k] & 0xFFu : e.ep[s][k & 1u] >> 24;      // L0 L1 A0 A1
        else val = (e.ep[s][k & 1u] >> (8u*(k >> 1))) & 0xFFu;       // r0 r1 g0 g1 b0 b1 a0 a1 (luminance: the r pair)
        const uint32_t rank = tab_u8(c, c.tab.off_cq_near + cl*256u + val);
        return tab_u8(c, c.tab.off_cq_enc + cl*256u + rank);
    });
    // weights: BISE from bit 0 of a scratch word, then mirrored into the top of the block
    Bits128 w; w.lo = w.hi = 0;
    const uint32_t L = m.level;
    const uint32_t planes = slot.dual_ch >= 0 ? 2u : 1u;
    if (planes == 2) b.put(below - 2u, static_cast<uint32_t>(slot.dual_ch), 2);
    ise_encode(c, w, 0, m.nw*planes, kWqBits[L], kWqTrits[L] != 0, kWqQuints[L] != 0, [&](uint32_t j) {
        if (k_lin) return tab_u8(c, c.tab.off_wq_enc + L*32u + k_lin[j]);    // the caller kept the ranks
        const uint32_t u = linear ? u_scr[j] : u_scr[scr_index(j, lane)];
        uint32_t k = 0;
        while (k + 1u < kWqN[L] && tab_u8(c, c.tab.off_wq_val + L*32u + k) != u) ++k;
        return tab_u8(c, c.tab.off_wq_enc + L*32u + k);
    });
    b.hi |= bitrev64(w.lo);
    b.lo |= bitrev64(w.hi);
    return make_uint4(static_cast<uint32_t>(b.lo), static_cast<uint32_t>(b.lo >> 32), static_cast<uint32_t>(b.hi),
        static_cast<uint32_t>(b.hi >> 32));
}

// Void-extent block: one UNORM16 colour for the whole block (texel values are 0..255 floats).
CFX_HD uint4 pack_void_extent(float4 v)
{
    const uint32_t r = static_cast<uint32_t>(__float2int_rn(v.x*257.0f)), g = static_cast<uint32_t>(__float2int_rn(v.y*257.0f));
    const uint32_t b = static_cast<uint32_t>(__float2int_rn(v.z*257.0f)), a = static_cast<uint32_t>(__float2int_rn(v.w*257.0f));
    return make_uint4(0xFFFFFDFCu, 0xFFFFFFFFu, r | (g << 16), b | (a << 16));
}

} // namespace astc
} // namespace cfx

// ---- block-level driver state and steps (a warp on the device, a loop over lanes on the host) ----
namespace cfx {
namespace astc {

struct BlockState {
    float4 cf[kMaxTexels];       // texels, 0..255 per channel
    Slot slots[kSlots];          // 0: one subset; 1,2: two-subset picks; 3,4: three-subset picks
    uint64_t km[3];              // k-means masks: [0] two clusters; [1],[2] three clusters
    uint32_t keys[2][32];        // per lane: best (mismatch << 10 | seed) for 2 and 3 subsets
    float scores[2][32];         // per lane: line-fit residual of that seed
    uint32_t has_alpha, constant, first_rgba8;
};

// Which search effort a quality level buys.
struct Plan {
    uint32_t n_cand[4];          // candidates evaluated per slot of each type (1 / 2 / 3 subsets, dual plane)
    uint32_t refine;             // refinement rounds on each lane's best candidate
    uint32_t slots;              // slots searched: 1 = one subset only; 5 = + two/three subsets; 9 = + dual plane
};

CFX_HD uint32_t slot_type(uint32_t slot) { return slot == 0 ? 0u : (slot < 3 ? 1u : (slot < 5 ? 2u : 3u)); }

// The search plan of a quality level (AstcConverter's preset mapping, lib/src/AstcConverter.cpp:174-195:
// Lowest/Low/Normal/High/Highest -> fastest/fast/medium/thorough/exhaustive).
inline Plan make_plan(uint32_t quality, const AstcTab& t)
{
    Plan p;
    for (int i = 0; i < 4; ++i) p.n_cand[i] = kPlanCounts[quality][i] < t.n_cand[i] ? kPlanCounts[quality][i] : t.n_cand[i];
    p.refine = quality >= 3 ? 3 : 2;
    p.slots = 9;
    return p;
}

// step 1: lane 0 builds the single-subset slot, lanes 1/2 cluster the texels into 2/3 groups
CFX_HD void step_init(const Ctx& c, BlockState& st, uint32_t lane)
{
    const uint32_t T = c.tab.texels;
    if (lane == 0) {
        fill_parts(st.slots[0], T, 1, 0, 0, 0);
        build_slot(st.cf, T, st.has_alpha != 0, st.slots[0]);
    } else if (lane == 1) {
        kmeans(st.cf, T, 2, &st.km[0]);
    } else if (lane == 2) {
        kmeans(st.cf, T, 3, &st.km[1]);
    } else if (lane >= 4 && lane < 8) {
        // dual-plane hypotheses: channel (lane - 4) on its own weight plane
        Slot& slot = st.slots[5 + (lane - 4)];
        fill_parts(slot, T, 1, 0, 0, 0);
        if (lane - 4 < (st.has_alpha ? 4u : 3u)) {
            slot.dual_ch = static_cast<int32_t>(lane - 4);
            build_slot(st.cf, T, st.has_alpha != 0, slot);
        }
    }
}

// step 2: every lane scans its share of the 1024 seeds for the one closest to the clustering
CFX_HD void step_rank(const Ctx& c, BlockState& st, uint32_t lane)
{
    const uint32_t T = c.tab.texels;
    const uint64_t full = T == 64 ? ~0ull : ((1ull << T) - 1ull);
    uint32_t b2 = 0xFFFFFFFFu, b3 = 0xFFFFFFFFu;
    for (uint32_t seed = lane; seed < 1024; seed += 32) {
        const uint64_t p = tab_u64(c, c.tab.off_part2 + seed*8u);
        if (p) b2 = min(b2, (mismatch2(st.km[0], p, T) << 10) | seed);
        const uint64_t p1 = tab_u64(c, c.tab.off_part3 + seed*16u), p2 = tab_u64(c, c.tab.off_part3 + seed*16u + 8u);
        if (p1) b3 = min(b3, (mismatch3(st.km[1], st.km[2], p1, p2, full) << 10) | seed);
    }
    st.keys[0][lane] = b2; st.keys[1][lane] = b3;
}

// step 3: exact line-fit residual of each lane's seeds
CFX_HD void step_score(const Ctx& c, BlockState& st, uint32_t lane)
{
    const uint32_t T = c.tab.texels;
    const bool ha = st.has_alpha != 0;
    st.scores[0][lane] = st.scores[1][lane] = 3.0e38f;
    if (st.keys[0][lane] != 0xFFFFFFFFu) {
        const uint32_t seed = st.keys[0][lane] & 1023u;
        st.scores[0][lane] = line_fit_residual(st.cf, T, 2, tab_u64(c, c.tab.off_part2 + seed*8u), 0, ha);
    }
    if (st.keys[1][lane] != 0xFFFFFFFFu) {
        const uint32_t seed = st.keys[1][lane] & 1023u;
        st.scores[1][lane] = line_fit_residual(st.cf, T, 3, tab_u64(c, c.tab.off_part3 + seed*16u),
            tab_u64(c, c.tab.off_part3 + seed*16u + 8u), ha);
    }
}

// step 4: lanes 0..3 turn the two best seeds of each subset count into slots 1..4
CFX_HD void step_slots(const Ctx& c, BlockState& st, uint32_t lane)
{
    if (lane >= 4) return;
    const uint32_t T = c.tab.texels;
    const uint32_t which = lane >> 1, rank = lane & 1u;
    // rank-th smallest (score, lane)
    uint32_t pick = 32;
    uint32_t skip = 32;
    for (uint32_t r = 0; r <= rank; ++r) {
        float best = 3.0e38f;
        pick = 32;
        for (uint32_t l = 0; l < 32; ++l)
            if (l != skip && st.scores[which][l] < best) { best = st.scores[which][l]; pick = l; }
        if (r < rank) skip = pick;
    }
    Slot& slot = st.slots[1 + lane];
    slot.valid = 0; slot.pc = which + 2;
    if (pick >= 32) return;
    const uint32_t seed = st.keys[which][pick] & 1023u;
    if (which == 0) fill_parts(slot, T, 2, seed, tab_u64(c, c.tab.off_part2 + seed*8u), 0);
    else fill_parts(slot, T, 3, seed, tab_u64(c, c.tab.off_part3 + seed*16u), tab_u64(c, c.tab.off_part3 + seed*16u + 8u));
    build_slot(st.cf, T, st.has_alpha != 0, slot);
}

} // namespace astc
} // namespace cfx
