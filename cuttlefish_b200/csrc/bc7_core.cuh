// Lane-local part of the BC7 encoder (see bc7.cu for the warp-level organisation): everything a
// single lane does for ITS (mode, shape) candidate.  Compiles for the device (bc7.cu) and, through
// hostdev.h, for the host, where tools/emu_bc7.cpp runs the same search block by block so that
// quality work can be iterated without a GPU.  The host build is a developer tool, never a
// fallback: libcfx.so contains only the device build.
#pragma once
#include "bc7_tables.cuh"
#include "hostdev.h"

namespace cfx {
namespace bc7 {

struct ModeInfo {
    uint32_t ns, cbits, abits, pmode, ibits;   // pmode: 0 none, 1 unique, 2 shared
};

CFX_HD ModeInfo mode_info(uint32_t mode)
{
    ModeInfo m;
    m.ns = mode == 6 ? 1u : 2u;
    m.cbits = mode == 1 ? 6u : (mode == 7 ? 5u : 7u);
    m.abits = mode == 6 ? 7u : (mode == 7 ? 5u : 0u);
    m.pmode = mode == 1 ? 2u : 1u;
    m.ibits = mode == 6 ? 4u : (mode == 1 ? 3u : 2u);
    return m;
}

// BC7 interpolation weight of index k for an ibits-bit index: {0,21,43,64}, {0,9,...,64},
// {0,4,9,...,64} == (k*64 + (N-1)/2) / (N-1); the division is done with a 17-bit reciprocal.
CFX_HD uint32_t index_weight(uint32_t k, uint32_t half, uint32_t recip)
{
    return ((k*64u + half)*recip) >> 17;
}

struct Fit {
    uint32_t e0[2], e1[2];     // quantised endpoints per subset, one byte per channel (incl. p-bit)
    uint32_t err[2];           // SSE per subset
    uint32_t sel_lo, sel_hi;   // 16 x 4-bit indices
    uint32_t sel2_lo, sel2_hi; // modes 4 / 5: the scalar channel's indices (sel_* are the colour indices)
};

CFX_HD uint32_t nibble_mask8(uint32_t m8)
{
    // spread 8 mask bits to 8 nibbles of 0xF
    uint32_t x = m8 & 0xFFu;
    x = (x | (x << 12)) & 0x000F000Fu;
    x = (x | (x << 6)) & 0x03030303u;
    x = (x | (x << 3)) & 0x11111111u;
    return x*15u;
}

// Principal axis of a symmetric 4x4 matrix (upper triangle c[10]: xx xy xz xw yy yz yw zz zw ww).
CFX_HD float4 principal_axis(const float* c, int iters)
{
    // start from the row with the largest diagonal so a degenerate start is impossible
    float4 v = make_float4(c[0], c[1], c[2], c[3]);
    float best = c[0];
    if (c[4] > best) { best = c[4]; v = make_float4(c[1], c[4], c[5], c[6]); }
    if (c[7] > best) { best = c[7]; v = make_float4(c[2], c[5], c[7], c[8]); }
    if (c[9] > best) { best = c[9]; v = make_float4(c[3], c[6], c[8], c[9]); }
    for (int it = 0; it < iters; ++it) {
        float n2 = v.x*v.x + v.y*v.y + v.z*v.z + v.w*v.w;
        float inv = n2 > 1e-20f ? rsqrtf(n2) : 0.0f;
        v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
        float4 r;
        r.x = c[0]*v.x + c[1]*v.y + c[2]*v.z + c[3]*v.w;
        r.y = c[1]*v.x + c[4]*v.y + c[5]*v.z + c[6]*v.w;
        r.z = c[2]*v.x + c[5]*v.y + c[7]*v.z + c[8]*v.w;
        r.w = c[3]*v.x + c[6]*v.y + c[8]*v.z + c[9]*v.w;
        v = r;
    }
    float n2 = v.x*v.x + v.y*v.y + v.z*v.z + v.w*v.w;
    float inv = n2 > 1e-20f ? rsqrtf(n2) : 0.0f;
    return make_float4(v.x*inv, v.y*inv, v.z*inv, v.w*inv);
}

// lambda_max estimate of the same matrix: |C v| for the unit v after `iters` power iterations.
CFX_HD float lambda_max(const float* c)
{
    float4 v = principal_axis(c, 3);
    float4 r;
    r.x = c[0]*v.x + c[1]*v.y + c[2]*v.z + c[3]*v.w;
    r.y = c[1]*v.x + c[4]*v.y + c[5]*v.z + c[6]*v.w;
    r.z = c[2]*v.x + c[5]*v.y + c[7]*v.z + c[8]*v.w;
    r.w = c[3]*v.x + c[6]*v.y + c[8]*v.z + c[9]*v.w;
    return v.x*r.x + v.y*r.y + v.z*r.z + v.w*r.w;
}

// What a subset's spread ALONG its line costs after index quantisation, as a fraction of that spread: texels spread
// evenly over the end point interval and rounded to one of N levels keep 1/(N-1)^2 of their variance as error (N = 8 for
// mode 1). Without this term every shape of a block whose texels are collinear (black-on-white text) scores zero and the
// ranking is arbitrary, although cutting the line into two short pieces is exactly what such blocks gain from.
#ifndef CFX_BC7_LINE_KEEP
#define CFX_BC7_LINE_KEEP (1.0f - 1.0f/49.0f)
#endif

// Phase-1 score of one two-subset shape: sum over subsets of (trace - lambda_max) of the subset's
// scatter matrix = squared distance of its texels from their best-fit line.  Returned as a sortable
// key with the shape number in the low 6 bits.  sT / cT: block totals of x and of the 10 products.
CFX_HD uint32_t score_shape(const float4* bxf, const float* sT, const float* cT, uint32_t shape)
{
    const uint32_t m1 = kBc7Part2[shape];
    float n1 = 0, s1[4] = {0, 0, 0, 0}, c1[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        if ((m1 >> i) & 1u) {
            float4 x = bxf[i];
            n1 += 1.0f;
            s1[0] += x.x; s1[1] += x.y; s1[2] += x.z; s1[3] += x.w;
            c1[0] += x.x*x.x; c1[1] += x.x*x.y; c1[2] += x.x*x.z; c1[3] += x.x*x.w;
            c1[4] += x.y*x.y; c1[5] += x.y*x.z; c1[6] += x.y*x.w;
            c1[7] += x.z*x.z; c1[8] += x.z*x.w; c1[9] += x.w*x.w;
        }
    }
    float score = 0.0f;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        float nn = s ? n1 : 16.0f - n1;
        float sm[4], cc[10];
#pragma unroll
        for (int k = 0; k < 4; ++k) sm[k] = s ? s1[k] : sT[k] - s1[k];
#pragma unroll
        for (int k = 0; k < 10; ++k) cc[k] = s ? c1[k] : cT[k] - c1[k];
        float inv = 1.0f/nn;     // every BC7 shape has both subsets non-empty
        cc[0] -= sm[0]*sm[0]*inv; cc[1] -= sm[0]*sm[1]*inv; cc[2] -= sm[0]*sm[2]*inv; cc[3] -= sm[0]*sm[3]*inv;
        cc[4] -= sm[1]*sm[1]*inv; cc[5] -= sm[1]*sm[2]*inv; cc[6] -= sm[1]*sm[3]*inv;
        cc[7] -= sm[2]*sm[2]*inv; cc[8] -= sm[2]*sm[3]*inv; cc[9] -= sm[3]*sm[3]*inv;
        score += (cc[0] + cc[4] + cc[7] + cc[9]) - CFX_BC7_LINE_KEEP*lambda_max(cc);
    }
    score = fmaxf(score, 0.0f);
    return (__float_as_uint(score) & ~63u) | shape;
}

// The same for a block whose alpha is constant: the alpha row and column of every scatter matrix are zero, so the
// products, the totals and the power iteration shrink from 4 to 3 dimensions (6 products instead of 10).
// c: xx xy xz yy yz zz.
CFX_HD float lambda_max3(const float* c)
{
    float v0 = c[0], v1 = c[1], v2 = c[2], best = c[0];
    if (c[3] > best) { best = c[3]; v0 = c[1]; v1 = c[3]; v2 = c[4]; }
    if (c[5] > best) { best = c[5]; v0 = c[2]; v1 = c[4]; v2 = c[5]; }
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const float n2 = v0*v0 + v1*v1 + v2*v2;
        const float inv = n2 > 1e-20f ? rsqrtf(n2) : 0.0f;
        v0 *= inv; v1 *= inv; v2 *= inv;
        const float r0 = c[0]*v0 + c[1]*v1 + c[2]*v2, r1 = c[1]*v0 + c[3]*v1 + c[4]*v2, r2 = c[2]*v0 + c[4]*v1 + c[5]*v2;
        v0 = r0; v1 = r1; v2 = r2;
    }
    const float n2 = v0*v0 + v1*v1 + v2*v2;
    const float inv = n2 > 1e-20f ? rsqrtf(n2) : 0.0f;
    v0 *= inv; v1 *= inv; v2 *= inv;
    const float r0 = c[0]*v0 + c[1]*v1 + c[2]*v2, r1 = c[1]*v0 + c[3]*v1 + c[4]*v2, r2 = c[2]*v0 + c[4]*v1 + c[5]*v2;
    return v0*r0 + v1*r1 + v2*r2;
}

// sT: block totals of x, y, z; cT: of the 6 products.
CFX_HD uint32_t score_shape_rgb(const float4* bxf, const float* sT, const float* cT, uint32_t shape)
{
    const uint32_t m1 = kBc7Part2[shape];
    float n1 = 0, s1[3] = {0, 0, 0}, c1[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        if ((m1 >> i) & 1u) {
            const float4 x = bxf[i];
            n1 += 1.0f;
            s1[0] += x.x; s1[1] += x.y; s1[2] += x.z;
            c1[0] += x.x*x.x; c1[1] += x.x*x.y; c1[2] += x.x*x.z;
            c1[3] += x.y*x.y; c1[4] += x.y*x.z; c1[5] += x.z*x.z;
        }
    }
    float score = 0.0f;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const float nn = s ? n1 : 16.0f - n1;
        float sm[3], cc[6];
#pragma unroll
        for (int k = 0; k < 3; ++k) sm[k] = s ? s1[k] : sT[k] - s1[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) cc[k] = s ? c1[k] : cT[k] - c1[k];
        const float inv = 1.0f/nn;
        cc[0] -= sm[0]*sm[0]*inv; cc[1] -= sm[0]*sm[1]*inv; cc[2] -= sm[0]*sm[2]*inv;
        cc[3] -= sm[1]*sm[1]*inv; cc[4] -= sm[1]*sm[2]*inv; cc[5] -= sm[2]*sm[2]*inv;
        score += (cc[0] + cc[3] + cc[5]) - CFX_BC7_LINE_KEEP*lambda_max3(cc);
    }
    score = fmaxf(score, 0.0f);
    return (__float_as_uint(score) & ~63u) | shape;
}

// Candidate descriptor of a lane: mode | rank << 4 | variant << 8 | LS rounds << 12, where rank is
// the position of the lane's partition shape in the phase-1 ranking (ignored for mode 6), variant 1
// = "extrapolating" first LS round, and rounds = number of least-squares refinement rounds.
#define CFX_BC7_CAND(mode, rank, variant, rounds) ((mode) | ((rank) << 4) | ((variant) << 8) | ((rounds) << 12))
CFX_HD uint32_t cand_mode(uint32_t d) { return d & 15u; }
CFX_HD bool cand_is_dual(uint32_t d) { return (d & 15u) == 4u || (d & 15u) == 5u; }      // modes 4, 5: rank field = rotation, variant bit 0 = index mode
CFX_HD uint32_t cand_rank(uint32_t d) { return (d & 15u) == 6u || cand_is_dual(d) ? 0xFFFFFFFFu : ((d >> 4) & 15u); }
CFX_HD uint32_t cand_rotation(uint32_t d) { return (d >> 4) & 3u; }
CFX_HD uint32_t cand_variant(uint32_t d) { return (d >> 8) & 15u; }
CFX_HD uint32_t cand_rounds(uint32_t d) { return (d >> 12) & 15u; }

// Quantise one float endpoint (0..255 per channel) to tb-bit codes with p-bit p (p = 2: mode has
// no p-bit), returning the codes packed one per byte and the squared quantisation error.
CFX_HD uint32_t quantize_endpoint(float4 x, uint32_t cbits, uint32_t abits,
    bool has_p, uint32_t p, float& qerr)
{
    float v[4] = {x.x, x.y, x.z, x.w};
    uint32_t out = 0;
    qerr = 0.0f;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        uint32_t bits = ch == 3 ? abits : cbits;
        if (bits == 0) continue;                       // mode has no alpha: decodes to 255
        uint32_t tb = bits + (has_p ? 1u : 0u);
        float maxv = static_cast<float>((1u << tb) - 1u);
        float f = fminf(fmaxf(v[ch], 0.0f), 255.0f)*(maxv*(1.0f/255.0f));
        uint32_t code;
        if (has_p) {
            int q = __float2int_rn((f - static_cast<float>(p))*0.5f);
            q = min(max(q, 0), static_cast<int>((1u << bits) - 1u));
            code = (static_cast<uint32_t>(q) << 1) | p;
        } else {
            int q = __float2int_rn(f);
            code = static_cast<uint32_t>(min(max(q, 0), static_cast<int>((1u << bits) - 1u)));
        }
        uint32_t dq = (code << (8u - tb)) | (code >> (2u*tb - 8u));
        float d = static_cast<float>(dq) - v[ch];
        qerr += d*d;
        out |= code << (8*ch);
    }
    return out;
}

// Expand packed tb-bit codes to the 8-bit values the decoder interpolates.
CFX_HD uint32_t dequant_endpoint(uint32_t codes, uint32_t ctb, uint32_t atb)
{
    // colour bytes
    uint32_t c = codes & 0x00FFFFFFu;
    uint32_t rmask = ((1u << (8u - ctb)) - 1u)*0x010101u;
    uint32_t rgb = ((c << (8u - ctb)) | ((c >> (2u*ctb - 8u)) & rmask)) & 0x00FFFFFFu;
    uint32_t a = 0xFFu;
    if (atb) {
        uint32_t av = codes >> 24;
        a = ((av << (8u - atb)) | (av >> (2u*atb - 8u))) & 0xFFu;
    }
    return rgb | (a << 24);
}

// Mixed-sign 2-way dot products: a = two signed 16-bit halves, b = unsigned bytes
// (lo: bytes 0,1; hi: bytes 2,3).  The CUDA intrinsics only offer same-sign variants.
CFX_HD int dp2a_lo_s16_u8(uint32_t a, uint32_t b, int c)
{
#ifdef __CUDA_ARCH__
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return c + (int)(int16_t)(a & 0xFFFF)*(int)(b & 0xFF) + (int)(int16_t)(a >> 16)*(int)((b >> 8) & 0xFF);
#endif
}
CFX_HD int dp2a_hi_s16_u8(uint32_t a, uint32_t b, int c)
{
#ifdef __CUDA_ARCH__
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return c + (int)(int16_t)(a & 0xFFFF)*(int)((b >> 16) & 0xFF) + (int)(int16_t)(a >> 16)*(int)(b >> 24);
#endif
}

struct SubsetEval {
    uint32_t lo_rg, lo_ba, hi_rg, hi_ba;   // dequantised endpoints as 2x16-bit pairs
    uint32_t d_rg, d_ba;                   // hi - lo as signed 16-bit pairs
    int c0;                                // dot(lo, d)
    float scale;                           // (N-1) / |d|^2
};

CFX_HD SubsetEval make_eval(uint32_t e0, uint32_t e1, uint32_t ctb, uint32_t atb,
    uint32_t nm1)
{
    uint32_t lo = dequant_endpoint(e0, ctb, atb), hi = dequant_endpoint(e1, ctb, atb);
    SubsetEval s;
    s.lo_rg = (lo & 0xFFu) | ((lo & 0xFF00u) << 8);
    s.lo_ba = ((lo >> 16) & 0xFFu) | ((lo >> 8) & 0xFF0000u);
    s.hi_rg = (hi & 0xFFu) | ((hi & 0xFF00u) << 8);
    s.hi_ba = ((hi >> 16) & 0xFFu) | ((hi >> 8) & 0xFF0000u);
    int dr = static_cast<int>(hi & 0xFF) - static_cast<int>(lo & 0xFF);
    int dg = static_cast<int>((hi >> 8) & 0xFF) - static_cast<int>((lo >> 8) & 0xFF);
    int db = static_cast<int>((hi >> 16) & 0xFF) - static_cast<int>((lo >> 16) & 0xFF);
    int da = static_cast<int>(hi >> 24) - static_cast<int>(lo >> 24);
    s.d_rg = (static_cast<uint32_t>(dr) & 0xFFFFu) | (static_cast<uint32_t>(dg) << 16);
    s.d_ba = (static_cast<uint32_t>(db) & 0xFFFFu) | (static_cast<uint32_t>(da) << 16);
    s.c0 = dr*static_cast<int>(lo & 0xFF) + dg*static_cast<int>((lo >> 8) & 0xFF) +
        db*static_cast<int>((lo >> 16) & 0xFF) + da*static_cast<int>(lo >> 24);
    int len2 = dr*dr + dg*dg + db*db + da*da;
    s.scale = len2 > 0 ? static_cast<float>(nm1)/static_cast<float>(len2) : 0.0f;
    return s;
}

// Exact SSE of texel x against palette entry k of the subset (weights 1 on enabled channels).
CFX_HD uint32_t entry_error(const SubsetEval& s, uint32_t x, uint32_t w, uint32_t chmask)
{
    uint32_t iw = 64u - w;
    uint32_t rg = ((s.lo_rg*iw + s.hi_rg*w + 0x00200020u) >> 6) & 0x00FF00FFu;
    uint32_t ba = ((s.lo_ba*iw + s.hi_ba*w + 0x00200020u) >> 6) & 0x00FF00FFu;
    uint32_t pal = __byte_perm(rg, ba, 0x6420);          // R G B A
    uint32_t d = __vabsdiffu4(pal, x) & chmask;
    return __dp4a(d, d, 0u);
}

// Assign indices and accumulate per-subset SSE for the lane's candidate.
CFX_HD void evaluate(const uint32_t* s_x, uint32_t m1, const SubsetEval& s0,
    const SubsetEval& s1, uint32_t nm1, uint32_t half, uint32_t recip, uint32_t chmask, uint32_t err[2],
    uint32_t& sel_lo, uint32_t& sel_hi)
{
    uint32_t e0 = 0, e1 = 0;
    uint64_t sel = 0;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        uint32_t x = s_x[i];
        bool in1 = (m1 >> i) & 1u;
        SubsetEval s;
        s.lo_rg = in1 ? s1.lo_rg : s0.lo_rg; s.lo_ba = in1 ? s1.lo_ba : s0.lo_ba;
        s.hi_rg = in1 ? s1.hi_rg : s0.hi_rg; s.hi_ba = in1 ? s1.hi_ba : s0.hi_ba;
        s.d_rg = in1 ? s1.d_rg : s0.d_rg; s.d_ba = in1 ? s1.d_ba : s0.d_ba;
        s.c0 = in1 ? s1.c0 : s0.c0; s.scale = in1 ? s1.scale : s0.scale;
        int dot = dp2a_lo_s16_u8(s.d_rg, x, 0);        // dR*R + dG*G
        dot = dp2a_hi_s16_u8(s.d_ba, x, dot);          // + dB*B + dA*A
        // nearest index along the endpoint axis, then the neighbour on the side the texel lies on
        const float t = static_cast<float>(dot - s.c0)*s.scale;
        int k = __float2int_rn(t);
        k = min(max(k, 0), static_cast<int>(nm1));
        const int kn = min(max(t > static_cast<float>(k) ? k + 1 : k - 1, 0), static_cast<int>(nm1));
        uint32_t er = entry_error(s, x, index_weight(k, half, recip), chmask);
        const uint32_t ern = entry_error(s, x, index_weight(kn, half, recip), chmask);
        uint32_t bk = k;
        if (ern < er) { er = ern; bk = kn; }
        if (in1) e1 += er; else e0 += er;
        sel |= static_cast<uint64_t>(bk) << (4*i);
    }
    err[0] = e0; err[1] = e1;
    sel_lo = static_cast<uint32_t>(sel); sel_hi = static_cast<uint32_t>(sel >> 32);
}

// Choose p-bits + quantise both endpoints of one subset.
CFX_HD void quantize_pair(float4 lo, float4 hi, const ModeInfo& mi, uint32_t& e0,
    uint32_t& e1)
{
    float q00, q01, q10, q11;
    uint32_t a0 = quantize_endpoint(lo, mi.cbits, mi.abits, true, 0, q00);
    uint32_t a1 = quantize_endpoint(lo, mi.cbits, mi.abits, true, 1, q01);
    uint32_t b0 = quantize_endpoint(hi, mi.cbits, mi.abits, true, 0, q10);
    uint32_t b1 = quantize_endpoint(hi, mi.cbits, mi.abits, true, 1, q11);
    if (mi.pmode == 2) {          // shared p-bit
        bool one = (q01 + q11) < (q00 + q10);
        e0 = one ? a1 : a0; e1 = one ? b1 : b0;
    } else {
        e0 = q01 < q00 ? a1 : a0;
        e1 = q11 < q10 ? b1 : b0;
    }
}

// Quantise the least-squares endpoints of one subset with the indices held fixed.  With fixed
// indices the SSE of channel c is the quadratic  A l^2 + 2B l h + C h^2 - 2 P_c l - 2 Q_c h  in the
// decoded endpoint values (l, h), so for every p-bit choice each channel independently tries the
// two representable values below/above the real-valued optimum for l and for h (4 pairs) and the
// p-bit combination with the smallest total wins.  This is what makes p-bit and rounding
// decisions exact instead of nearest-value guesses.
CFX_HD void quantize_ls(float A, float B, float C, const float* P, const float* Q,
    const float* lo, const float* hi, const ModeInfo& mi, uint32_t& e0, uint32_t& e1)
{
    // Every mode fitted here has cbits == abits or no alpha at all, so one lattice serves all channels.
    const uint32_t tb = mi.cbits + 1u;
    const float scale = static_cast<float>((1u << tb) - 1u)*(1.0f/255.0f);
    const int qmax = static_cast<int>((1u << mi.cbits) - 1u);
    const int nch = mi.abits ? 4 : 3;
    // E[pl + 2 ph] accumulates, over the channels, the best of the 2x2 codes around the optimum
    float E[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    uint32_t c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        if (ch >= nch) continue;
        const float fl = fminf(fmaxf(lo[ch], 0.0f), 255.0f)*scale, fh = fminf(fmaxf(hi[ch], 0.0f), 255.0f)*scale;
        // the four codes around each optimum: index = parity*2 + (floor, floor+1)
        uint32_t lc[4], hc[4];
        float lb[4], le[4], hv[4], he[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int par = k >> 1;
            const int ql = static_cast<int>(floorf((fl - static_cast<float>(par))*0.5f)) + (k & 1);
            const int qh = static_cast<int>(floorf((fh - static_cast<float>(par))*0.5f)) + (k & 1);
            lc[k] = (static_cast<uint32_t>(min(max(ql, 0), qmax)) << 1) | static_cast<uint32_t>(par);
            hc[k] = (static_cast<uint32_t>(min(max(qh, 0), qmax)) << 1) | static_cast<uint32_t>(par);
            const float l = static_cast<float>((lc[k] << (8u - tb)) | (lc[k] >> (2u*tb - 8u)));
            const float h = static_cast<float>((hc[k] << (8u - tb)) | (hc[k] >> (2u*tb - 8u)));
            le[k] = l*(A*l - 2.0f*P[ch]); lb[k] = 2.0f*B*l;
            he[k] = h*(C*h - 2.0f*Q[ch]); hv[k] = h;
        }
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
            if (mi.pmode == 2 && (pc == 1 || pc == 2)) continue;      // shared p-bit: pl == ph
            const int pl = pc & 1, ph = pc >> 1;
            float bestc = 3.4e38f;
            uint32_t bl = 0, bh = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = pl*2 + (k & 1), j = ph*2 + (k >> 1);
                const float e = lb[i]*hv[j] + (le[i] + he[j]);
                if (e < bestc) { bestc = e; bl = lc[i]; bh = hc[j]; }
            }
            E[pc] += bestc;
            c0[pc] |= bl << (8*ch); c1[pc] |= bh << (8*ch);
        }
    }
    uint32_t b0 = c0[0], b1 = c1[0];
    float bestE = E[0];
    if (mi.pmode != 2) {
        if (E[1] < bestE) { bestE = E[1]; b0 = c0[1]; b1 = c1[1]; }
        if (E[2] < bestE) { bestE = E[2]; b0 = c0[2]; b1 = c1[2]; }
    }
    if (E[3] < bestE) { bestE = E[3]; b0 = c0[3]; b1 = c1[3]; }
    e0 = b0; e1 = b1;
}

// Endpoints that reproduce ONE colour as closely as the mode's endpoint lattice allows: every texel
// of the subset uses index k (weight w), and per channel the pair (L, H) is searched around the
// solutions of ((64-w) L + w H + 32) >> 6 == c.  This is what makes flat subsets (a constant colour,
// or a spread too small to separate) exact instead of "nearest representable endpoint".
CFX_HD uint32_t flat_fit(uint32_t target, const ModeInfo& mi, uint32_t w, uint32_t chmask, uint32_t& e0,
    uint32_t& e1)
{
    uint32_t bestE = 0xFFFFFFFFu;
    const int combos = mi.pmode == 2 ? 2 : 4;
    const float invw = 1.0f/static_cast<float>(w);
#pragma unroll 1
    for (int pc = 0; pc < combos; ++pc) {
        const uint32_t pl = mi.pmode == 2 ? pc : (pc & 1), ph = mi.pmode == 2 ? pc : (pc >> 1);
        uint32_t E = 0, c0 = 0, c1 = 0;
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            const uint32_t bits = ch == 3 ? mi.abits : mi.cbits;
            if (bits == 0) continue;
            const uint32_t tb = bits + 1u;
            const int qmax = static_cast<int>((1u << bits) - 1u);
            const float scale = static_cast<float>((1u << tb) - 1u)*(1.0f/255.0f);
            const int c = static_cast<int>((target >> (8*ch)) & 0xFFu);
            const int q0 = __float2int_rn((static_cast<float>(c)*scale - static_cast<float>(pl))*0.5f);
            uint32_t bestc = 0xFFFFFFFFu, bl = 0, bh = 0;
#pragma unroll 1
            for (int dl = -2; dl <= 2; ++dl) {
                const uint32_t ca = (static_cast<uint32_t>(min(max(q0 + dl, 0), qmax)) << 1) | pl;
                const int L = static_cast<int>((ca << (8u - tb)) | (ca >> (2u*tb - 8u)));
                const float Hf = static_cast<float>(64*c - (64 - static_cast<int>(w))*L)*invw;
                const int qh0 = static_cast<int>(floorf((fminf(fmaxf(Hf, 0.0f), 255.0f)*scale - static_cast<float>(ph))*0.5f));
#pragma unroll
                for (int dh = 0; dh < 2; ++dh) {
                    const uint32_t cb = (static_cast<uint32_t>(min(max(qh0 + dh, 0), qmax)) << 1) | ph;
                    const int H = static_cast<int>((cb << (8u - tb)) | (cb >> (2u*tb - 8u)));
                    const int v = ((64 - static_cast<int>(w))*L + static_cast<int>(w)*H + 32) >> 6;
                    const uint32_t e = static_cast<uint32_t>((v - c)*(v - c));
                    if (e < bestc) { bestc = e; bl = ca; bh = cb; }
                }
            }
            if ((chmask >> (8*ch)) & 1u) E += bestc;
            c0 |= bl << (8*ch); c1 |= bh << (8*ch);
        }
        if (E < bestE) { bestE = E; e0 = c0; e1 = c1; }
    }
    return bestE;
}

// The whole fit of one (mode, shape) candidate by one lane.
CFX_HD void fit_candidate(const float4* s_xf, const uint32_t* s_x, uint32_t mode,
    uint32_t m1, uint32_t variant, uint32_t rounds, uint32_t chmask, Fit& best)
{
    const ModeInfo mi = mode_info(mode);
    const uint32_t ctb = mi.cbits + 1u, atb = mi.abits ? mi.abits + 1u : 0u;
    const uint32_t nm1 = (1u << mi.ibits) - 1u;
    const uint32_t half = nm1 >> 1;
    const uint32_t recip = nm1 == 3 ? 43691u : (nm1 == 7 ? 18725u : 8739u);

    // ---- statistics: totals and subset 1; subset 0 = total - subset 1
    float nT = 16.0f, n1 = 0.0f;
    float sT[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
    float cT[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, c1[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 2
    for (int i = 0; i < 16; ++i) {
        float4 x = s_xf[i];
        float f = ((m1 >> i) & 1u) ? 1.0f : 0.0f;
        float o[10] = {x.x*x.x, x.x*x.y, x.x*x.z, x.x*x.w, x.y*x.y, x.y*x.z, x.y*x.w, x.z*x.z, x.z*x.w, x.w*x.w};
        sT[0] += x.x; sT[1] += x.y; sT[2] += x.z; sT[3] += x.w;
        s1[0] += f*x.x; s1[1] += f*x.y; s1[2] += f*x.z; s1[3] += f*x.w;
        n1 += f;
#pragma unroll
        for (int k = 0; k < 10; ++k) { cT[k] += o[k]; c1[k] += f*o[k]; }
    }
    float4 mean[2], axis[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        float n = s ? n1 : nT - n1;
        float sm[4], cc[10];
#pragma unroll
        for (int k = 0; k < 4; ++k) sm[k] = s ? s1[k] : sT[k] - s1[k];
#pragma unroll
        for (int k = 0; k < 10; ++k) cc[k] = s ? c1[k] : cT[k] - c1[k];
        float inv = n > 0.0f ? 1.0f/n : 0.0f;
        float m[4] = {sm[0]*inv, sm[1]*inv, sm[2]*inv, sm[3]*inv};
        cc[0] -= sm[0]*m[0]; cc[1] -= sm[0]*m[1]; cc[2] -= sm[0]*m[2]; cc[3] -= sm[0]*m[3];
        cc[4] -= sm[1]*m[1]; cc[5] -= sm[1]*m[2]; cc[6] -= sm[1]*m[3];
        cc[7] -= sm[2]*m[2]; cc[8] -= sm[2]*m[3]; cc[9] -= sm[3]*m[3];
        mean[s] = make_float4(m[0], m[1], m[2], m[3]);
        axis[s] = principal_axis(cc, 4);
    }

    // ---- extent of each subset along its axis
    float tmin[2] = {1e30f, 1e30f}, tmax[2] = {-1e30f, -1e30f};
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        float4 x = s_xf[i];
        bool in1 = (m1 >> i) & 1u;
        float4 mu = in1 ? mean[1] : mean[0];
        float4 ax = in1 ? axis[1] : axis[0];
        float t = (x.x - mu.x)*ax.x + (x.y - mu.y)*ax.y + (x.z - mu.z)*ax.z + (x.w - mu.w)*ax.w;
        if (in1) { tmin[1] = fminf(tmin[1], t); tmax[1] = fmaxf(tmax[1], t); }
        else { tmin[0] = fminf(tmin[0], t); tmax[0] = fmaxf(tmax[0], t); }
    }

    uint32_t e0[2], e1[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        float lo_t = tmin[s] < 1e29f ? tmin[s] : 0.0f, hi_t = tmax[s] > -1e29f ? tmax[s] : 0.0f;
        float4 lo = make_float4(mean[s].x + lo_t*axis[s].x, mean[s].y + lo_t*axis[s].y,
            mean[s].z + lo_t*axis[s].z, mean[s].w + lo_t*axis[s].w);
        float4 hi = make_float4(mean[s].x + hi_t*axis[s].x, mean[s].y + hi_t*axis[s].y,
            mean[s].z + hi_t*axis[s].z, mean[s].w + hi_t*axis[s].w);
        quantize_pair(lo, hi, mi, e0[s], e1[s]);
    }

    SubsetEval ev0 = make_eval(e0[0], e1[0], ctb, atb, nm1);
    SubsetEval ev1 = make_eval(e0[1], e1[1], ctb, atb, nm1);
    evaluate(s_x, m1, ev0, ev1, nm1, half, recip, chmask, best.err, best.sel_lo, best.sel_hi);
    best.e0[0] = e0[0]; best.e0[1] = e0[1]; best.e1[0] = e1[0]; best.e1[1] = e1[1];

    // ---- least-squares refinement: solve for endpoints given the indices, keep per-subset wins
    uint32_t cur_lo = best.sel_lo, cur_hi = best.sel_hi;
    for (uint32_t round = 0; round < rounds; ++round) {
        // normal equations per subset: [A B; B C] [lo hi]^T = [P Q]^T
        float AT = 0, BT = 0, CT = 0, A1 = 0, B1 = 0, C1 = 0;
        float PT[4] = {0, 0, 0, 0}, QT[4] = {0, 0, 0, 0}, P1[4] = {0, 0, 0, 0}, Q1[4] = {0, 0, 0, 0};
#pragma unroll 2
        for (int i = 0; i < 16; ++i) {
            float4 x = s_xf[i];
            uint32_t k = ((i < 8 ? cur_lo : cur_hi) >> (4*(i & 7))) & 15u;
            // variant 1, first round: pull the extreme indices inwards so the solved endpoints
            // extrapolate beyond the texel range (finds the wide-endpoint encodings that make
            // near-flat blocks exact)
            if ((variant & 1u) && round == 0) k = min(max(k, 1u), nm1 - 1u);
            float w = static_cast<float>(index_weight(k, half, recip))*(1.0f/64.0f);
            float iw = 1.0f - w;
            float f = ((m1 >> i) & 1u) ? 1.0f : 0.0f;
            float a = iw*iw, b = iw*w, c = w*w;
            AT += a; BT += b; CT += c; A1 += f*a; B1 += f*b; C1 += f*c;
            float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float p = iw*xs[ch], q = w*xs[ch];
                PT[ch] += p; QT[ch] += q; P1[ch] += f*p; Q1[ch] += f*q;
            }
        }
        uint32_t n0[2], n1e[2];
        bool ok[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            float A = s ? A1 : AT - A1, B = s ? B1 : BT - B1, C = s ? C1 : CT - C1;
            float det = A*C - B*B;
            ok[s] = fabsf(det) > 1e-4f;
            float id = ok[s] ? 1.0f/det : 0.0f;
            float lo[4], hi[4], Ps[4], Qs[4];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                Ps[ch] = s ? P1[ch] : PT[ch] - P1[ch]; Qs[ch] = s ? Q1[ch] : QT[ch] - Q1[ch];
                lo[ch] = (C*Ps[ch] - B*Qs[ch])*id;
                hi[ch] = (A*Qs[ch] - B*Ps[ch])*id;
            }
            // the intermediate rounds only have to produce good indices for the next solve: nearest
            // endpoint codes are enough there; the last round decides p-bits and rounding exactly
            if (round + 1 < rounds)
                quantize_pair(make_float4(lo[0], lo[1], lo[2], lo[3]), make_float4(hi[0], hi[1], hi[2], hi[3]), mi, n0[s], n1e[s]);
            else
                quantize_ls(A, B, C, Ps, Qs, lo, hi, mi, n0[s], n1e[s]);
            if (!ok[s]) { n0[s] = best.e0[s]; n1e[s] = best.e1[s]; }
        }
        ev0 = make_eval(n0[0], n1e[0], ctb, atb, nm1);
        ev1 = make_eval(n0[1], n1e[1], ctb, atb, nm1);
        uint32_t nerr[2], nlo, nhi;
        evaluate(s_x, m1, ev0, ev1, nm1, half, recip, chmask, nerr, nlo, nhi);
        cur_lo = nlo; cur_hi = nhi;
        uint32_t take = 0;                      // texel mask of subsets that improved
        if (nerr[0] < best.err[0]) { best.err[0] = nerr[0]; best.e0[0] = n0[0]; best.e1[0] = n1e[0]; take |= ~m1 & 0xFFFFu; }
        if (nerr[1] < best.err[1]) { best.err[1] = nerr[1]; best.e0[1] = n0[1]; best.e1[1] = n1e[1]; take |= m1; }
        uint32_t tl = nibble_mask8(take), th = nibble_mask8(take >> 8);
        best.sel_lo = (best.sel_lo & ~tl) | (nlo & tl);
        best.sel_hi = (best.sel_hi & ~th) | (nhi & th);
    }

    // ---- flat subsets: every texel of the subset ended up on one index (constant colour, or a
    // spread the endpoints cannot resolve) and is still not exact -> fit that single colour exactly
    {
        uint32_t n0[2] = {best.e0[0], best.e0[1]}, n1e[2] = {best.e1[0], best.e1[1]};
        bool any = false;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const uint32_t mask = s ? m1 : (~m1 & 0xFFFFu);
            if (mask == 0 || best.err[s] == 0) continue;
            const uint32_t nl = nibble_mask8(mask), nh = nibble_mask8(mask >> 8);
            // all indices of the subset equal?  compare against the first texel's index replicated
            const uint32_t first = static_cast<uint32_t>(__ffs(mask) - 1);
            const uint32_t k0 = ((first < 8 ? best.sel_lo : best.sel_hi) >> (4*(first & 7))) & 15u;
            const uint32_t rep = k0*0x11111111u;
            if (((best.sel_lo ^ rep) & nl) | ((best.sel_hi ^ rep) & nh)) continue;
            const float4 mu = mean[s];
            const uint32_t target = static_cast<uint32_t>(__float2int_rn(mu.x)) |
                (static_cast<uint32_t>(__float2int_rn(mu.y)) << 8) |
                (static_cast<uint32_t>(__float2int_rn(mu.z)) << 16) |
                (static_cast<uint32_t>(__float2int_rn(mu.w)) << 24);
            const uint32_t kA = nm1 == 3 ? 1u : (nm1 == 7 ? 2u : 5u), kB = nm1 == 3 ? 2u : (nm1 == 7 ? 3u : 7u);
            uint32_t a0, a1, b0, b1;
            const uint32_t EA = flat_fit(target, mi, index_weight(kA, half, recip), chmask, a0, a1);
            const uint32_t EB = EA ? flat_fit(target, mi, index_weight(kB, half, recip), chmask, b0, b1) : 1u;
            n0[s] = EB < EA ? b0 : a0; n1e[s] = EB < EA ? b1 : a1;
            any = true;
        }
        if (any) {
            ev0 = make_eval(n0[0], n1e[0], ctb, atb, nm1);
            ev1 = make_eval(n0[1], n1e[1], ctb, atb, nm1);
            uint32_t nerr[2], nlo, nhi;
            evaluate(s_x, m1, ev0, ev1, nm1, half, recip, chmask, nerr, nlo, nhi);
            uint32_t take = 0;
            if (nerr[0] < best.err[0]) { best.err[0] = nerr[0]; best.e0[0] = n0[0]; best.e1[0] = n1e[0]; take |= ~m1 & 0xFFFFu; }
            if (nerr[1] < best.err[1]) { best.err[1] = nerr[1]; best.e0[1] = n0[1]; best.e1[1] = n1e[1]; take |= m1; }
            uint32_t tl = nibble_mask8(take), th = nibble_mask8(take >> 8);
            best.sel_lo = (best.sel_lo & ~tl) | (nlo & tl);
            best.sel_hi = (best.sel_hi & ~th) | (nhi & th);
        }
    }
}

struct BitWriter {
    uint64_t lo, hi;
    uint32_t pos;
    CFX_HD void init() { lo = hi = 0; pos = 0; }
    CFX_HD void put(uint32_t v, uint32_t bits)
    {
        uint64_t vv = v;
        if (pos < 64) {
            lo |= vv << pos;
            if (pos + bits > 64) hi |= vv >> (64u - pos);
        } else {
            hi |= vv << (pos - 64u);
        }
        pos += bits;
    }
};

// ---- modes 4 and 5: one subset, colour and one scalar channel on SEPARATE index sets, channel rotation ----------------
// (bc7enc uses mode 5 for blocks with alpha, lib/bc7enc_rdo/bc7enc.cpp:2039-2137; bc7e all of them,
// lib/bc7enc_rdo/bc7e.ispc:3909-4603.)  Rotation r swaps alpha with channel r-1 after decoding, so the encoder swaps
// them before: the "scalar" is then R, G or B and the colour triple carries alpha in its place -- what a block needs
// when one channel runs independently of the others.  Mode 5: colour 7.7.7 + scalar 8 bits, two 2-bit index sets.
// Mode 4: colour 5.5.5 + scalar 6 bits, a 2-bit and a 3-bit set; idx_mode says which one the colour gets.
// No p-bits.  Fit::e0[0] / e1[0] hold the colour codes in bytes 0..2 and the scalar code in byte 3; err[0] is the colour
// SSE, err[1] the scalar's; sel_* the colour indices, sel2_* the scalar's.
CFX_HD uint32_t rotation_selector(uint32_t rot) { return rot == 0 ? 0x3210u : (rot == 1 ? 0x0213u : (rot == 2 ? 0x1230u : 0x2310u)); }

CFX_HD uint32_t eval_colour3(const uint32_t* s_x, uint32_t psel, const SubsetEval& s, uint32_t nm1, uint32_t half, uint32_t recip,
    uint32_t cmask, uint32_t& sel_lo, uint32_t& sel_hi)
{
    uint32_t err = 0;
    uint64_t sel = 0;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const uint32_t x = __byte_perm(s_x[i], 0u, psel);
        int dot = dp2a_lo_s16_u8(s.d_rg, x, 0);
        dot = dp2a_hi_s16_u8(s.d_ba, x, dot);              // the alpha delta is zero: both end points decode to 255
        const float t = static_cast<float>(dot - s.c0)*s.scale;
        int k = min(max(__float2int_rn(t), 0), static_cast<int>(nm1));
        const int kn = min(max(t > static_cast<float>(k) ? k + 1 : k - 1, 0), static_cast<int>(nm1));
        uint32_t er = entry_error(s, x, index_weight(k, half, recip), cmask);
        const uint32_t ern = entry_error(s, x, index_weight(kn, half, recip), cmask);
        if (ern < er) { er = ern; k = kn; }
        err += er;
        sel |= static_cast<uint64_t>(k) << (4*i);
    }
    sel_lo = static_cast<uint32_t>(sel); sel_hi = static_cast<uint32_t>(sel >> 32);
    return err;
}

CFX_HD uint32_t eval_scalar(const uint32_t* s_x, uint32_t psel, int lo, int hi, uint32_t nm1, uint32_t half, uint32_t recip,
    bool enabled, uint32_t& sel_lo, uint32_t& sel_hi)
{
    uint32_t err = 0;
    uint64_t sel = 0;
    const float scale = hi != lo ? static_cast<float>(nm1)/static_cast<float>(hi - lo) : 0.0f;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const int v = static_cast<int>(__byte_perm(s_x[i], 0u, psel) >> 24);
        const float t = static_cast<float>(v - lo)*scale;
        int k = min(max(__float2int_rn(t), 0), static_cast<int>(nm1));
        const int kn = min(max(t > static_cast<float>(k) ? k + 1 : k - 1, 0), static_cast<int>(nm1));
        const int w = static_cast<int>(index_weight(k, half, recip)), wn = static_cast<int>(index_weight(kn, half, recip));
        const int d = ((lo*(64 - w) + hi*w + 32) >> 6) - v, dn = ((lo*(64 - wn) + hi*wn + 32) >> 6) - v;
        uint32_t er = static_cast<uint32_t>(d*d);
        if (static_cast<uint32_t>(dn*dn) < er) { er = static_cast<uint32_t>(dn*dn); k = kn; }
        err += er;
        sel |= static_cast<uint64_t>(k) << (4*i);
    }
    sel_lo = static_cast<uint32_t>(sel); sel_hi = static_cast<uint32_t>(sel >> 32);
    return enabled ? err : 0u;
}

CFX_HD uint32_t scalar_code(float v, uint32_t bits)
{
    const float maxv = static_cast<float>((1u << bits) - 1u);
    return static_cast<uint32_t>(min(max(__float2int_rn(fminf(fmaxf(v, 0.0f), 255.0f)*(maxv*(1.0f/255.0f))), 0), static_cast<int>(maxv)));
}
CFX_HD int scalar_value(uint32_t code, uint32_t bits) { return static_cast<int>(((code << (8u - bits)) | (code >> (2u*bits - 8u))) & 0xFFu); }

CFX_HD_NOINLINE void fit_dual(const uint32_t* s_x, uint32_t mode, uint32_t rot, uint32_t idx_mode, uint32_t rounds, uint32_t chmask, Fit& best)
{
    const uint32_t psel = rotation_selector(rot);
    const uint32_t cmask = __byte_perm(chmask, 0u, psel);            // the channel enables travel with their channels
    const uint32_t cbits = mode == 5 ? 7u : 5u, abits = mode == 5 ? 8u : 6u;
    const uint32_t cib = mode == 5 ? 2u : (idx_mode ? 3u : 2u), aib = mode == 5 ? 2u : (idx_mode ? 2u : 3u);
    const uint32_t cn = (1u << cib) - 1u, an = (1u << aib) - 1u;
    const uint32_t chalf = cn >> 1, ahalf = an >> 1;
    const uint32_t crecip = cn == 3 ? 43691u : 18725u, arecip = an == 3 ? 43691u : 18725u;

    // ---- colour triple: principal axis, quantise, indices, least squares
    float sm[3] = {0, 0, 0}, cc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    int amin = 255, amax = 0;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const uint32_t x = __byte_perm(s_x[i], 0u, psel);
        const float r = static_cast<float>(x & 0xFFu), g = static_cast<float>((x >> 8) & 0xFFu), b = static_cast<float>((x >> 16) & 0xFFu);
        sm[0] += r; sm[1] += g; sm[2] += b;
        cc[0] += r*r; cc[1] += r*g; cc[2] += r*b; cc[4] += g*g; cc[5] += g*b; cc[7] += b*b;
        amin = min(amin, static_cast<int>(x >> 24)); amax = max(amax, static_cast<int>(x >> 24));
    }
    const float m[3] = {sm[0]*(1.0f/16.0f), sm[1]*(1.0f/16.0f), sm[2]*(1.0f/16.0f)};
    cc[0] -= sm[0]*m[0]; cc[1] -= sm[0]*m[1]; cc[2] -= sm[0]*m[2]; cc[4] -= sm[1]*m[1]; cc[5] -= sm[1]*m[2]; cc[7] -= sm[2]*m[2];
    const float4 ax = principal_axis(cc, 4);
    float tmin = 1e30f, tmax = -1e30f;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const uint32_t x = __byte_perm(s_x[i], 0u, psel);
        const float t = (static_cast<float>(x & 0xFFu) - m[0])*ax.x + (static_cast<float>((x >> 8) & 0xFFu) - m[1])*ax.y +
            (static_cast<float>((x >> 16) & 0xFFu) - m[2])*ax.z;
        tmin = fminf(tmin, t); tmax = fmaxf(tmax, t);
    }
    float q;
    uint32_t e0 = quantize_endpoint(make_float4(m[0] + tmin*ax.x, m[1] + tmin*ax.y, m[2] + tmin*ax.z, 255.0f), cbits, 0u, false, 0u, q);
    uint32_t e1 = quantize_endpoint(make_float4(m[0] + tmax*ax.x, m[1] + tmax*ax.y, m[2] + tmax*ax.z, 255.0f), cbits, 0u, false, 0u, q);
    SubsetEval ev = make_eval(e0, e1, cbits, 0u, cn);
    uint32_t clo, chi;
    uint32_t cerr = eval_colour3(s_x, psel, ev, cn, chalf, crecip, cmask & 0x00FFFFFFu, clo, chi);
    uint32_t cur_lo = clo, cur_hi = chi;
    for (uint32_t round = 0; round < rounds && cerr; ++round) {
        float A = 0, B = 0, C = 0, P[3] = {0, 0, 0}, Q[3] = {0, 0, 0};
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const uint32_t x = __byte_perm(s_x[i], 0u, psel);
            const uint32_t k = ((i < 8 ? cur_lo : cur_hi) >> (4*(i & 7))) & 15u;
            const float w = static_cast<float>(index_weight(k, chalf, crecip))*(1.0f/64.0f), iw = 1.0f - w;
            A += iw*iw; B += iw*w; C += w*w;
            const float xs[3] = {static_cast<float>(x & 0xFFu), static_cast<float>((x >> 8) & 0xFFu), static_cast<float>((x >> 16) & 0xFFu)};
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) { P[ch] += iw*xs[ch]; Q[ch] += w*xs[ch]; }
        }
        const float det = A*C - B*B;
        if (!(fabsf(det) > 1e-4f)) break;
        const float id = 1.0f/det;
        const uint32_t n0 = quantize_endpoint(make_float4((C*P[0] - B*Q[0])*id, (C*P[1] - B*Q[1])*id, (C*P[2] - B*Q[2])*id, 255.0f), cbits, 0u, false, 0u, q);
        const uint32_t n1 = quantize_endpoint(make_float4((A*Q[0] - B*P[0])*id, (A*Q[1] - B*P[1])*id, (A*Q[2] - B*P[2])*id, 255.0f), cbits, 0u, false, 0u, q);
        ev = make_eval(n0, n1, cbits, 0u, cn);
        const uint32_t nerr = eval_colour3(s_x, psel, ev, cn, chalf, crecip, cmask & 0x00FFFFFFu, cur_lo, cur_hi);
        if (nerr < cerr) { cerr = nerr; e0 = n0; e1 = n1; clo = cur_lo; chi = cur_hi; }
    }

    // ---- scalar channel: range, quantise, indices, least squares
    const bool enabled = (cmask >> 24) != 0u;
    uint32_t a0 = scalar_code(static_cast<float>(amin), abits), a1 = scalar_code(static_cast<float>(amax), abits);
    uint32_t alo, ahi;
    uint32_t aerr = eval_scalar(s_x, psel, scalar_value(a0, abits), scalar_value(a1, abits), an, ahalf, arecip, enabled, alo, ahi);
    cur_lo = alo; cur_hi = ahi;
    for (uint32_t round = 0; round < rounds && aerr; ++round) {
        float A = 0, B = 0, C = 0, P = 0, Q = 0;
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const float v = static_cast<float>(__byte_perm(s_x[i], 0u, psel) >> 24);
            const uint32_t k = ((i < 8 ? cur_lo : cur_hi) >> (4*(i & 7))) & 15u;
            const float w = static_cast<float>(index_weight(k, ahalf, arecip))*(1.0f/64.0f), iw = 1.0f - w;
            A += iw*iw; B += iw*w; C += w*w; P += iw*v; Q += w*v;
        }
        const float det = A*C - B*B;
        if (!(fabsf(det) > 1e-4f)) break;
        const float id = 1.0f/det;
        const uint32_t n0 = scalar_code((C*P - B*Q)*id, abits), n1 = scalar_code((A*Q - B*P)*id, abits);
        const uint32_t nerr = eval_scalar(s_x, psel, scalar_value(n0, abits), scalar_value(n1, abits), an, ahalf, arecip, enabled, cur_lo, cur_hi);
        if (nerr < aerr) { aerr = nerr; a0 = n0; a1 = n1; alo = cur_lo; ahi = cur_hi; }
    }
    best.e0[0] = (e0 & 0x00FFFFFFu) | (a0 << 24); best.e1[0] = (e1 & 0x00FFFFFFu) | (a1 << 24);
    best.e0[1] = best.e1[1] = 0u;
    best.err[0] = cerr; best.err[1] = aerr;
    best.sel_lo = clo; best.sel_hi = chi; best.sel2_lo = alo; best.sel2_hi = ahi;
}

// Pack modes 4, 5.
CFX_HD_NOINLINE uint4 pack_dual(uint32_t mode, uint32_t rot, uint32_t idx_mode, Fit f)
{
    const uint32_t cbits = mode == 5 ? 7u : 5u, abits = mode == 5 ? 8u : 6u;
    const uint32_t cib = mode == 5 ? 2u : (idx_mode ? 3u : 2u), aib = mode == 5 ? 2u : (idx_mode ? 2u : 3u);
    uint64_t csel = (static_cast<uint64_t>(f.sel_hi) << 32) | f.sel_lo, asel = (static_cast<uint64_t>(f.sel2_hi) << 32) | f.sel2_lo;
    uint32_t c0 = f.e0[0] & 0x00FFFFFFu, c1 = f.e1[0] & 0x00FFFFFFu, a0 = f.e0[0] >> 24, a1 = f.e1[0] >> 24;
    // the first index of each set must have a clear MSB: swap that set's end points and invert its indices
    if ((static_cast<uint32_t>(csel) & 15u) >> (cib - 1u)) {
        const uint32_t t = c0; c0 = c1; c1 = t;
        csel = 0x1111111111111111ull*((1u << cib) - 1u) - csel;
    }
    if ((static_cast<uint32_t>(asel) & 15u) >> (aib - 1u)) {
        const uint32_t t = a0; a0 = a1; a1 = t;
        asel = 0x1111111111111111ull*((1u << aib) - 1u) - asel;
    }
    BitWriter bw; bw.init();
    bw.put(1u << mode, mode + 1u);
    bw.put(rot, 2);
    if (mode == 4) bw.put(idx_mode, 1);
#pragma unroll 1
    for (uint32_t ch = 0; ch < 3; ++ch) { bw.put((c0 >> (8*ch)) & 0xFFu, cbits); bw.put((c1 >> (8*ch)) & 0xFFu, cbits); }
    bw.put(a0, abits); bw.put(a1, abits);
    // mode 5: colour indices then scalar indices; mode 4: the 2-bit set then the 3-bit set, whoever owns them
    const bool colour_first = mode == 5 || idx_mode == 0;
    const uint64_t first = colour_first ? csel : asel, second = colour_first ? asel : csel;
    const uint32_t fb = colour_first ? cib : aib, sb = colour_first ? aib : cib;
#pragma unroll 1
    for (uint32_t i = 0; i < 16; ++i) bw.put(static_cast<uint32_t>(first >> (4*i)) & 15u, i == 0 ? fb - 1u : fb);
#pragma unroll 1
    for (uint32_t i = 0; i < 16; ++i) bw.put(static_cast<uint32_t>(second >> (4*i)) & 15u, i == 0 ? sb - 1u : sb);
    return make_uint4(static_cast<uint32_t>(bw.lo), static_cast<uint32_t>(bw.lo >> 32),
        static_cast<uint32_t>(bw.hi), static_cast<uint32_t>(bw.hi >> 32));
}

// Pack modes 1, 3, 6, 7 (one index set, p-bits).
CFX_HD_NOINLINE uint4 pack_block(uint32_t mode, uint32_t part, uint32_t m1, Fit f)
{
    const ModeInfo mi = mode_info(mode);
    const uint32_t nm1 = (1u << mi.ibits) - 1u;
    uint64_t sel = (static_cast<uint64_t>(f.sel_hi) << 32) | f.sel_lo;
    uint32_t anchor1 = mi.ns == 2 ? kBc7Anchor2[part] : 0u;
    // anchor indices must have a clear MSB: swap that subset's endpoints and invert its indices
    uint32_t msb = 1u << (mi.ibits - 1u);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        if (s >= static_cast<int>(mi.ns)) continue;
        uint32_t a = s ? anchor1 : 0u;
        if ((static_cast<uint32_t>(sel >> (4*a)) & 15u) & msb) {
            uint32_t t = f.e0[s]; f.e0[s] = f.e1[s]; f.e1[s] = t;
            uint32_t mask = s ? m1 : (~m1 & 0xFFFFu);
            uint64_t nm = (static_cast<uint64_t>(nibble_mask8(mask >> 8)) << 32) | nibble_mask8(mask);
            // idx -> nm1 - idx on that subset's texels
            uint64_t inv = (0x1111111111111111ull*nm1) & nm;
            sel = (sel & ~nm) | ((inv - (sel & nm)) & nm);
        }
    }
    BitWriter bw; bw.init();
    bw.put(1u << mode, mode + 1u);
    if (mi.ns == 2) bw.put(part, 6);
    const uint32_t pshift = 1u;   // every mode handled here has p-bits
#pragma unroll 1
    for (uint32_t ch = 0; ch < 4; ++ch) {
        uint32_t bits = ch == 3 ? mi.abits : mi.cbits;
        if (!bits) continue;
#pragma unroll 1
        for (uint32_t s = 0; s < mi.ns; ++s) {
            bw.put(((f.e0[s] >> (8*ch)) & 0xFFu) >> pshift, bits);
            bw.put(((f.e1[s] >> (8*ch)) & 0xFFu) >> pshift, bits);
        }
    }
#pragma unroll 1
    for (uint32_t s = 0; s < mi.ns; ++s) {
        if (mi.pmode == 2) bw.put(f.e0[s] & 1u, 1);
        else { bw.put(f.e0[s] & 1u, 1); bw.put(f.e1[s] & 1u, 1); }
    }
#pragma unroll 1
    for (uint32_t i = 0; i < 16; ++i) {
        uint32_t k = static_cast<uint32_t>(sel >> (4*i)) & 15u;
        bool is_anchor = i == 0 || (mi.ns == 2 && i == anchor1);
        bw.put(k, is_anchor ? mi.ibits - 1u : mi.ibits);
    }
    return make_uint4(static_cast<uint32_t>(bw.lo), static_cast<uint32_t>(bw.lo >> 32),
        static_cast<uint32_t>(bw.hi), static_cast<uint32_t>(bw.hi >> 32));
}

} // namespace bc7
} // namespace cfx
