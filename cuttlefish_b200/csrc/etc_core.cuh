// Lane-local ETC1 / ETC2 RGB / ETC2 RGBA8 (EAC alpha) encoder: ONE LANE OWNS ONE BLOCK.
//
// PARITY STATUS: our own search, held to "RGB(A) PSNR >= reference - 0.1 dB" against etc2comp as
// EtcConverter::process drives it (lib/src/EtcConverter.cpp:120-152: one Etc::Image::Encode per
// block; at Quality::Normal only encoding iteration 0 runs, lib/etc2comp/EtcLib/Etc/EtcImage.cpp:282).
// It is NOT bit-identical to etc2comp's float pipeline (Block4x4Encoding_ETC1::PerformFirstIteration,
// EtcBlock4x4Encoding_ETC1.cpp:311-338); see DESIGN.md.
//
// Search: both flips x {differential 555+333, individual 444+444}, each half fitted by a +-1
// descent around its mean over all 8 modifier tables with exact decoded error; ETC2 adds the planar
// mode (least-squares plane per channel, 676 quantisation, +-1 descent) and the T / H modes
// (two-colour clustering); EAC alpha searches table x multiplier x base around the block's range.
// Partial edge blocks: `vm` (bit t = texel t lies inside the image) -- the reference hands etc2comp a smaller image for
// them (lib/src/EtcConverter.cpp:122-150) and etc2comp gives the missing texels no weight (NaN alpha "border" pixels);
// here the texels outside the image (clamp-to-edge replicas in xs) are left out of every mean and every error sum and
// only receive selectors.
// Compiles for the device and, through hostdev.h, for tools/emu_etc.cpp.
#pragma once
#include "hostdev.h"

namespace cfx {
namespace etc {

CFX_CONST int16_t kMod[8][2] = {{2, 8}, {5, 17}, {9, 29}, {13, 42}, {18, 60}, {24, 80}, {33, 106}, {47, 183}};
CFX_CONST uint8_t kDist[8] = {3, 6, 11, 16, 23, 32, 41, 64};
CFX_CONST int8_t kEac[16][8] = {
    {-3, -6, -9, -15, 2, 5, 8, 14}, {-3, -7, -10, -13, 2, 6, 9, 12}, {-2, -5, -8, -13, 1, 4, 7, 12}, {-2, -4, -6, -13, 1, 3, 5, 12},
    {-3, -6, -8, -12, 2, 5, 7, 11}, {-3, -7, -9, -11, 2, 6, 8, 10}, {-4, -7, -8, -11, 3, 6, 7, 10}, {-3, -5, -8, -11, 2, 4, 7, 10},
    {-2, -6, -8, -10, 1, 5, 7, 9}, {-2, -5, -8, -10, 1, 4, 7, 9}, {-2, -4, -8, -10, 1, 3, 7, 9}, {-2, -5, -7, -10, 1, 4, 6, 9},
    {-3, -4, -7, -10, 2, 3, 6, 9}, {-1, -2, -3, -10, 0, 1, 2, 9}, {-4, -6, -8, -9, 3, 5, 7, 8}, {-3, -5, -7, -9, 2, 4, 6, 8}};

// texel t = y*4 + x, channel c (0..3), values 0..255 as floats
CFX_HD float& px(float* xs, uint32_t lane, uint32_t t, uint32_t c) { return xs[(t*4u + c)*32u + lane]; }

CFX_HD int clamp255(int v) { return min(max(v, 0), 255); }
CFX_HD int expand5(int v) { return (v << 3) | (v >> 2); }
CFX_HD int expand4(int v) { return (v << 4) | v; }
CFX_HD int expand6(int v) { return (v << 2) | (v >> 4); }
CFX_HD int expand7(int v) { return (v << 1) | (v >> 6); }

// T/H colour descent at the short search levels (up to Quality::Normal): off by default. Measured with it on
// (-DCFX_ETC_TH_AT_NORMAL=1, gate 1.1): ETC2 Normal +0.11 dB over etc2comp on noise+grad instead of -0.06, +1.12 instead
// of +1.04 on the screenshot probe, at 2.38 instead of 2.80 GTexel/s.
#ifndef CFX_ETC_TH_AT_NORMAL
#define CFX_ETC_TH_AT_NORMAL 0
#endif
#ifndef CFX_ETC_TH_GATE
#define CFX_ETC_TH_GATE 1.1f
#endif
#ifndef CFX_ETC_NARROW
#define CFX_ETC_NARROW 1
#endif
struct HalfFit { float err; uint32_t table; };

// The colour error metric.  Linear textures: plain squared RGB distance (etc2comp's RGBX).  sRGB textures (`perc`): the
// reference switches etc2comp to REC709 (lib/src/EtcConverter.cpp:61-88; Block4x4Encoding::CalcPixelError,
// EtcBlock4x4Encoding.cpp:157-180): 3 dL^2 + dCr^2 + 0.5 dCb^2 with L = 0.2126 r + 0.7152 g + 0.0722 b,
// Cr = 0.5 (r - L)/(1 - 0.2126), Cb = 0.5 (b - L)/(1 - 0.0722) -- a quadratic form d^T A d of the RGB difference d,
// A = M^T diag(3, 1, 0.5) M; its six coefficients below.  The search is the same, every error sum is taken in this metric.
namespace rec709 {
constexpr double kLr = 0.2126, kLg = 0.7152, kLb = 0.0722;
constexpr double kCr = 0.5/(1.0 - kLr), kCb = 0.5/(1.0 - kLb);
// rows of M: luma, chroma red, chroma blue
constexpr double m[3][3] = {{kLr, kLg, kLb}, {kCr*(1.0 - kLr), -kCr*kLg, -kCr*kLb}, {-kCb*kLr, -kCb*kLg, kCb*(1.0 - kLb)}};
constexpr double w[3] = {3.0, 1.0, 0.5};
constexpr double a(int i, int j) { return w[0]*m[0][i]*m[0][j] + w[1]*m[1][i]*m[1][j] + w[2]*m[2][i]*m[2][j]; }
constexpr float a00 = static_cast<float>(a(0, 0)), a11 = static_cast<float>(a(1, 1)), a22 = static_cast<float>(a(2, 2));
constexpr float a01 = static_cast<float>(a(0, 1)), a02 = static_cast<float>(a(0, 2)), a12 = static_cast<float>(a(1, 2));
}
// d^T A d
CFX_HD float perc_norm(float d0, float d1, float d2)
{
    return rec709::a00*d0*d0 + rec709::a11*d1*d1 + rec709::a22*d2*d2 + 2.0f*(rec709::a01*d0*d1 + rec709::a02*d0*d2 + rec709::a12*d1*d2);
}
CFX_HD float color_err(float d0, float d1, float d2, bool perc) { return perc ? perc_norm(d0, d1, d2) : d0*d0 + d1*d1 + d2*d2; }

// The four colours of modifier table tb around an 8-bit base colour (clamped), kept as -2 p_k and |p_k|^2: the error of
// texel x to colour k is |x|^2 + (|p_k|^2 - 2 p_k.x), three multiply-adds per candidate. For 8-bit sources every term is
// an integer below 2^24, so this is exactly the sum of squared differences.
// tmask != 0: punch-through block (ETC2 RGB8A1 with the opaque bit clear): texels of tmask take selector 2
// (transparent) at no cost, the others choose among {+0, +big, -big} (selectors 0, 1, 3).
struct TableColours { float n0[4], n1[4], n2[4], cc[4]; };

template <bool PERC>
CFX_HD void table_colours(const int* base, uint32_t tb, bool punch, TableColours& p)
{
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k) {
        int m = (k & 2u) ? -static_cast<int>(kMod[tb][k & 1u]) : static_cast<int>(kMod[tb][k & 1u]);
        if (punch && k == 0u) m = 0;
        const float p0 = static_cast<float>(clamp255(base[0] + m)), p1 = static_cast<float>(clamp255(base[1] + m)),
            p2 = static_cast<float>(clamp255(base[2] + m));
        if (PERC) {
            // |p - x|_A^2 = p^T A p - 2 (A p).x + x^T A x
            const float q0 = rec709::a00*p0 + rec709::a01*p1 + rec709::a02*p2, q1 = rec709::a01*p0 + rec709::a11*p1 + rec709::a12*p2,
                q2 = rec709::a02*p0 + rec709::a12*p1 + rec709::a22*p2;
            p.n0[k] = -2.0f*q0; p.n1[k] = -2.0f*q1; p.n2[k] = -2.0f*q2;
            p.cc[k] = p0*q0 + p1*q1 + p2*q2;
            continue;
        }
        p.n0[k] = -2.0f*p0; p.n1[k] = -2.0f*p1; p.n2[k] = -2.0f*p2;
        p.cc[k] = p0*p0 + p1*p1 + p2*p2;
    }
}

// Best modifier table (of tb0 .. tb1) of one half (texel mask) for an 8-bit base colour, and its error. The selectors are
// not kept: most fits lose, half_selectors() recomputes them for the one that ends up in the block.
// (PERC is a template parameter: as a run-time flag in this loop it cost the linear path 11 %)
template <bool PERC>
CFX_HD void half_fit_t(float* xs, uint32_t lane, uint32_t mask, const int* base, float limit, HalfFit& out, uint32_t tmask,
    uint32_t tb0, uint32_t tb1)
{
    out.err = 3.0e38f; out.table = 0;
#pragma unroll 1
    for (uint32_t tb = tb0; tb <= tb1; ++tb) {
        TableColours p;
        table_colours<PERC>(base, tb, tmask != 0, p);
        float err = 0.0f;
        for (uint32_t left = mask & ~tmask; left; left &= left - 1u) {
            const uint32_t t = static_cast<uint32_t>(__ffs(left)) - 1u;
            const float x0 = px(xs, lane, t, 0), x1 = px(xs, lane, t, 1), x2 = px(xs, lane, t, 2);
            float be = p.cc[0] + p.n0[0]*x0 + p.n1[0]*x1 + p.n2[0]*x2;
#pragma unroll
            for (uint32_t k = 1; k < 4; ++k) {
                if (tmask && k == 2u) continue;
                be = fminf(be, p.cc[k] + p.n0[k]*x0 + p.n1[k]*x1 + p.n2[k]*x2);
            }
            err += fmaxf(be + (PERC ? perc_norm(x0, x1, x2) : x0*x0 + x1*x1 + x2*x2), 0.0f);
            if (err >= out.err || err >= limit) break;
        }
        if (err < out.err) { out.err = err; out.table = tb; }
    }
}

CFX_HD void half_fit(float* xs, uint32_t lane, uint32_t mask, const int* base, float limit, HalfFit& out, uint32_t tmask = 0,
    uint32_t tb0 = 0, uint32_t tb1 = 7, bool perc = false)
{
    if (perc) half_fit_t<true>(xs, lane, mask, base, limit, out, tmask, tb0, tb1);
    else half_fit_t<false>(xs, lane, mask, base, limit, out, tmask, tb0, tb1);
}

// Selectors (2 bits per texel t, only the half's texels) of a base colour and table: the first minimum over k.
template <bool PERC = false>
CFX_HD uint32_t half_selectors(float* xs, uint32_t lane, uint32_t mask, const int* base, uint32_t tb, uint32_t tmask = 0)
{
    TableColours p;
    table_colours<PERC>(base, tb, tmask != 0, p);
    uint32_t sel = 0;
    for (uint32_t left = mask; left; left &= left - 1u) {
        const uint32_t t = static_cast<uint32_t>(__ffs(left)) - 1u;
        if ((tmask >> t) & 1u) { sel |= 2u << (2*t); continue; }
        const float x0 = px(xs, lane, t, 0), x1 = px(xs, lane, t, 1), x2 = px(xs, lane, t, 2);
        float be = 3.0e38f;
        uint32_t bk = 0;
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {
            if (tmask && k == 2u) continue;
            const float e = p.cc[k] + p.n0[k]*x0 + p.n1[k]*x1 + p.n2[k]*x2;
            if (e < be) { be = e; bk = k; }
        }
        sel |= bk << (2*t);
    }
    return sel;
}

// Descent of one half's quantised base colour (bits = 4 or 5) within [lo, hi] per channel.
template <bool PERC = false>
CFX_HD void half_search(float* xs, uint32_t lane, uint32_t mask, int bits, int* q /* in/out */, const int* lo, const int* hi,
    int rounds, HalfFit& best, uint32_t tmask = 0)
{
    int base[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { q[c] = min(max(q[c], lo[c]), hi[c]); base[c] = bits == 5 ? expand5(q[c]) : expand4(q[c]); }
    half_fit_t<PERC>(xs, lane, mask, base, 3.0e38f, best, tmask, 0, 7);
    for (int round = 0; round < rounds && best.err > 0.0f; ++round) {
        bool improved = false;
#pragma unroll 1
        const int moves = rounds >= 5 ? 20 : 8;          // Quality::Highest also moves two channels at once
        for (int k = 0; k < moves; ++k) {
            // k 0,1: all channels -1 / +1 (luma); k 2..7: one channel -1 / +1; k 8..19: two channels, the four sign pairs
            int t[3] = {q[0], q[1], q[2]};
            const int d = (k & 1) ? 1 : -1;
            if (k < 2) { t[0] += d; t[1] += d; t[2] += d; }
            else if (k < 8) t[(k - 2) >> 1] += d;
            else {
                const int pr = (k - 8) >> 2, c1 = pr == 2 ? 1 : 0, c2 = pr == 0 ? 1 : 2;
                t[c1] += d; t[c2] += (k & 2) ? 1 : -1;
            }
            if (t[0] < lo[0] || t[0] > hi[0] || t[1] < lo[1] || t[1] > hi[1] || t[2] < lo[2] || t[2] > hi[2]) continue;
#pragma unroll
            for (int c = 0; c < 3; ++c) base[c] = bits == 5 ? expand5(t[c]) : expand4(t[c]);
            HalfFit f;
            // a one-step move of the base colour rarely changes the best table by more than one: the short descents
            // (up to Quality::Normal) only look at the incumbent's neighbours, the long ones at all eight
            const bool narrow = CFX_ETC_NARROW && rounds <= 1;
            half_fit_t<PERC>(xs, lane, mask, base, best.err, f, tmask, narrow ? (best.table ? best.table - 1u : 0u) : 0u,
                narrow ? min(best.table + 1u, 7u) : 7u);
            if (f.err < best.err) { best = f; q[0] = t[0]; q[1] = t[1]; q[2] = t[2]; improved = true; }
        }
        if (!improved) break;
    }
}

CFX_HD void put_be(uint32_t& hi, uint32_t& lo, int bit /* 63..0 */, uint32_t v, int n)
{
    // field occupies bits [bit, bit-n+1] of the 64-bit big-endian word (hi = bits 63..32)
    for (int i = 0; i < n; ++i) {
        const int b = bit - i;
        const uint32_t one = (v >> (n - 1 - i)) & 1u;
        if (b >= 32) hi |= one << (b - 32); else lo |= one << b;
    }
}

// pixel index bits: selector k of texel t=(y*4+x) goes to pixel p = x*4 + y: MSB at bit 16+p, LSB at bit p
CFX_HD uint32_t pixel_bits(uint32_t sel)
{
    uint32_t out = 0;
#pragma unroll
    for (uint32_t t = 0; t < 16; ++t) {
        const uint32_t k = (sel >> (2*t)) & 3u, p = (t & 3u)*4u + (t >> 2);
        out |= ((k >> 1) & 1u) << (16u + p);
        out |= (k & 1u) << p;
    }
    return out;
}

CFX_HD uint2 to_bytes(uint32_t hi, uint32_t lo)
{
    // the block is stored big-endian: byte 0 = bits 63..56
    return make_uint2(__byte_perm(hi, 0, 0x0123), __byte_perm(lo, 0, 0x0123));
}

struct ColorResult { float err; uint32_t hi, lo; };

// ---- ETC1 part: both flips, differential and individual -----------------------------------------
// diff_only: no individual (444+444) mode -- ETC2 RGB8A1, where that bit is the opaque flag.
// tmask != 0: punch-through block of RGB8A1 (opaque flag clear, see half_fit).
template <bool PERC = false>
CFX_HD void encode_etc1(float* xs, uint32_t lane, int rounds, ColorResult& out, bool diff_only = false, uint32_t tmask = 0,
    uint32_t vm = 0xFFFFu)
{
    out.err = 3.0e38f; out.hi = out.lo = 0;
    const uint32_t skip = tmask | (~vm & 0xFFFFu);       // texels without a say in the base colours
#pragma unroll 1
    for (uint32_t flip = 0; flip < 2; ++flip) {
        const uint32_t maskA = flip ? 0x00FFu : 0x3333u, maskB = ~maskA & 0xFFFFu;
        float mA[3] = {0, 0, 0}, mB[3] = {0, 0, 0}, nA = 0.0f, nB = 0.0f;
        for (uint32_t t = 0; t < 16; ++t) {
            if ((skip >> t) & 1u) continue;
            const bool a = (maskA >> t) & 1u;
            if (a) nA += 1.0f; else nB += 1.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) { const float v = px(xs, lane, t, c); if (a) mA[c] += v; else mB[c] += v; }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) { mA[c] *= nA > 0.0f ? 1.0f/nA : 0.0f; mB[c] *= nB > 0.0f ? 1.0f/nB : 0.0f; }
        if (skip) {                  // a half without opaque (or inside-the-image) texels follows the other one
#pragma unroll
            for (int c = 0; c < 3; ++c) { if (nA == 0.0f) mA[c] = mB[c]; if (nB == 0.0f) mB[c] = mA[c]; }
        }
#pragma unroll 1
        for (int diff = 1; diff >= (diff_only ? 1 : 0); --diff) {
            const int bits = diff ? 5 : 4, maxq = diff ? 31 : 15;
            int qA[3], qB[3], lo[3] = {0, 0, 0}, hi[3] = {maxq, maxq, maxq};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                qA[c] = __float2int_rn(mA[c]*static_cast<float>(maxq)*(1.0f/255.0f));
                qB[c] = __float2int_rn(mB[c]*static_cast<float>(maxq)*(1.0f/255.0f));
            }
            if (diff) {
                // pull the two bases together until every delta fits in [-4, 3]
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    int d = qB[c] - qA[c];
                    while (d > 3) { if ((d & 1) == 0) ++qA[c]; else --qB[c]; d = qB[c] - qA[c]; }
                    while (d < -4) { if ((d & 1) == 0) --qA[c]; else ++qB[c]; d = qB[c] - qA[c]; }
                }
            }
            HalfFit fA, fB;
            half_search<PERC>(xs, lane, maskA & vm, bits, qA, lo, hi, rounds, fA, tmask);
            if (diff) {
#pragma unroll
                for (int c = 0; c < 3; ++c) { lo[c] = max(qA[c] - 4, 0); hi[c] = min(qA[c] + 3, 31); }
            }
            half_search<PERC>(xs, lane, maskB & vm, bits, qB, lo, hi, rounds, fB, tmask);
            const float err = fA.err + fB.err;
            if (err < out.err) {
                uint32_t h = 0, l = 0;
                if (diff) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        put_be(h, l, 63 - 8*c, static_cast<uint32_t>(qA[c]), 5);
                        put_be(h, l, 58 - 8*c, static_cast<uint32_t>(qB[c] - qA[c]) & 7u, 3);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        put_be(h, l, 63 - 8*c, static_cast<uint32_t>(qA[c]), 4);
                        put_be(h, l, 59 - 8*c, static_cast<uint32_t>(qB[c]), 4);
                    }
                }
                put_be(h, l, 39, fA.table, 3);
                put_be(h, l, 36, fB.table, 3);
                put_be(h, l, 33, tmask ? 0u : static_cast<uint32_t>(diff), 1);      // RGB8A1: this bit is the opaque flag
                put_be(h, l, 32, flip, 1);
                int bA[3], bB[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) { bA[c] = diff ? expand5(qA[c]) : expand4(qA[c]); bB[c] = diff ? expand5(qB[c]) : expand4(qB[c]); }
                l = pixel_bits(half_selectors<PERC>(xs, lane, maskA, bA, fA.table, tmask) | half_selectors<PERC>(xs, lane, maskB, bB, fB.table, tmask));
                out.err = err; out.hi = h; out.lo = l;
            }
        }
    }
}

// ---- ETC2 planar ---------------------------------------------------------------------------------
template <bool PERC>
CFX_HD float planar_error_t(float* xs, uint32_t lane, const int* O, const int* H, const int* V, uint32_t vm)
{
    float err = 0.0f;
    for (uint32_t t = 0; t < 16; ++t) {
        if (!((vm >> t) & 1u)) continue;
        const int x = t & 3, y = t >> 2;
        float d[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int v = clamp255((x*(H[c] - O[c]) + y*(V[c] - O[c]) + 4*O[c] + 2) >> 2);
            d[c] = static_cast<float>(v) - px(xs, lane, t, c);
            if (!PERC) err += d[c]*d[c];
        }
        if (PERC) err += perc_norm(d[0], d[1], d[2]);
    }
    return err;
}

CFX_HD float planar_error(float* xs, uint32_t lane, const int* O, const int* H, const int* V, uint32_t vm = 0xFFFFu, bool perc = false)
{
    return perc ? planar_error_t<true>(xs, lane, O, H, V, vm) : planar_error_t<false>(xs, lane, O, H, V, vm);
}

template <bool PERC = false>
CFX_HD void encode_planar(float* xs, uint32_t lane, int rounds, ColorResult& out, uint32_t vm = 0xFFFFu)
{
    int q[9];     // RO GO BO RH GH BH RV GV BV (6/7/6 bits)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float m = 0, sx = 0, sy = 0;
        for (uint32_t t = 0; t < 16; ++t) {
            const float v = px(xs, lane, t, c);
            m += v; sx += (static_cast<float>(t & 3) - 1.5f)*v; sy += (static_cast<float>(t >> 2) - 1.5f)*v;
        }
        m *= (1.0f/16.0f); sx *= (1.0f/20.0f); sy *= (1.0f/20.0f);
        const float o = m - 1.5f*sx - 1.5f*sy, h = o + 4.0f*sx, v = o + 4.0f*sy;
        const float scale = c == 1 ? 127.0f/255.0f : 63.0f/255.0f;
        const int maxq = c == 1 ? 127 : 63;
        q[c] = min(max(__float2int_rn(o*scale), 0), maxq);
        q[3 + c] = min(max(__float2int_rn(h*scale), 0), maxq);
        q[6 + c] = min(max(__float2int_rn(v*scale), 0), maxq);
    }
    int O[3], H[3], V[3];
    auto expand_all = [&]() {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            O[c] = c == 1 ? expand7(q[c]) : expand6(q[c]);
            H[c] = c == 1 ? expand7(q[3 + c]) : expand6(q[3 + c]);
            V[c] = c == 1 ? expand7(q[6 + c]) : expand6(q[6 + c]);
        }
    };
    expand_all();
    float best = planar_error_t<PERC>(xs, lane, O, H, V, vm);
    for (int round = 0; round < rounds && best > 0.0f; ++round) {
        bool improved = false;
#pragma unroll 1
        for (int k = 0; k < 18; ++k) {
            const int i = k >> 1, d = (k & 1) ? 1 : -1;
            const int maxq = (i % 3) == 1 ? 127 : 63;
            if (q[i] + d < 0 || q[i] + d > maxq) continue;
            q[i] += d;
            expand_all();
            const float e = planar_error_t<PERC>(xs, lane, O, H, V, vm);
            if (e < best) { best = e; improved = true; } else q[i] -= d;
        }
        if (!improved) break;
    }
    out.err = best;
    const uint32_t RO = q[0], GO = q[1], BO = q[2], RH = q[3], GH = q[4], BH = q[5], RV = q[6], GV = q[7], BV = q[8];
    uint32_t h = 0, l = 0;
    // R: bits 62..57 = RO; no overflow of (63..59) + signed(58..56)
    put_be(h, l, 62, RO, 6);
    put_be(h, l, 56, GO >> 6, 1);
    {
        const int dr = static_cast<int>(((RO & 3u) << 1) | (GO >> 6));          // bits 58..56 as signed 3
        put_be(h, l, 63, (dr & 4) ? 1u : 0u, 1);
    }
    put_be(h, l, 54, GO & 63u, 6);
    put_be(h, l, 48, BO >> 5, 1);
    {
        const int dg = static_cast<int>(((GO & 3u) << 1) | (BO >> 5));
        put_be(h, l, 55, (dg & 4) ? 1u : 0u, 1);
    }
    put_be(h, l, 44, (BO >> 3) & 3u, 2);
    put_be(h, l, 41, BO & 7u, 3);
    {
        // B must overflow: 5-bit (47..43) + signed 3-bit (42..40)
        const uint32_t a = (BO >> 3) & 3u, b = (BO >> 1) & 3u;
        if (a + b < 4) { put_be(h, l, 47, 0u, 3); put_be(h, l, 42, 1u, 1); }
        else { put_be(h, l, 47, 7u, 3); put_be(h, l, 42, 0u, 1); }
    }
    put_be(h, l, 38, RH >> 1, 5);
    put_be(h, l, 33, 1u, 1);
    put_be(h, l, 32, RH & 1u, 1);
    put_be(h, l, 31, GH, 7);
    put_be(h, l, 24, BH, 6);
    put_be(h, l, 18, RV, 6);
    put_be(h, l, 12, GV, 7);
    put_be(h, l, 5, BV, 6);
    out.hi = h; out.lo = l;
}

// ---- ETC2 T and H modes --------------------------------------------------------------------------
// Two 444 colours from a 2-means split of the block along its principal axis.
// Error and selectors of one T / H configuration: kind 0: T with A single, B +-d; kind 1: T with B single, A +-d;
// kind 2: H.  qA, qB: the two RGB444 colours; di: distance index.  Stops early once `limit` is exceeded.
template <bool PERC>
CFX_HD float th_eval_t(float* xs, uint32_t lane, uint32_t kind, uint32_t di, const int* qA, const int* qB, float limit, uint32_t& sel_out,
    uint32_t vm)
{
    const int d = kDist[di];
    int pal[4][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int eA = expand4(qA[c]), eB = expand4(qB[c]);
        const int s = kind == 1 ? eB : eA, o = kind == 1 ? eA : eB;
        if (kind < 2) { pal[0][c] = s; pal[1][c] = clamp255(o + d); pal[2][c] = o; pal[3][c] = clamp255(o - d); }
        else { pal[0][c] = clamp255(eA + d); pal[1][c] = clamp255(eA - d); pal[2][c] = clamp255(eB + d); pal[3][c] = clamp255(eB - d); }
    }
    float err = 0.0f;
    uint32_t sel = 0;
    for (uint32_t t = 0; t < 16 && err < limit; ++t) {
        float be = 3.0e38f;
        uint32_t bk = 0;
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {
            const float d0 = static_cast<float>(pal[k][0]) - px(xs, lane, t, 0), d1 = static_cast<float>(pal[k][1]) - px(xs, lane, t, 1),
                d2 = static_cast<float>(pal[k][2]) - px(xs, lane, t, 2);
            const float e = PERC ? perc_norm(d0, d1, d2) : d0*d0 + d1*d1 + d2*d2;
            if (e < be) { be = e; bk = k; }
        }
        if ((vm >> t) & 1u) err += be;
        sel |= bk << (2*t);
    }
    sel_out = sel;
    return err;
}

CFX_HD float th_eval(float* xs, uint32_t lane, uint32_t kind, uint32_t di, const int* qA, const int* qB, float limit, uint32_t& sel_out,
    uint32_t vm = 0xFFFFu, bool perc = false)
{
    return perc ? th_eval_t<true>(xs, lane, kind, di, qA, qB, limit, sel_out, vm) : th_eval_t<false>(xs, lane, kind, di, qA, qB, limit, sel_out, vm);
}

// rounds > 0 (Quality::High and up): +-1 descent on the six RGB444 components of the winner, distance index +-1
// (etc2comp widens its T / H search the same way in its later iterations, EtcBlock4x4Encoding_RGB8.cpp:370-...).
// rounds: +-1 descent rounds over the two colours; the descent only runs while the T/H error is below `gate` (the short
// searches pass a small multiple of the incumbent's error: a T/H block that far behind will not win).
template <bool PERC = false>
CFX_HD void encode_th(float* xs, uint32_t lane, ColorResult& out, int rounds = 0, float gate = 3.0e38f, uint32_t vm = 0xFFFFu)
{
    out.err = 3.0e38f; out.hi = out.lo = 0;
    float m[3] = {0, 0, 0};
    for (uint32_t t = 0; t < 16; ++t) { m[0] += px(xs, lane, t, 0); m[1] += px(xs, lane, t, 1); m[2] += px(xs, lane, t, 2); }
    m[0] *= (1.0f/16.0f); m[1] *= (1.0f/16.0f); m[2] *= (1.0f/16.0f);
    float cv[6] = {0, 0, 0, 0, 0, 0};
    for (uint32_t t = 0; t < 16; ++t) {
        const float d0 = px(xs, lane, t, 0) - m[0], d1 = px(xs, lane, t, 1) - m[1], d2 = px(xs, lane, t, 2) - m[2];
        cv[0] += d0*d0; cv[1] += d0*d1; cv[2] += d0*d2; cv[3] += d1*d1; cv[4] += d1*d2; cv[5] += d2*d2;
    }
    float v[3] = {cv[0], cv[1], cv[2]};
    float bestd = cv[0];
    if (cv[3] > bestd) { bestd = cv[3]; v[0] = cv[1]; v[1] = cv[3]; v[2] = cv[4]; }
    if (cv[5] > bestd) { bestd = cv[5]; v[0] = cv[2]; v[1] = cv[4]; v[2] = cv[5]; }
    for (int it = 0; it < 4; ++it) {
        const float n2 = v[0]*v[0] + v[1]*v[1] + v[2]*v[2];
        const float s = n2 > 1e-20f ? rsqrtf(n2) : 0.0f;
        const float a0 = v[0]*s, a1 = v[1]*s, a2 = v[2]*s;
        v[0] = cv[0]*a0 + cv[1]*a1 + cv[2]*a2; v[1] = cv[1]*a0 + cv[3]*a1 + cv[4]*a2; v[2] = cv[2]*a0 + cv[4]*a1 + cv[5]*a2;
    }
    // split at the mean projection, two Lloyd iterations
    uint32_t side = 0;
    for (uint32_t t = 0; t < 16; ++t) {
        const float p = (px(xs, lane, t, 0) - m[0])*v[0] + (px(xs, lane, t, 1) - m[1])*v[1] + (px(xs, lane, t, 2) - m[2])*v[2];
        if (p > 0.0f) side |= 1u << t;
    }
    float cA[3] = {m[0], m[1], m[2]}, cB[3] = {m[0], m[1], m[2]};
    for (int it = 0; it < 3; ++it) {
        float sA[3] = {0, 0, 0}, sB[3] = {0, 0, 0}, nA = 0, nB = 0;
        for (uint32_t t = 0; t < 16; ++t) {
            const bool b = (side >> t) & 1u;
#pragma unroll
            for (int c = 0; c < 3; ++c) { const float x = px(xs, lane, t, c); if (b) sB[c] += x; else sA[c] += x; }
            if (b) nB += 1.0f; else nA += 1.0f;
        }
        if (nA == 0.0f || nB == 0.0f) return;
#pragma unroll
        for (int c = 0; c < 3; ++c) { cA[c] = sA[c]/nA; cB[c] = sB[c]/nB; }
        side = 0;
        for (uint32_t t = 0; t < 16; ++t) {
            float dA = 0, dB = 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) { const float x = px(xs, lane, t, c); dA += (x - cA[c])*(x - cA[c]); dB += (x - cB[c])*(x - cB[c]); }
            if (dB < dA) side |= 1u << t;
        }
    }
    int qA[3], qB[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        qA[c] = min(max(__float2int_rn(cA[c]*(15.0f/255.0f)), 0), 15); qB[c] = min(max(__float2int_rn(cB[c]*(15.0f/255.0f)), 0), 15);
    }
    float best = 3.0e38f;
    uint32_t best_kind = 0, best_d = 0, best_sel = 0;
#pragma unroll 1
    for (uint32_t kind = 0; kind < 3; ++kind) {
#pragma unroll 1
        for (uint32_t di = 0; di < 8; ++di) {
            uint32_t sel;
            const float err = th_eval_t<PERC>(xs, lane, kind, di, qA, qB, best, sel, vm);
            if (err < best) { best = err; best_kind = kind; best_d = di; best_sel = sel; }
        }
    }
    for (int round = 0; round < rounds && best > 0.0f && best < gate; ++round) {
        bool improved = false;
#pragma unroll 1
        for (int k = 0; k < 12; ++k) {
            int tA[3] = {qA[0], qA[1], qA[2]}, tB[3] = {qB[0], qB[1], qB[2]};
            int* tgt = k < 6 ? tA : tB;
            const int c = (k % 6) >> 1, d = (k & 1) ? 1 : -1;
            if (tgt[c] + d < 0 || tgt[c] + d > 15) continue;
            tgt[c] += d;
#pragma unroll 1
            for (int dd = -1; dd <= 1; ++dd) {
                const int di = static_cast<int>(best_d) + dd;
                if (di < 0 || di > 7) continue;
                uint32_t sel;
                const float err = th_eval_t<PERC>(xs, lane, best_kind, static_cast<uint32_t>(di), tA, tB, best, sel, vm);
                if (err < best) {
                    best = err; best_d = static_cast<uint32_t>(di); best_sel = sel; improved = true;
#pragma unroll
                    for (int q = 0; q < 3; ++q) { qA[q] = tA[q]; qB[q] = tB[q]; }
                }
            }
        }
        if (!improved) break;
    }
    if (best >= 3.0e38f) return;
    out.err = best;
    uint32_t h = 0, l = 0;
    if (best_kind < 2) {
        const int* s = best_kind == 1 ? qB : qA;
        const int* o = best_kind == 1 ? qA : qB;
        const uint32_t R1 = s[0], a = R1 >> 2, b = R1 & 3u;
        if (a + b < 4) { put_be(h, l, 63, 0u, 3); put_be(h, l, 58, 1u, 1); } else { put_be(h, l, 63, 7u, 3); put_be(h, l, 58, 0u, 1); }
        put_be(h, l, 60, a, 2); put_be(h, l, 57, b, 2);
        put_be(h, l, 55, static_cast<uint32_t>(s[1]), 4); put_be(h, l, 51, static_cast<uint32_t>(s[2]), 4);
        put_be(h, l, 47, static_cast<uint32_t>(o[0]), 4); put_be(h, l, 43, static_cast<uint32_t>(o[1]), 4); put_be(h, l, 39, static_cast<uint32_t>(o[2]), 4);
        put_be(h, l, 35, best_d >> 1, 2); put_be(h, l, 33, 1u, 1); put_be(h, l, 32, best_d & 1u, 1);
        l = pixel_bits(best_sel);
    } else {
        // H: the low bit of the distance index is carried by the ORDER of the two colours
        int c1[3] = {qA[0], qA[1], qA[2]}, c2[3] = {qB[0], qB[1], qB[2]};
        uint32_t sel = best_sel;
        const uint32_t v1 = (c1[0] << 8) | (c1[1] << 4) | c1[2], v2 = (c2[0] << 8) | (c2[1] << 4) | c2[2];
        const bool want_ge = best_d & 1u;
        if (v1 == v2 && !want_ge) { out.err = 3.0e38f; return; }        // cannot express an even index with equal colours
        if ((v1 >= v2) != want_ge) {
#pragma unroll
            for (int c = 0; c < 3; ++c) { const int tmp = c1[c]; c1[c] = c2[c]; c2[c] = tmp; }
            // colours swapped: selectors 0,1 <-> 2,3
            sel ^= 0xAAAAAAAAu;
        }
        const uint32_t R1 = c1[0], G1 = c1[1], B1 = c1[2];
        put_be(h, l, 62, R1, 4);
        put_be(h, l, 58, G1 >> 1, 3);
        put_be(h, l, 63, (G1 & 8u) ? 1u : 0u, 1);          // keep R + dR in range
        put_be(h, l, 52, G1 & 1u, 1);
        put_be(h, l, 51, B1 >> 3, 1);
        put_be(h, l, 49, B1 & 7u, 3);
        {
            const uint32_t a = ((G1 & 1u) << 1) | (B1 >> 3), b = (B1 >> 1) & 3u;
            if (a + b < 4) { put_be(h, l, 55, 0u, 3); put_be(h, l, 50, 1u, 1); } else { put_be(h, l, 55, 7u, 3); put_be(h, l, 50, 0u, 1); }
        }
        put_be(h, l, 46, static_cast<uint32_t>(c2[0]), 4); put_be(h, l, 42, static_cast<uint32_t>(c2[1]), 4); put_be(h, l, 38, static_cast<uint32_t>(c2[2]), 4);
        put_be(h, l, 34, best_d >> 2, 1); put_be(h, l, 33, 1u, 1); put_be(h, l, 32, (best_d >> 1) & 1u, 1);
        l = pixel_bits(sel);
    }
    out.hi = h; out.lo = l;
}

// ---- EAC alpha (ETC2 RGBA8) ----------------------------------------------------------------------
CFX_HD uint2 encode_eac_alpha(float* xs, uint32_t lane, int radius, uint32_t vm = 0xFFFFu)
{
    float lo = 3.0e38f, hi = -3.0e38f;
    for (uint32_t t = 0; t < 16; ++t) { if (!((vm >> t) & 1u)) continue; const float a = px(xs, lane, t, 3); lo = fminf(lo, a); hi = fmaxf(hi, a); }
    uint32_t best_base = 255, best_mul = 1, best_tab = 13;
    float best = 3.0e38f;
    if (lo == hi && lo == 255.0f) {
        best = 0.0f;                 // opaque: base 255, table 13 has a zero modifier at selector 4
    } else {
#pragma unroll 1
        for (uint32_t tab = 0; tab < 16; ++tab) {
            const float tmin = static_cast<float>(kEac[tab][3]), tmax = static_cast<float>(kEac[tab][7]);
            const float range = tmax - tmin;
            const int mul0 = min(max(__float2int_rn((hi - lo)/range), 1), 15);
#pragma unroll 1
            for (int dm = -1; dm <= 1; ++dm) {
                const int mul = mul0 + dm;
                if (mul < 1 || mul > 15) continue;
                // centre the table's span on the block's alpha range
                const int base0 = __float2int_rn(0.5f*(lo + hi) - 0.5f*(tmin + tmax)*static_cast<float>(mul));
#pragma unroll 1
                for (int db = -radius; db <= radius; ++db) {
                    const int base = base0 + db;
                    if (base < 0 || base > 255) continue;
                    float err = 0.0f;
                    for (uint32_t t = 0; t < 16 && err < best; ++t) {
                        if (!((vm >> t) & 1u)) continue;
                        const float a = px(xs, lane, t, 3);
                        float be = 3.0e38f;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float d = static_cast<float>(clamp255(base + kEac[tab][k]*mul)) - a;
                            be = fminf(be, d*d);
                        }
                        err += be;
                    }
                    if (err < best) { best = err; best_base = base; best_mul = mul; best_tab = tab; }
                }
            }
        }
    }
    // selectors for the winner; 48 bits, pixel p = x*4 + y first (most significant)
    uint64_t bits = 0;
    for (uint32_t p = 0; p < 16; ++p) {
        const uint32_t t = (p & 3u)*4u + (p >> 2);
        const float a = px(xs, lane, t, 3);
        float be = 3.0e38f;
        uint32_t bk = 0;
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) {
            const float d = static_cast<float>(clamp255(static_cast<int>(best_base) + kEac[best_tab][k]*static_cast<int>(best_mul))) - a;
            if (d*d < be) { be = d*d; bk = k; }
        }
        bits = (bits << 3) | bk;
    }
    const uint32_t hi32 = (best_base << 24) | (best_mul << 20) | (best_tab << 16) | static_cast<uint32_t>(bits >> 32);
    const uint32_t lo32 = static_cast<uint32_t>(bits);
    return to_bytes(hi32, lo32);
}


// ---- ETC2 RGB8A1 (punch-through alpha) --------------------------------------------------------------
// Replaces Block4x4Encoding_RGB8A1 / _Opaque / _Transparent (lib/etc2comp/EtcLib/EtcCodec/EtcBlock4x4Encoding_RGB8A1.cpp;
// EtcConverter picks it for ETC2_R8G8B8A1, lib/src/EtcConverter.cpp:76-88).  Texels with alpha < 0.5 are transparent
// (same threshold, :96, :754).  Opaque blocks: the ETC2 RGB search without the individual mode (that bit is the
// opaque flag); mixed blocks: differential mode with the opaque flag clear, transparent texels on selector 2 and the
// others on {+0, +big, -big}; fully transparent blocks: selector 2 everywhere.  Our own search: PSNR parity.
template <bool PERC = false>
CFX_HD uint2 encode_color_a1(float* xs, uint32_t lane, int rounds, uint32_t vm = 0xFFFFu)
{
    uint32_t tmask = 0;
    for (uint32_t t = 0; t < 16; ++t) if (px(xs, lane, t, 3) < 127.5f) tmask |= 1u << t;
    ColorResult best;
    if (tmask == 0xFFFFu) {
        uint32_t sel = 0;
        for (uint32_t t = 0; t < 16; ++t) sel |= 2u << (2*t);
        return to_bytes(0u, pixel_bits(sel));
    }
    encode_etc1<PERC>(xs, lane, rounds, best, true, tmask, vm);
    if (tmask == 0 && best.err > 0.0f) {
        ColorResult r;
        encode_planar<PERC>(xs, lane, rounds, r, vm);
        if (r.err < best.err) best = r;
        if (best.err > 0.0f) {
            encode_th<PERC>(xs, lane, r, rounds >= (CFX_ETC_TH_AT_NORMAL ? 1 : 2) ? rounds : 0, rounds >= 2 ? 3.0e38f : best.err*CFX_ETC_TH_GATE, vm);
            if (r.err < best.err) best = r;
        }
    }
    return to_bytes(best.hi, best.lo);
}

// ---- EAC R11 / RG11 (one or two 11-bit channels, unsigned or signed) --------------------------------
// Replaces Block4x4Encoding_R11 / _RG11 (lib/etc2comp/EtcLib/EtcCodec/EtcBlock4x4Encoding_R11.cpp:170-392),
// which EtcConverter selects for EAC_R11 / EAC_R11G11 (lib/src/EtcConverter.cpp:89-115; signed inputs are remapped
// to [0,1] there, :139-143).  Same table x multiplier x base search as the alpha block above, but scored against the
// 11-bit decode of the format: unsigned clamp(base*8 + 4 + modifier*multiplier*8, 0, 2047) / 2047, signed
// clamp(base*8 + modifier*multiplier*8, -1023, 1023) / 1023 with an int8 base.  Our own search: PSNR parity.
// xs channel `chan` holds v*255 with v clamped to [0,1] (unsigned) or [-1,1] (signed).
template <bool SIGNED>
CFX_HD uint2 encode_eac_r11(float* xs, uint32_t lane, uint32_t chan, int radius, uint32_t vm = 0xFFFFu)
{
    const float scale = SIGNED ? 1023.0f/255.0f : 2047.0f/255.0f;
    const int off = SIGNED ? 0 : 4, vmin = SIGNED ? -1023 : 0, vmax = SIGNED ? 1023 : 2047;
    const int bmin = SIGNED ? -127 : 0, bmax = SIGNED ? 127 : 255;
    float lo = 3.0e38f, hi = -3.0e38f;
    for (uint32_t t = 0; t < 16; ++t) { if (!((vm >> t) & 1u)) continue; const float a = px(xs, lane, t, chan)*scale; lo = fminf(lo, a); hi = fmaxf(hi, a); }
    int best_base = 0, best_mul = 1;
    uint32_t best_tab = 13;
    float best = 3.0e38f;
#pragma unroll 1
    for (uint32_t tab = 0; tab < 16; ++tab) {
        const float tmin = static_cast<float>(kEac[tab][3]), tmax = static_cast<float>(kEac[tab][7]);
        const float range = (tmax - tmin)*8.0f;
        const int mul0 = min(max(__float2int_rn((hi - lo)/range), 1), 15);
#pragma unroll 1
        for (int dm = -1; dm <= 1; ++dm) {
            const int mul = mul0 + dm;
            if (mul < 1 || mul > 15) continue;
            const int base0 = __float2int_rn((0.5f*(lo + hi) - static_cast<float>(off) - 4.0f*(tmin + tmax)*static_cast<float>(mul))*0.125f);
#pragma unroll 1
            for (int db = -radius; db <= radius; ++db) {
                const int base = base0 + db;
                if (base < bmin || base > bmax) continue;
                float err = 0.0f;
                for (uint32_t t = 0; t < 16 && err < best; ++t) {
                    if (!((vm >> t) & 1u)) continue;
                    const float a = px(xs, lane, t, chan)*scale;
                    float be = 3.0e38f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float d = static_cast<float>(min(max(base*8 + off + kEac[tab][k]*mul*8, vmin), vmax)) - a;
                        be = fminf(be, d*d);
                    }
                    err += be;
                }
                if (err < best) { best = err; best_base = base; best_mul = mul; best_tab = tab; }
            }
        }
    }
    uint64_t bits = 0;
    for (uint32_t p = 0; p < 16; ++p) {
        const uint32_t t = (p & 3u)*4u + (p >> 2);
        const float a = px(xs, lane, t, chan)*scale;
        float be = 3.0e38f;
        uint32_t bk = 0;
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) {
            const float d = static_cast<float>(min(max(best_base*8 + off + kEac[best_tab][k]*best_mul*8, vmin), vmax)) - a;
            if (d*d < be) { be = d*d; bk = k; }
        }
        bits = (bits << 3) | bk;
    }
    const uint32_t hi32 = ((static_cast<uint32_t>(best_base) & 0xFFu) << 24) | (static_cast<uint32_t>(best_mul) << 20) | (best_tab << 16) |
        static_cast<uint32_t>(bits >> 32);
    return to_bytes(hi32, static_cast<uint32_t>(bits));
}

// format: 37 ETC1, 38 ETC2 RGB, 40 ETC2 RGBA8 (colour part); returns the 8 colour bytes
template <bool PERC = false>
CFX_HD uint2 encode_color(float* xs, uint32_t lane, bool etc2, int rounds, uint32_t vm = 0xFFFFu)
{
    ColorResult best;
    encode_etc1<PERC>(xs, lane, rounds, best, false, 0, vm);
    if (etc2 && best.err > 0.0f) {
        ColorResult r;
        encode_planar<PERC>(xs, lane, rounds, r, vm);
        if (r.err < best.err) best = r;
        if (best.err > 0.0f) {
            encode_th<PERC>(xs, lane, r, rounds >= (CFX_ETC_TH_AT_NORMAL ? 1 : 2) ? rounds : 0, rounds >= 2 ? 3.0e38f : best.err*CFX_ETC_TH_GATE, vm);
            if (r.err < best.err) best = r;
        }
    }
    return to_bytes(best.hi, best.lo);
}

} // namespace etc
} // namespace cfx
