// Byte-exact BC1 colour block for Quality::Normal: a restatement of what
// rgbcx::encode_bc1(level 9, ...) computes (lib/bc7enc_rdo/rgbcx.cpp:1692-1811 level table,
// :2263-2604 driver, :1813-2175 initial end points, :1000-1045 selector search, :727-773 least squares,
// :1194-1228 565 rounding, :1348-1690 three-colour trials, :1245-1338 packing, :629-691 solid blocks),
// as Bc1Converter / Bc1AConverter / Bc2Converter / Bc3Converter::compressBlock call it
// (lib/src/S3tcConverter.cpp:263-376).  ONE LANE OWNS ONE BLOCK.
//
// Byte parity depends on doing the same IEEE single-precision operations in the same order with no
// fused multiply-add (the reference binary has none): bc1_bc3.cu is compiled with -fmad=false, casts
// are truncating, divisions are IEEE.  The data tables (single-colour matches, rounding midpoints,
// total-ordering hash / factors and rgbcx's TRAINED "likely total orderings") are not ours and are
// not in this repository: tools/gen_rgbcx_tables.py dumps them from the reference's rgbcx.cpp at
// build time into csrc/generated/rgbcx_tables.inc (git-ignored).  Without that file this header is
// not compiled and the library uses its own search (bc1_core.cuh, PSNR parity) for every quality.
#pragma once
#include "hostdev.h"

namespace cfx {
namespace rgbcx9 {

#include "generated/rgbcx_tables.inc"

CFX_HD float tabf(const uint32_t* t, uint32_t i) { return __uint_as_float(t[i]); }
CFX_HD int to_5(uint32_t v) { v = v*31u + 128u; return static_cast<int>(((v + (v >> 8)) >> 8) & 0xFFu); }
CFX_HD int to_6(uint32_t v) { v = v*63u + 128u; return static_cast<int>(((v + (v >> 8)) >> 8) & 0xFFu); }
CFX_HD int sq(int v) { return v*v; }

struct Block { int r[16], g[16], b[16]; };
struct Ends { int lr, lg, lb, hr, hg, hb; };

CFX_HD bool same(const Ends& a, const Ends& b)
{
    return a.lr == b.lr && a.lg == b.lg && a.lb == b.lb && a.hr == b.hr && a.hg == b.hg && a.hb == b.hb;
}

CFX_HD void colors4(const Ends& e, int* br, int* bg, int* bb)
{
    br[0] = (e.lr << 3) | (e.lr >> 2); bg[0] = (e.lg << 2) | (e.lg >> 4); bb[0] = (e.lb << 3) | (e.lb >> 2);
    br[3] = (e.hr << 3) | (e.hr >> 2); bg[3] = (e.hg << 2) | (e.hg >> 4); bb[3] = (e.hb << 3) | (e.hb >> 2);
    br[1] = (br[0]*2 + br[3])/3; bg[1] = (bg[0]*2 + bg[3])/3; bb[1] = (bb[0]*2 + bb[3])/3;
    br[2] = (br[3]*2 + br[0])/3; bg[2] = (bg[3]*2 + bg[0])/3; bb[2] = (bb[3]*2 + bb[0])/3;
}

CFX_HD void colors3(const Ends& e, int* br, int* bg, int* bb)
{
    br[0] = (e.lr << 3) | (e.lr >> 2); bg[0] = (e.lg << 2) | (e.lg >> 4); bb[0] = (e.lb << 3) | (e.lb >> 2);
    br[1] = (e.hr << 3) | (e.hr >> 2); bg[1] = (e.hg << 2) | (e.hg >> 4); bb[1] = (e.hb << 3) | (e.hb >> 2);
    br[2] = (br[0] + br[1])/2; bg[2] = (bg[0] + bg[1])/2; bb[2] = (bb[0] + bb[1])/2;
}

// bc1_find_sels4_check2_err: sels = 2 bits per texel, linear order (0 = low ... 3 = high)
CFX_HD uint32_t find_sels4(const Block& p, const Ends& e, uint32_t& sels, uint32_t cur_err)
{
    int br[4], bg[4], bb[4];
    colors4(e, br, bg, bb);
    const int dr = br[3] - br[0], dg = bg[3] - bg[0], db = bb[3] - bb[0];
    const float f = 4.0f/(static_cast<float>(sq(dr) + sq(dg) + sq(db)) + .00000125f);
    uint32_t total = 0, out = sels;
    for (int i = 0; i < 16; ++i) {
        const int r = p.r[i], g = p.g[i], b = p.b[i];
        int sel = static_cast<int>(static_cast<float>((r - br[0])*dr + (g - bg[0])*dg + (b - bb[0])*db)*f + .5f);
        sel = min(max(sel, 1), 3);
        const int lo = sel - 1;
        const int r0 = lo == 0 ? br[0] : (lo == 1 ? br[1] : br[2]), g0 = lo == 0 ? bg[0] : (lo == 1 ? bg[1] : bg[2]),
            b0 = lo == 0 ? bb[0] : (lo == 1 ? bb[1] : bb[2]);
        const int r1 = sel == 1 ? br[1] : (sel == 2 ? br[2] : br[3]), g1 = sel == 1 ? bg[1] : (sel == 2 ? bg[2] : bg[3]),
            b1 = sel == 1 ? bb[1] : (sel == 2 ? bb[2] : bb[3]);
        const uint32_t err0 = static_cast<uint32_t>(sq(r0 - r) + sq(g0 - g) + sq(b0 - b));
        const uint32_t err1 = static_cast<uint32_t>(sq(r1 - r) + sq(g1 - g) + sq(b1 - b));
        int best_sel = sel;
        uint32_t best_err = err1;
        if (err0 == err1) { if (best_sel - 1 == 0) best_sel = 0; }
        else if (err0 < best_err) { best_sel = sel - 1; best_err = err0; }
        total += best_err;
        if (total >= cur_err) break;
        out = (out & ~(3u << (2*i))) | (static_cast<uint32_t>(best_sel) << (2*i));
    }
    sels = out;
    return total;
}

// bc1_find_sels4_noerr (rgbcx.cpp:915-950): selectors by projection on the end point axis, no error
CFX_HD void find_sels4_noerr(const Block& p, const Ends& e, uint32_t& sels)
{
    int br[4], bg[4], bb[4];
    colors4(e, br, bg, bb);
    int ar = br[3] - br[0], ag = bg[3] - bg[0], ab = bb[3] - bb[0];
    int dots[4];
    for (int i = 0; i < 4; ++i) dots[i] = br[i]*ar + bg[i]*ag + bb[i]*ab;
    const int t0 = dots[0] + dots[1], t1 = dots[1] + dots[2], t2 = dots[2] + dots[3];
    ar *= 2; ag *= 2; ab *= 2;
    uint32_t out = 0;
    for (int i = 0; i < 16; ++i) {
        const int d = p.r[i]*ar + p.g[i]*ag + p.b[i]*ab;
        out |= static_cast<uint32_t>(3 - ((d <= t0 ? 1 : 0) + (d < t1 ? 1 : 0) + (d < t2 ? 1 : 0))) << (2*i);
    }
    sels = out;
}

// bc1_find_sels4_fullerr (rgbcx.cpp:1047-1081): every texel against all four colours, ties go to colour 3
CFX_HD uint32_t find_sels4_full(const Block& p, const Ends& e, uint32_t& sels, uint32_t cur_err)
{
    int br[4], bg[4], bb[4];
    colors4(e, br, bg, bb);
    uint32_t total = 0, out = sels;
    for (int i = 0; i < 16; ++i) {
        const int r = p.r[i], g = p.g[i], b = p.b[i];
        uint32_t best_err = static_cast<uint32_t>(sq(br[0] - r) + sq(bg[0] - g) + sq(bb[0] - b)), best_sel = 0;
        for (uint32_t j = 1; j < 4 && best_err; ++j) {
            const uint32_t err = static_cast<uint32_t>(sq(br[j] - r) + sq(bg[j] - g) + sq(bb[j] - b));
            if (err < best_err || (err == best_err && j == 3)) { best_err = err; best_sel = j; }
        }
        total += best_err;
        if (total >= cur_err) break;
        out = (out & ~(3u << (2*i))) | (best_sel << (2*i));
    }
    sels = out;
    return total;
}

// bc1_find_sels3_fullerr
CFX_HD uint32_t find_sels3(bool use_black, const Block& p, const Ends& e, uint32_t& sels, uint32_t cur_err)
{
    int br[3], bg[3], bb[3];
    colors3(e, br, bg, bb);
    uint32_t total = 0, out = sels;
    for (int i = 0; i < 16; ++i) {
        const int r = p.r[i], g = p.g[i], b = p.b[i];
        uint32_t best_err = static_cast<uint32_t>(sq(br[0] - r) + sq(bg[0] - g) + sq(bb[0] - b)), best_sel = 0;
        const uint32_t err1 = static_cast<uint32_t>(sq(br[1] - r) + sq(bg[1] - g) + sq(bb[1] - b));
        if (err1 < best_err) { best_err = err1; best_sel = 1; }
        const uint32_t err2 = static_cast<uint32_t>(sq(br[2] - r) + sq(bg[2] - g) + sq(bb[2] - b));
        if (err2 < best_err) { best_err = err2; best_sel = 2; }
        if (use_black) {
            const uint32_t err3 = static_cast<uint32_t>(sq(r) + sq(g) + sq(b));
            if (err3 < best_err) { best_err = err3; best_sel = 3; }
        }
        total += best_err;
        if (total >= cur_err) { sels = out; return total; }
        out = (out & ~(3u << (2*i))) | (best_sel << (2*i));
    }
    sels = out;
    return total;
}

// compute_least_squares_endpoints4_rgb (selector version)
CFX_HD bool ls4(const Block& p, uint32_t sels, int total_r, int total_g, int total_b, float* xl, float* xh)
{
    uint32_t uq_r = 0, uq_g = 0, uq_b = 0, wacc = 0;
    for (int i = 0; i < 16; ++i) {
        const uint32_t sel = (sels >> (2*i)) & 3u;
        wacc += sel == 0 ? 0x000009u : (sel == 1 ? 0x010204u : (sel == 2 ? 0x040201u : 0x090000u));
        uq_r += sel*static_cast<uint32_t>(p.r[i]); uq_g += sel*static_cast<uint32_t>(p.g[i]); uq_b += sel*static_cast<uint32_t>(p.b[i]);
    }
    const int q_r = total_r*3 - static_cast<int>(uq_r), q_g = total_g*3 - static_cast<int>(uq_g), q_b = total_b*3 - static_cast<int>(uq_b);
    const float z00 = static_cast<float>((wacc >> 16) & 0xFF), z10 = static_cast<float>((wacc >> 8) & 0xFF), z11 = static_cast<float>(wacc & 0xFF);
    const float z01 = z10;
    float det = z00*z11 - z01*z10;
    if (fabsf(det) < 1e-8f) return false;
    det = (3.0f/255.0f)/det;
    const float iz00 = z11*det, iz01 = -z01*det, iz10 = -z10*det, iz11 = z00*det;
    xl[0] = iz00*static_cast<float>(uq_r) + iz01*static_cast<float>(q_r); xh[0] = iz10*static_cast<float>(uq_r) + iz11*static_cast<float>(q_r);
    xl[1] = iz00*static_cast<float>(uq_g) + iz01*static_cast<float>(q_g); xh[1] = iz10*static_cast<float>(uq_g) + iz11*static_cast<float>(q_g);
    xl[2] = iz00*static_cast<float>(uq_b) + iz01*static_cast<float>(q_b); xh[2] = iz10*static_cast<float>(uq_b) + iz11*static_cast<float>(q_b);
    return true;
}

// compute_least_squares_endpoints3_rgb (selector version)
CFX_HD bool ls3(bool use_black, const Block& p, uint32_t sels, float* xl, float* xh)
{
    int uq_r = 0, uq_g = 0, uq_b = 0, total_r = 0, total_g = 0, total_b = 0;
    uint32_t wacc = 0;
    for (int i = 0; i < 16; ++i) {
        const int r = p.r[i], g = p.g[i], b = p.b[i];
        if (use_black && ((r | g | b) < 4)) continue;
        const uint32_t sel = (sels >> (2*i)) & 3u;
        if (sel == 3) continue;
        wacc += sel == 0 ? 0x000004u : (sel == 1 ? 0x040000u : 0x010101u);
        const int tsel = sel == 0 ? 0 : (sel == 1 ? 2 : 1);
        uq_r += tsel*r; uq_g += tsel*g; uq_b += tsel*b;
        total_r += r; total_g += g; total_b += b;
    }
    const int q_r = total_r*2 - uq_r, q_g = total_g*2 - uq_g, q_b = total_b*2 - uq_b;
    const float z00 = static_cast<float>((wacc >> 16) & 0xFF), z10 = static_cast<float>((wacc >> 8) & 0xFF), z11 = static_cast<float>(wacc & 0xFF);
    const float z01 = z10;
    float det = z00*z11 - z01*z10;
    if (fabsf(det) < 1e-8f) return false;
    det = (2.0f/255.0f)/det;
    const float iz00 = z11*det, iz01 = -z01*det, iz10 = -z10*det, iz11 = z00*det;
    xl[0] = iz00*static_cast<float>(uq_r) + iz01*static_cast<float>(q_r); xh[0] = iz10*static_cast<float>(uq_r) + iz11*static_cast<float>(q_r);
    xl[1] = iz00*static_cast<float>(uq_g) + iz01*static_cast<float>(q_g); xh[1] = iz10*static_cast<float>(uq_g) + iz11*static_cast<float>(q_g);
    xl[2] = iz00*static_cast<float>(uq_b) + iz01*static_cast<float>(q_b); xh[2] = iz10*static_cast<float>(uq_b) + iz11*static_cast<float>(q_b);
    return true;
}

// precise_round_565(xl, xh, l.., h..): first triple from xl, second from xh
CFX_HD void round565(const float* xl, const float* xh, int& ar, int& ag, int& ab, int& cr, int& cg, int& cb)
{
    ar = static_cast<int>(xl[0]*31.0f); ag = static_cast<int>(xl[1]*63.0f); ab = static_cast<int>(xl[2]*31.0f);
    cr = static_cast<int>(xh[0]*31.0f); cg = static_cast<int>(xh[1]*63.0f); cb = static_cast<int>(xh[2]*31.0f);
    if (static_cast<uint32_t>(ar | ab | cr | cb) > 31u) {
        ar = (static_cast<uint32_t>(ar) > 31u) ? (~ar >> 31) & 31 : ar;
        cr = (static_cast<uint32_t>(cr) > 31u) ? (~cr >> 31) & 31 : cr;
        ab = (static_cast<uint32_t>(ab) > 31u) ? (~ab >> 31) & 31 : ab;
        cb = (static_cast<uint32_t>(cb) > 31u) ? (~cb >> 31) & 31 : cb;
    }
    if (static_cast<uint32_t>(ag | cg) > 63u) {
        ag = (static_cast<uint32_t>(ag) > 63u) ? (~ag >> 31) & 63 : ag;
        cg = (static_cast<uint32_t>(cg) > 63u) ? (~cg >> 31) & 63 : cg;
    }
    ar = (ar + (xl[0] > tabf(kMidpoint5, ar) ? 1 : 0)) & 31;
    ag = (ag + (xl[1] > tabf(kMidpoint6, ag) ? 1 : 0)) & 63;
    ab = (ab + (xl[2] > tabf(kMidpoint5, ab) ? 1 : 0)) & 31;
    cr = (cr + (xh[0] > tabf(kMidpoint5, cr) ? 1 : 0)) & 31;
    cg = (cg + (xh[1] > tabf(kMidpoint6, cg) ? 1 : 0)) & 63;
    cb = (cb + (xh[2] > tabf(kMidpoint5, cb) ? 1 : 0)) & 31;
}

// The call sites pass (xl, xh, trial_h*, trial_l*): the HIGH end point is rounded from xl.
CFX_HD void round_to_ends(const float* xl, const float* xh, Ends& e)
{
    round565(xl, xh, e.hr, e.hg, e.hb, e.lr, e.lg, e.lb);
}

CFX_HD void match_eq1(int avg_r, int avg_g, int avg_b, Ends& e)
{
    e.lr = kMatch5Eq1[avg_r] & 0xFF; e.lg = kMatch6Eq1[avg_g] & 0xFF; e.lb = kMatch5Eq1[avg_b] & 0xFF;
    e.hr = (kMatch5Eq1[avg_r] >> 8) & 0xFF; e.hg = (kMatch6Eq1[avg_g] >> 8) & 0xFF; e.hb = (kMatch5Eq1[avg_b] >> 8) & 0xFF;
}

CFX_HD void match_half(int avg_r, int avg_g, int avg_b, Ends& e)
{
    e.lr = kMatch5Half[avg_r] & 0xFF; e.lg = kMatch6Half[avg_g] & 0xFF; e.lb = kMatch5Half[avg_b] & 0xFF;
    e.hr = (kMatch5Half[avg_r] >> 8) & 0xFF; e.hg = (kMatch6Half[avg_g] >> 8) & 0xFF; e.hb = (kMatch5Half[avg_b] >> 8) & 0xFF;
}

// power-iteration axis shared by encode_bc1_pick_initial and try_3color_block_useblack
CFX_HD void principal_axis(const int* icov, int max_r, int min_r, int max_g, int min_g, int max_b, int min_b, float scale,
    int& sr, int& sg, int& sb, int iters = 4)
{
    float xr = static_cast<float>(max_r - min_r), xg = static_cast<float>(max_g - min_g), xb = static_cast<float>(max_b - min_b);
    if (icov[2] < 0) xr = -xr;
    if (icov[4] < 0) xg = -xg;
    float cov[6];
    for (int i = 0; i < 6; ++i) cov[i] = static_cast<float>(icov[i])*(1.0f/255.0f);
    for (int it = 0; it < iters; ++it) {                  // (cEncodeBC1Use6PowerIters: 6)
        const float r = xr*cov[0] + xg*cov[1] + xb*cov[2];
        const float g = xr*cov[1] + xg*cov[3] + xb*cov[4];
        const float b = xr*cov[2] + xg*cov[4] + xb*cov[5];
        xr = r; xg = g; xb = b;
    }
    const float k = fmaxf(fmaxf(fabsf(xr), fabsf(xg)), fabsf(xb));
    sr = 306; sg = 601; sb = 117;
    if (k >= 2) {
        const float m = scale/k;
        sr = static_cast<int>(xr*m); sg = static_cast<int>(xg*m); sb = static_cast<int>(xb*m);
    }
}

CFX_HD void pick_initial(const Block& p, bool grayscale, int min_r, int min_g, int min_b, int max_r, int max_g, int max_b,
    int avg_r, int avg_g, int avg_b, Ends& e, int power_iters = 4, bool bbox_int = false, bool bbox_float = false)
{
    if (grayscale) {
        const int fr = p.r[0];
        if (max_r - min_r < 2) { e.lr = e.lb = e.hr = e.hb = to_5(fr); e.lg = e.hg = to_6(fr); }
        else { e.lr = e.lb = to_5(min_r); e.lg = to_6(min_r); e.hr = e.hb = to_5(max_r); e.hg = to_6(max_r); }
        return;
    }
    if (bbox_float) {
        // cEncodeBC1BoundingBox (rgbcx.cpp:1987-2032; the second initial guess of cEncodeBC1TryAllInitialEndponts)
        float l[3] = {static_cast<float>(min_r)*(1.0f/255.0f), static_cast<float>(min_g)*(1.0f/255.0f), static_cast<float>(min_b)*(1.0f/255.0f)};
        float h[3] = {static_cast<float>(max_r)*(1.0f/255.0f), static_cast<float>(max_g)*(1.0f/255.0f), static_cast<float>(max_b)*(1.0f/255.0f)};
        const float bias = 8.0f/255.0f;
        for (int c = 0; c < 3; ++c) {
            const float inset = (h[c] - l[c] - bias)*(1.0f/16.0f);
            l[c] = fminf(fmaxf(l[c] + inset, 0.0f), 1.0f);
            h[c] = fminf(fmaxf(h[c] - inset, 0.0f), 1.0f);
        }
        int icov_xz = 0, icov_yz = 0;
        for (int i = 0; i < 16; ++i) {
            const int r = p.r[i] - avg_r, g = p.g[i] - avg_g, b = p.b[i] - avg_b;
            icov_xz += r*b; icov_yz += g*b;
        }
        if (icov_xz < 0) { const float t = l[0]; l[0] = h[0]; h[0] = t; }
        if (icov_yz < 0) { const float t = l[1]; l[1] = h[1]; h[1] = t; }
        round565(l, h, e.lr, e.lg, e.lb, e.hr, e.hg, e.hb);
        return;
    }
    if (bbox_int) {
        // cEncodeBC1BoundingBoxInt (rgbcx.cpp:2034-2085): the inset bounding box, its red / green corners swapped by the
        // sign of their covariance with blue
        const int inset_r = (max_r - min_r - 8) >> 4, inset_g = (max_g - min_g - 8) >> 4, inset_b = (max_b - min_b - 8) >> 4;
        min_r += inset_r; min_g += inset_g; min_b += inset_b;
        if (static_cast<uint32_t>(min_r | min_g | min_b) > 255u) { min_r = min(max(min_r, 0), 255); min_g = min(max(min_g, 0), 255); min_b = min(max(min_b, 0), 255); }
        max_r -= inset_r; max_g -= inset_g; max_b -= inset_b;
        if (static_cast<uint32_t>(max_r | max_g | max_b) > 255u) { max_r = min(max(max_r, 0), 255); max_g = min(max(max_g, 0), 255); max_b = min(max(max_b, 0), 255); }
        int icov_xz = 0, icov_yz = 0;
        for (int i = 0; i < 16; ++i) {
            const int r = p.r[i] - avg_r, g = p.g[i] - avg_g, b = p.b[i] - avg_b;
            icov_xz += r*b; icov_yz += g*b;
        }
        int x0 = min_r, y0 = min_g, x1 = max_r, y1 = max_g;
        if (icov_xz < 0) { const int t = x0; x0 = x1; x1 = t; }
        if (icov_yz < 0) { const int t = y0; y0 = y1; y1 = t; }
        e.lr = to_5(static_cast<uint32_t>(x0)); e.lg = to_6(static_cast<uint32_t>(y0)); e.lb = to_5(static_cast<uint32_t>(min_b));
        e.hr = to_5(static_cast<uint32_t>(x1)); e.hg = to_6(static_cast<uint32_t>(y1)); e.hb = to_5(static_cast<uint32_t>(max_b));
        return;
    }
    int icov[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 16; ++i) {
        const int r = p.r[i] - avg_r, g = p.g[i] - avg_g, b = p.b[i] - avg_b;
        icov[0] += r*r; icov[1] += r*g; icov[2] += r*b; icov[3] += g*g; icov[4] += g*b; icov[5] += b*b;
    }
    int sr, sg, sb;
    principal_axis(icov, max_r, min_r, max_g, min_g, max_b, min_b, 2048.0f, sr, sg, sb, power_iters);
    sr = static_cast<int>(static_cast<uint32_t>(sr) << 4); sg = static_cast<int>(static_cast<uint32_t>(sg) << 4);
    sb = static_cast<int>(static_cast<uint32_t>(sb) << 4);
    int low_dot = 2147483647, high_dot = -2147483647 - 1;
    for (int i = 0; i < 16; ++i) {
        const int dot = ((p.r[i]*sr + p.g[i]*sg + p.b[i]*sb) & ~0xF) + i;
        low_dot = min(low_dot, dot); high_dot = max(high_dot, dot);
    }
    const int lc = low_dot & 15, hc = high_dot & 15;
    e.lr = to_5(p.r[lc]); e.lg = to_6(p.g[lc]); e.lb = to_5(p.b[lc]);
    e.hr = to_5(p.r[hc]); e.hg = to_6(p.g[hc]); e.hb = to_5(p.b[hc]);
}

// texels sorted along the end-point axis -> prefix sums of their colours
CFX_HD void prefix_sums(const Block& p, const Ends& e, int total_r, int total_g, int total_b, uint32_t* rs, uint32_t* gs, uint32_t* bs)
{
    const int r0 = (e.lr << 3) | (e.lr >> 2), g0 = (e.lg << 2) | (e.lg >> 4), b0 = (e.lb << 3) | (e.lb >> 2);
    const int r3 = (e.hr << 3) | (e.hr >> 2), g3 = (e.hg << 2) | (e.hg >> 4), b3 = (e.hb << 3) | (e.hb >> 2);
    const int ar = r3 - r0, ag = g3 - g0, ab = b3 - b0;
    int dots[16];
    for (int i = 0; i < 16; ++i) {
        const int d = 0x1000000 + (p.r[i]*ar + p.g[i]*ag + p.b[i]*ab);
        dots[i] = (d << 4) + i;
    }
    // keys are unique (low nibble = texel), so any correct sort gives std::sort's order
    for (int i = 1; i < 16; ++i) {
        const int v = dots[i];
        int j = i - 1;
        while (j >= 0 && dots[j] > v) { dots[j + 1] = dots[j]; --j; }
        dots[j + 1] = v;
    }
    uint32_t r = 0, g = 0, b = 0;
    for (int i = 0; i < 16; ++i) {
        const int q = dots[i] & 15;
        rs[i] = r; gs[i] = g; bs[i] = b;
        r += static_cast<uint32_t>(p.r[q]); g += static_cast<uint32_t>(p.g[q]); b += static_cast<uint32_t>(p.b[q]);
    }
    rs[16] = static_cast<uint32_t>(total_r); gs[16] = static_cast<uint32_t>(total_g); bs[16] = static_cast<uint32_t>(total_b);
}

struct Result { Ends e; uint32_t sels; bool three; };

// try_3color_block (no black)
CFX_HD void try_3color(const Block& p, uint32_t& cur_err, int avg_r, int avg_g, int avg_b, Ends e, int total_r, int total_g,
    int total_b, uint32_t orderings, Result& res)
{
    uint32_t trial_sels = 0;
    uint32_t trial_err = find_sels3(false, p, e, trial_sels, 0xFFFFFFFFu);
    if (trial_err) {
        for (int pass = 0; pass < 2; ++pass) {
            float xl[3], xh[3];
            Ends e2;
            if (!ls3(false, p, trial_sels, xl, xh)) match_half(avg_r, avg_g, avg_b, e2);
            else round_to_ends(xl, xh, e2);
            if (same(e, e2)) break;
            uint32_t sels2 = 0;
            const uint32_t err2 = find_sels3(false, p, e2, sels2, trial_err);
            if (err2 < trial_err) { trial_err = err2; e = e2; trial_sels = sels2; } else break;
        }
    }
    if (trial_err && orderings) {
        uint32_t h[3] = {0, 0, 0};
        for (int i = 0; i < 16; ++i) h[(trial_sels >> (2*i)) & 3u]++;
        uint32_t idx;
        if (h[0] == 16) idx = CFX_RGBCX_TOTAL_ORDER_3_0_16;
        else if (h[1] == 16) idx = CFX_RGBCX_TOTAL_ORDER_3_1_16;
        else if (h[2] == 16) idx = CFX_RGBCX_TOTAL_ORDER_3_2_16;
        else idx = kOrderHash3[h[0] | (h[1] << 4)];
        uint32_t rs[17], gs[17], bs[17];
        prefix_sums(p, e, total_r, total_g, total_b, rs, gs, bs);
        for (uint32_t q = 0; q < orderings; ++q) {
            const uint32_t s = kBestOrders3[idx*32u + q];
            Ends t;
            if (s == CFX_RGBCX_TOTAL_ORDER_3_0_16 || s == CFX_RGBCX_TOTAL_ORDER_3_1_16 || s == CFX_RGBCX_TOTAL_ORDER_3_2_16) {
                match_half(avg_r, avg_g, avg_b, t);
            } else {
                const float iz00 = tabf(kSelFactors3, s*3u), iz10 = tabf(kSelFactors3, s*3u + 1u), iz11 = tabf(kSelFactors3, s*3u + 2u);
                const float iz01 = iz10;
                const uint32_t f1 = kUniqueOrders3[s*3u], f2 = kUniqueOrders3[s*3u] + kUniqueOrders3[s*3u + 2u];
                const uint32_t uq_r = (rs[16] - rs[f2])*2u + (rs[f2] - rs[f1]);
                const uint32_t uq_g = (gs[16] - gs[f2])*2u + (gs[f2] - gs[f1]);
                const uint32_t uq_b = (bs[16] - bs[f2])*2u + (bs[f2] - bs[f1]);
                const float q_r = static_cast<float>(static_cast<uint32_t>(total_r*2) - uq_r), q_g = static_cast<float>(static_cast<uint32_t>(total_g*2) - uq_g),
                    q_b = static_cast<float>(static_cast<uint32_t>(total_b*2) - uq_b);
                float xl[3], xh[3];
                xl[0] = iz00*static_cast<float>(uq_r) + iz01*q_r; xh[0] = iz10*static_cast<float>(uq_r) + iz11*q_r;
                xl[1] = iz00*static_cast<float>(uq_g) + iz01*q_g; xh[1] = iz10*static_cast<float>(uq_g) + iz11*q_g;
                xl[2] = iz00*static_cast<float>(uq_b) + iz01*q_b; xh[2] = iz10*static_cast<float>(uq_b) + iz11*q_b;
                round_to_ends(xl, xh, t);
            }
            uint32_t sels2 = 0;
            const uint32_t err2 = find_sels3(false, p, t, sels2, 0xFFFFFFFFu);
            if (err2 < trial_err) { trial_err = err2; e = t; trial_sels = sels2; }
        }
    }
    if (trial_err < cur_err) { res.three = true; res.e = e; res.sels = trial_sels; cur_err = trial_err; }
}

// try_3color_block_useblack
CFX_HD void try_3color_black(const Block& p, uint32_t& cur_err, Result& res)
{
    int total_r = 0, total_g = 0, total_b = 0, max_r = 0, max_g = 0, max_b = 0, min_r = 255, min_g = 255, min_b = 255, n = 0;
    for (int i = 0; i < 16; ++i) {
        const int r = p.r[i], g = p.g[i], b = p.b[i];
        if ((r | g | b) < 4) continue;
        max_r = max(max_r, r); max_g = max(max_g, g); max_b = max(max_b, b);
        min_r = min(min_r, r); min_g = min(min_g, g); min_b = min(min_b, b);
        total_r += r; total_g += g; total_b += b; ++n;
    }
    if (!n) return;
    const int half = n >> 1;
    const int avg_r = (total_r + half)/n, avg_g = (total_g + half)/n, avg_b = (total_b + half)/n;
    int icov[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 16; ++i) {
        int r = p.r[i], g = p.g[i], b = p.b[i];
        if ((r | g | b) < 4) continue;
        r -= avg_r; g -= avg_g; b -= avg_b;
        icov[0] += r*r; icov[1] += r*g; icov[2] += r*b; icov[3] += g*g; icov[4] += g*b; icov[5] += b*b;
    }
    int sr, sg, sb;
    principal_axis(icov, max_r, min_r, max_g, min_g, max_b, min_b, 1024.0f, sr, sg, sb);
    int low_dot = 2147483647, high_dot = -2147483647 - 1, lc = 0, hc = 0;
    for (int i = 0; i < 16; ++i) {
        const int r = p.r[i], g = p.g[i], b = p.b[i];
        if ((r | g | b) < 4) continue;
        const int dot = r*sr + g*sg + b*sb;
        if (dot < low_dot) { low_dot = dot; lc = i; }
        if (dot > high_dot) { high_dot = dot; hc = i; }
    }
    Ends e;
    e.lr = to_5(p.r[lc]); e.lg = to_6(p.g[lc]); e.lb = to_5(p.b[lc]);
    e.hr = to_5(p.r[hc]); e.hg = to_6(p.g[hc]); e.hb = to_5(p.b[hc]);
    uint32_t trial_sels = 0;
    uint32_t trial_err = find_sels3(true, p, e, trial_sels, 0xFFFFFFFFu);
    if (trial_err) {
        for (int pass = 0; pass < 2; ++pass) {
            float xl[3], xh[3];
            Ends e2;
            if (!ls3(true, p, trial_sels, xl, xh)) match_half(avg_r, avg_g, avg_b, e2);
            else round_to_ends(xl, xh, e2);
            if (same(e, e2)) break;
            uint32_t sels2 = 0;
            const uint32_t err2 = find_sels3(true, p, e2, sels2, trial_err);
            if (err2 < trial_err) { trial_err = err2; e = e2; trial_sels = sels2; } else break;
        }
    }
    if (trial_err < cur_err) { res.three = true; res.e = e; res.sels = trial_sels; cur_err = trial_err; }
}

CFX_HD uint2 pack_words(uint32_t c0, uint32_t c1, uint32_t sel) { return make_uint2(c0 | (c1 << 16), sel); }

// encode_bc1_solid_block
CFX_HD uint2 solid_block(uint32_t fr, uint32_t fg, uint32_t fb, bool allow_3color)
{
    uint32_t mask = 0xAA;
    int max16 = -1, min16 = 0;
    if (allow_3color) {
        const uint32_t err4 = (kMatch5Eq1[fr] >> 16) + (kMatch6Eq1[fg] >> 16) + (kMatch5Eq1[fb] >> 16);
        const uint32_t err3 = (kMatch5Half[fr] >> 16) + (kMatch6Half[fg] >> 16) + (kMatch5Half[fb] >> 16);
        if (err3 < err4) {
            max16 = static_cast<int>(((kMatch5Half[fr] & 0xFF) << 11) | ((kMatch6Half[fg] & 0xFF) << 5) | (kMatch5Half[fb] & 0xFF));
            min16 = static_cast<int>((((kMatch5Half[fr] >> 8) & 0xFF) << 11) | (((kMatch6Half[fg] >> 8) & 0xFF) << 5) | ((kMatch5Half[fb] >> 8) & 0xFF));
            if (max16 > min16) { const int t = max16; max16 = min16; min16 = t; }
        }
    }
    if (max16 == -1) {
        max16 = static_cast<int>(((kMatch5Eq1[fr] & 0xFF) << 11) | ((kMatch6Eq1[fg] & 0xFF) << 5) | (kMatch5Eq1[fb] & 0xFF));
        min16 = static_cast<int>((((kMatch5Eq1[fr] >> 8) & 0xFF) << 11) | (((kMatch6Eq1[fg] >> 8) & 0xFF) << 5) | ((kMatch5Eq1[fb] >> 8) & 0xFF));
        if (min16 == max16) {
            mask = 0;
            if (min16 > 0) min16--;
            else { max16 = 1; min16 = 0; mask = 0x55; }
        }
        if (max16 < min16) { const int t = max16; max16 = min16; min16 = t; mask ^= 0x55; }
    }
    return pack_words(static_cast<uint32_t>(max16), static_cast<uint32_t>(min16), mask*0x01010101u);
}

// bc1_encode4 / bc1_encode3
CFX_HD uint2 pack_result(const Result& r)
{
    uint32_t lc = static_cast<uint32_t>((r.e.lr << 11) | (r.e.lg << 5) | r.e.lb), hc = static_cast<uint32_t>((r.e.hr << 11) | (r.e.hg << 5) | r.e.hb);
    if (!r.three) {
        if (lc == hc) {
            uint32_t mask = 0;
            if (hc > 0) hc--;
            else { hc = 0; lc = 1; mask = 0x55; }
            return pack_words(lc, hc, mask*0x01010101u);
        }
        uint32_t invert = 0;
        if (lc < hc) { const uint32_t t = lc; lc = hc; hc = t; invert = 0x55555555u; }
        uint32_t packed = 0;
        for (int i = 0; i < 16; ++i) {
            const uint32_t s = (r.sels >> (2*i)) & 3u;
            packed |= (s == 0 ? 0u : (s == 1 ? 2u : (s == 2 ? 3u : 1u))) << (2*i);
        }
        return pack_words(lc, hc, packed ^ invert);
    }
    bool inv = false;
    if (lc > hc) { const uint32_t t = lc; lc = hc; hc = t; inv = true; }
    uint32_t packed = 0;
    for (int i = 0; i < 16; ++i) {
        const uint32_t s = (r.sels >> (2*i)) & 3u;
        packed |= (inv ? (s == 0 ? 1u : (s == 1 ? 0u : s)) : s) << (2*i);
    }
    return pack_words(lc, hc, packed);
}

// rgbcx::encode_bc1(9, dst, px, allow_3color, allow_transparent_texels_for_black)
CFX_HD uint2 encode_bc1_level9(const uint32_t* px, bool allow3, bool allow_black)
{
    Block p;
    for (int i = 0; i < 16; ++i) { p.r[i] = px[i] & 0xFF; p.g[i] = (px[i] >> 8) & 0xFF; p.b[i] = (px[i] >> 16) & 0xFF; }
    const int fr = p.r[0], fg = p.g[0], fb = p.b[0];
    int j;
    for (j = 15; j >= 1; --j) if (p.r[j] != fr || p.g[j] != fg || p.b[j] != fb) break;
    if (j == 0) return solid_block(fr, fg, fb, allow3 || allow_black);

    int total_r = fr, total_g = fg, total_b = fb, max_r = fr, max_g = fg, max_b = fb, min_r = fr, min_g = fg, min_b = fb;
    bool grayscale = fr == fg && fr == fb, any_black = (fr | fg | fb) < 4;
    for (int i = 1; i < 16; ++i) {
        const int r = p.r[i], g = p.g[i], b = p.b[i];
        grayscale = grayscale && r == g && r == b;
        any_black = any_black || ((r | g | b) < 4);
        max_r = max(max_r, r); max_g = max(max_g, g); max_b = max(max_b, b);
        min_r = min(min_r, r); min_g = min(min_g, g); min_b = min(min_b, b);
        total_r += r; total_g += g; total_b += b;
    }
    const int avg_r = (total_r + 8) >> 4, avg_g = (total_g + 8) >> 4, avg_b = (total_b + 8) >> 4;

    Result res; res.three = false; res.sels = 0;
    Ends round_e, orig;
    pick_initial(p, grayscale, min_r, min_g, min_b, max_r, max_g, max_b, avg_r, avg_g, avg_b, round_e);
    orig = round_e;
    uint32_t round_sels = 0;
    uint32_t round_err = find_sels4(p, round_e, round_sels, 0xFFFFFFFFu);
    for (int pass = 0; pass < 2; ++pass) {
        float xl[3], xh[3];
        Ends t;
        if (!ls4(p, round_sels, total_r, total_g, total_b, xl, xh)) {
            match_eq1(avg_r, avg_g, avg_b, t);                                                     // trial_l = m_hi, trial_h = m_lo
        } else round_to_ends(xl, xh, t);
        if (same(round_e, t)) break;
        uint32_t tsels = 0;
        const uint32_t terr = find_sels4(p, t, tsels, round_err);
        if (terr < round_err) { round_e = t; round_err = terr; round_sels = tsels; } else break;
    }
    uint32_t cur_err = round_err;
    res.e = round_e; res.sels = round_sels;

    if (cur_err) {
        uint32_t h[4] = {0, 0, 0, 0};
        for (int i = 0; i < 16; ++i) h[(res.sels >> (2*i)) & 3u]++;
        uint32_t idx;
        if (h[0] == 16) idx = CFX_RGBCX_TOTAL_ORDER_4_0_16;
        else if (h[1] == 16) idx = CFX_RGBCX_TOTAL_ORDER_4_1_16;
        else if (h[2] == 16) idx = CFX_RGBCX_TOTAL_ORDER_4_2_16;
        else if (h[3] == 16) idx = CFX_RGBCX_TOTAL_ORDER_4_3_16;
        else idx = kOrderHash4[h[0] | (h[1] << 4) | (h[2] << 8)];
        uint32_t rs[17], gs[17], bs[17];
        prefix_sums(p, res.e, total_r, total_g, total_b, rs, gs, bs);
        for (uint32_t q = 0; q < 11; ++q) {
            const uint32_t s = kBestOrders4[idx*CFX_RGBCX_ORDERS4_STRIDE + q];
            Ends t;
            if (s == CFX_RGBCX_TOTAL_ORDER_4_0_16 || s == CFX_RGBCX_TOTAL_ORDER_4_1_16 || s == CFX_RGBCX_TOTAL_ORDER_4_2_16 ||
                s == CFX_RGBCX_TOTAL_ORDER_4_3_16) {
                match_eq1(avg_r, avg_g, avg_b, t);
            } else {
                const float iz00 = tabf(kSelFactors4, s*3u), iz10 = tabf(kSelFactors4, s*3u + 1u), iz11 = tabf(kSelFactors4, s*3u + 2u);
                const float iz01 = iz10;
                const uint32_t f1 = kUniqueOrders4[s*4u], f2 = f1 + kUniqueOrders4[s*4u + 1u], f3 = f2 + kUniqueOrders4[s*4u + 2u];
                const uint32_t uq_r = (rs[f2] - rs[f1]) + (rs[f3] - rs[f2])*2u + (rs[16] - rs[f3])*3u;
                const uint32_t uq_g = (gs[f2] - gs[f1]) + (gs[f3] - gs[f2])*2u + (gs[16] - gs[f3])*3u;
                const uint32_t uq_b = (bs[f2] - bs[f1]) + (bs[f3] - bs[f2])*2u + (bs[16] - bs[f3])*3u;
                const float q_r = static_cast<float>(static_cast<uint32_t>(total_r*3) - uq_r), q_g = static_cast<float>(static_cast<uint32_t>(total_g*3) - uq_g),
                    q_b = static_cast<float>(static_cast<uint32_t>(total_b*3) - uq_b);
                float xl[3], xh[3];
                xl[0] = iz00*static_cast<float>(uq_r) + iz01*q_r; xh[0] = iz10*static_cast<float>(uq_r) + iz11*q_r;
                xl[1] = iz00*static_cast<float>(uq_g) + iz01*q_g; xh[1] = iz10*static_cast<float>(uq_g) + iz11*q_g;
                xl[2] = iz00*static_cast<float>(uq_b) + iz01*q_b; xh[2] = iz10*static_cast<float>(uq_b) + iz11*q_b;
                round_to_ends(xl, xh, t);
            }
            uint32_t tsels = 0;
            const uint32_t terr = find_sels4(p, t, tsels, cur_err);
            if (terr < cur_err) { cur_err = terr; res.e = t; res.sels = tsels; }
        }
    }
    if ((allow3 || allow_black) && cur_err) {
        if (allow3) try_3color(p, cur_err, avg_r, avg_g, avg_b, orig, total_r, total_g, total_b, 3, res);
        if (any_black && allow_black) try_3color_black(p, cur_err, res);
    }
    return pack_result(res);
}

// rgbcx::encode_bc1(level 0 / level 4, ...): Quality::Lowest / Quality::Low (S3tcConverter.cpp:70 maps the five levels to
// rgbcx levels 0, 4, 9, 13, 18).  Level 0 = cEncodeBC1BoundingBoxInt: no block error anywhere, selectors by projection,
// one least-squares pass (rgbcx.cpp:2321-2371).  Level 4 = two least-squares passes, full error evaluation, six power
// iterations (the level-9 flow without total orderings and three-colour trials, :2373-2465).  Neither level looks at
// allow_3color / allow_transparent_texels_for_black.
CFX_HD uint2 encode_bc1_level04(const uint32_t* px, bool level4)
{
    Block p;
    for (int i = 0; i < 16; ++i) { p.r[i] = px[i] & 0xFF; p.g[i] = (px[i] >> 8) & 0xFF; p.b[i] = (px[i] >> 16) & 0xFF; }
    const int fr = p.r[0], fg = p.g[0], fb = p.b[0];
    int j;
    for (j = 15; j >= 1; --j) if (p.r[j] != fr || p.g[j] != fg || p.b[j] != fb) break;
    if (j == 0) return solid_block(fr, fg, fb, false);
    int total_r = fr, total_g = fg, total_b = fb, max_r = fr, max_g = fg, max_b = fb, min_r = fr, min_g = fg, min_b = fb;
    bool grayscale = fr == fg && fr == fb;
    for (int i = 1; i < 16; ++i) {
        const int r = p.r[i], g = p.g[i], b = p.b[i];
        grayscale = grayscale && r == g && r == b;
        max_r = max(max_r, r); max_g = max(max_g, g); max_b = max(max_b, b);
        min_r = min(min_r, r); min_g = min(min_g, g); min_b = min(min_b, b);
        total_r += r; total_g += g; total_b += b;
    }
    const int avg_r = (total_r + 8) >> 4, avg_g = (total_g + 8) >> 4, avg_b = (total_b + 8) >> 4;
    Result res; res.three = false; res.sels = 0;
    Ends e;
    pick_initial(p, grayscale, min_r, min_g, min_b, max_r, max_g, max_b, avg_r, avg_g, avg_b, e, level4 ? 6 : 4, !level4);
    uint32_t sels = 0, err = 0xFFFFFFFFu;
    if (level4) err = find_sels4_full(p, e, sels, 0xFFFFFFFFu); else find_sels4_noerr(p, e, sels);
    const int passes = level4 ? 2 : 1;
    for (int pass = 0; pass < passes; ++pass) {
        float xl[3], xh[3];
        Ends t;
        if (!ls4(p, sels, total_r, total_g, total_b, xl, xh)) match_eq1(avg_r, avg_g, avg_b, t);
        else round_to_ends(xl, xh, t);
        if (same(e, t)) break;
        if (level4) {
            uint32_t tsels = 0;
            const uint32_t terr = find_sels4_full(p, t, tsels, err);
            if (terr < err) { e = t; err = terr; sels = tsels; } else break;
        } else {
            find_sels4_noerr(p, t, sels);
            e = t;
        }
    }
    res.e = e; res.sels = sels;
    return pack_result(res);
}

// the 16 neighbouring voxels of an end point the search below visits: dr, dg, db, index of the opposite step (the lattice
// directions of icbc's high quality mode, as rgbcx.cpp:2176-2194 lists them)
CFX_CONST int8_t kAdjVoxels[16][4] = {{1, 0, 0, 3}, {0, 1, 0, 4}, {0, 0, 1, 5}, {-1, 0, 0, 0}, {0, -1, 0, 1}, {0, 0, -1, 2}, {1, 1, 0, 9}, {1, 0, 1, 10},
    {0, 1, 1, 11}, {-1, -1, 0, 6}, {-1, 0, -1, 7}, {0, -1, -1, 8}, {-1, 1, 0, 13}, {1, -1, 0, 12}, {0, -1, 1, 15}, {0, 1, -1, 14}};

// encode_bc1_endpoint_search (rgbcx.cpp:2196-2260): a walk over the 16 neighbouring voxels of each end point, `rounds` trials
CFX_HD void endpoint_search(const Block& p, bool black_sels, int rounds, Result& res, uint32_t cur_err)
{
    int prev_improvement = 0, forbidden = -1;
    for (int i = 0; i < rounds; ++i) {
        if (forbidden == (i & 31)) continue;
        const int dr = kAdjVoxels[i & 15][0], dg = kAdjVoxels[i & 15][1], db = kAdjVoxels[i & 15][2];
        Ends t = res.e;
        if ((i >> 4) & 1) { t.lr = min(max(t.lr + dr, 0), 31); t.lg = min(max(t.lg + dg, 0), 63); t.lb = min(max(t.lb + db, 0), 31); }
        else { t.hr = min(max(t.hr + dr, 0), 31); t.hg = min(max(t.hg + dg, 0), 63); t.hb = min(max(t.hb + db, 0), 31); }
        uint32_t tsels = 0;
        const uint32_t terr = res.three ? find_sels3(black_sels, p, t, tsels, cur_err) : find_sels4_full(p, t, tsels, cur_err);
        if (terr < cur_err) {
            cur_err = terr;
            forbidden = kAdjVoxels[i & 15][3] | (i & 16);
            res.e = t; res.sels = tsels;
            prev_improvement = i;
        }
        if (i - prev_improvement > 32) break;
    }
}

// rgbcx::encode_bc1(level 13 / level 18, ...): Quality::High / Quality::Highest.  Full error evaluation, six power
// iterations, BOTH initial guesses (principal axis, then the float bounding box: cEncodeBC1TryAllInitialEndponts), 32 / 128
// likely total orderings (level 18: iterated once more from the improved selectors), 32 three-colour orderings, and the
// end point voxel walk with 20 / 256 trials (rgbcx.cpp:1765-1769, :1795-1799, :2373-2603).
CFX_HD uint2 encode_bc1_level1318(const uint32_t* px, bool level18, bool allow3, bool allow_black)
{
    Block p;
    for (int i = 0; i < 16; ++i) { p.r[i] = px[i] & 0xFF; p.g[i] = (px[i] >> 8) & 0xFF; p.b[i] = (px[i] >> 16) & 0xFF; }
    const int fr = p.r[0], fg = p.g[0], fb = p.b[0];
    int j;
    for (j = 15; j >= 1; --j) if (p.r[j] != fr || p.g[j] != fg || p.b[j] != fb) break;
    if (j == 0) return solid_block(fr, fg, fb, allow3 || allow_black);
    int total_r = fr, total_g = fg, total_b = fb, max_r = fr, max_g = fg, max_b = fb, min_r = fr, min_g = fg, min_b = fb;
    bool grayscale = fr == fg && fr == fb, any_black = (fr | fg | fb) < 4;
    for (int i = 1; i < 16; ++i) {
        const int r = p.r[i], g = p.g[i], b = p.b[i];
        grayscale = grayscale && r == g && r == b;
        any_black = any_black || ((r | g | b) < 4);
        max_r = max(max_r, r); max_g = max(max_g, g); max_b = max(max_b, b);
        min_r = min(min_r, r); min_g = min(min_g, g); min_b = min(min_b, b);
        total_r += r; total_g += g; total_b += b;
    }
    const int avg_r = (total_r + 8) >> 4, avg_g = (total_g + 8) >> 4, avg_b = (total_b + 8) >> 4;

    Result res; res.three = false; res.sels = 0; res.e.lr = res.e.lg = res.e.lb = res.e.hr = res.e.hg = res.e.hb = 0;
    Ends orig = res.e;
    uint32_t cur_err = 0xFFFFFFFFu;
    for (int round = 0; round < 2; ++round) {
        Ends round_e;
        pick_initial(p, grayscale, min_r, min_g, min_b, max_r, max_g, max_b, avg_r, avg_g, avg_b, round_e, 6, false, round == 1);
        const Ends orig_round = round_e;
        uint32_t round_sels = 0;
        uint32_t round_err = find_sels4_full(p, round_e, round_sels, 0xFFFFFFFFu);
        for (int pass = 0; pass < 2; ++pass) {
            float xl[3], xh[3];
            Ends t;
            if (!ls4(p, round_sels, total_r, total_g, total_b, xl, xh)) match_eq1(avg_r, avg_g, avg_b, t);
            else round_to_ends(xl, xh, t);
            if (same(round_e, t)) break;
            uint32_t tsels = 0;
            const uint32_t terr = find_sels4_full(p, t, tsels, round_err);
            if (terr < round_err) { round_e = t; round_err = terr; round_sels = tsels; } else break;
        }
        if (round_err <= cur_err) { cur_err = round_err; res.e = round_e; orig = orig_round; res.sels = round_sels; }
    }
    if (cur_err) {
        const uint32_t q_total = level18 ? 128u : 32u;
        for (int iter = 0; iter < (level18 ? 2 : 1); ++iter) {
            const uint32_t orig_err = cur_err;
            uint32_t h[4] = {0, 0, 0, 0};
            for (int i = 0; i < 16; ++i) h[(res.sels >> (2*i)) & 3u]++;
            uint32_t idx;
            if (h[0] == 16) idx = CFX_RGBCX_TOTAL_ORDER_4_0_16;
            else if (h[1] == 16) idx = CFX_RGBCX_TOTAL_ORDER_4_1_16;
            else if (h[2] == 16) idx = CFX_RGBCX_TOTAL_ORDER_4_2_16;
            else if (h[3] == 16) idx = CFX_RGBCX_TOTAL_ORDER_4_3_16;
            else idx = kOrderHash4[h[0] | (h[1] << 4) | (h[2] << 8)];
            uint32_t rs[17], gs[17], bs[17];
            prefix_sums(p, res.e, total_r, total_g, total_b, rs, gs, bs);
            for (uint32_t q = 0; q < q_total; ++q) {
                const uint32_t s = kBestOrders4[idx*CFX_RGBCX_ORDERS4_STRIDE + q];
                Ends t;
                if (s == CFX_RGBCX_TOTAL_ORDER_4_0_16 || s == CFX_RGBCX_TOTAL_ORDER_4_1_16 || s == CFX_RGBCX_TOTAL_ORDER_4_2_16 ||
                    s == CFX_RGBCX_TOTAL_ORDER_4_3_16) {
                    match_eq1(avg_r, avg_g, avg_b, t);
                } else {
                    const float iz00 = tabf(kSelFactors4, s*3u), iz10 = tabf(kSelFactors4, s*3u + 1u), iz11 = tabf(kSelFactors4, s*3u + 2u);
                    const float iz01 = iz10;
                    const uint32_t f1 = kUniqueOrders4[s*4u], f2 = f1 + kUniqueOrders4[s*4u + 1u], f3 = f2 + kUniqueOrders4[s*4u + 2u];
                    const uint32_t uq_r = (rs[f2] - rs[f1]) + (rs[f3] - rs[f2])*2u + (rs[16] - rs[f3])*3u;
                    const uint32_t uq_g = (gs[f2] - gs[f1]) + (gs[f3] - gs[f2])*2u + (gs[16] - gs[f3])*3u;
                    const uint32_t uq_b = (bs[f2] - bs[f1]) + (bs[f3] - bs[f2])*2u + (bs[16] - bs[f3])*3u;
                    const float q_r = static_cast<float>(static_cast<uint32_t>(total_r*3) - uq_r), q_g = static_cast<float>(static_cast<uint32_t>(total_g*3) - uq_g),
                        q_b = static_cast<float>(static_cast<uint32_t>(total_b*3) - uq_b);
                    float xl[3], xh[3];
                    xl[0] = iz00*static_cast<float>(uq_r) + iz01*q_r; xh[0] = iz10*static_cast<float>(uq_r) + iz11*q_r;
                    xl[1] = iz00*static_cast<float>(uq_g) + iz01*q_g; xh[1] = iz10*static_cast<float>(uq_g) + iz11*q_g;
                    xl[2] = iz00*static_cast<float>(uq_b) + iz01*q_b; xh[2] = iz10*static_cast<float>(uq_b) + iz11*q_b;
                    round_to_ends(xl, xh, t);
                }
                uint32_t tsels = 0;
                const uint32_t terr = find_sels4_full(p, t, tsels, cur_err);
                if (terr < cur_err) { cur_err = terr; res.e = t; res.sels = tsels; }
            }
            if (!cur_err || cur_err == orig_err) break;
        }
    }
    if ((allow3 || allow_black) && cur_err) {
        if (allow3) try_3color(p, cur_err, avg_r, avg_g, avg_b, orig, total_r, total_g, total_b, 32, res);
        if (any_black && allow_black) try_3color_black(p, cur_err, res);
    }
    if (cur_err) endpoint_search(p, any_black && allow_black, level18 ? 256 : 20, res, cur_err);
    return pack_result(res);
}

// the exact restatement for a Texture::Quality (0 = Lowest ... 4 = Highest)
CFX_HD uint2 encode_bc1_exact(const uint32_t* px, uint32_t quality, bool allow3, bool allow_black)
{
    if (quality >= 3u) return encode_bc1_level1318(px, quality >= 4u, allow3, allow_black);
    return quality == 2u ? encode_bc1_level9(px, allow3, allow_black) : encode_bc1_level04(px, quality == 1u);
}

} // namespace rgbcx9
} // namespace cfx
