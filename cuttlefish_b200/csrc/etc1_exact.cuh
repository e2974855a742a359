// Byte-exact ETC1, every Texture::Quality, linear and sRGB colour space (etc2comp's RGBX / REC709 error metric): a restatement of what etc2comp computes when
// EtcConverter::process (lib/src/EtcConverter.cpp:120-152) encodes one block as its own Etc::Image.
//   effort <= 40 (Lowest / Low / Normal): only encoding iteration 0 runs (lib/etc2comp/EtcLib/Etc/EtcImage.cpp:276-330 --
//   with one block per image the effort percentage rounds to zero blocks; Block4x4Encoding_ETC1::PerformFirstIteration,
//   EtcCodec/EtcBlock4x4Encoding_ETC1.cpp:311-338):
//     source averages of the four halves (:402-410), most likely flip from "gray line" distances
//     (:350-386, EtcBlock4x4Encoding_ETC1.h:138-152), then differential and individual tries at radius 0
//     for that flip and for the other one (:545-690, :807-900; base colours from
//     EtcDifferentialTrys.cpp:41-150 / EtcIndividualTrys.cpp:41-64), error metric RGBX
//     (EtcBlock4x4Encoding.cpp:144-154), strict '<' everywhere, SetDoneIfPerfect early exits.
//   effort 70 / 100 (High / Highest): the block is iterated until PerformIteration (:232-306) reports done -- radius-1
//     differential and individual tries (27 base colours per half, the best pair within the differential range) on the
//     likely flip and the other one, then the "degenerate" tries (:1000-1061: differential tries with gray offsets of
//     2 and 4 on either base colour); High stops after the first degenerate set, Highest runs all four.
// Byte parity needs the same IEEE single-precision operations in the same order and no fused
// multiply-add: etc.cu is compiled with -fmad=false.  ONE LANE OWNS ONE BLOCK.
#pragma once
#include "hostdev.h"

namespace cfx {
namespace etc1x {

struct Px { float r, g, b, a; };      // a = NaN marks a border texel (outside the image)

CFX_HD float clamp01(float v) { if (v < 0.0f) v = 0.0f; if (v > 1.0f) v = 1.0f; return v; }

// codeword table entries as the reference's compile-time floats k/255
CFX_CONST float kCwSmall[8] = {2.0f/255.0f, 5.0f/255.0f, 9.0f/255.0f, 13.0f/255.0f, 18.0f/255.0f, 24.0f/255.0f, 33.0f/255.0f, 47.0f/255.0f};
CFX_CONST float kCwLarge[8] = {8.0f/255.0f, 17.0f/255.0f, 29.0f/255.0f, 42.0f/255.0f, 60.0f/255.0f, 80.0f/255.0f, 106.0f/255.0f, 183.0f/255.0f};
CFX_HD float cw_delta(uint32_t cw, uint32_t sel)
{
    const float v = (sel & 1u) ? kCwLarge[cw] : kCwSmall[cw];
    return (sel & 2u) ? -v : v;
}

CFX_HD float gray_distance2(const Px& p, const float* t)
{
    const float dg = ((p.r - t[0]) + (p.g - t[1]) + (p.b - t[2]))/3.0f;
    const float lr = clamp01(t[0] + dg), lg = clamp01(t[1] + dg), lb = clamp01(t[2] + dg);
    const float dr = p.r - lr, dgg = p.g - lg, db = p.b - lb;
    return (dr*dr) + (dgg*dgg) + (db*db);
}

struct HalfTry { int r, g, b; uint32_t cw; uint32_t sel /* 2 bits x 8, pixel order of the half */; float err; };

// TryDifferentialHalf / TryIndividualHalf at radius 0: one base colour, all 8 codewords
// rec709: the REC709 error metric etc2comp uses for sRGB textures (Block4x4Encoding::CalcPixelError,
// EtcBlock4x4Encoding.cpp:157-180: luma weighted 3, blue chroma 0.5; source and decoded alpha are 1 for ETC1)
CFX_HD float pixel_error_rec709(float dr, float dg, float db, const Px& s)
{
    const float luma1 = s.r*0.2126f + s.g*0.7152f + s.b*0.0722f;
    const float cr1 = 0.5f*((s.r - luma1)*(1.0f/(1.0f - 0.2126f)));
    const float cb1 = 0.5f*((s.b - luma1)*(1.0f/(1.0f - 0.0722f)));
    const float luma2 = dr*0.2126f + dg*0.7152f + db*0.0722f;
    const float cr2 = 0.5f*((dr - luma2)*(1.0f/(1.0f - 0.2126f)));
    const float cb2 = 0.5f*((db - luma2)*(1.0f/(1.0f - 0.0722f)));
    const float dl = s.a*luma1 - 1.0f*luma2, dcr = s.a*cr1 - 1.0f*cr2, dcb = s.a*cb1 - 1.0f*cb2, da = 1.0f - s.a;
    return 3.0f*dl*dl + dcr*dcr + 0.5f*dcb*dcb + da*da;
}

// The half's eight texels are read once into registers (this function is the whole cost of the later iterations: 54 calls
// per differential try, 8 tables x 8 texels x 4 selectors each); same operations in the same order as the loops in
// TryDifferentialHalf / TryIndividualHalf.
template <bool REC709>
CFX_HD_NOINLINE void try_half_t(const Px* src, const uint32_t* mapping, float cr, float cg, float cb, HalfTry& t)
{
    Px px[8];
    float a2[8];
    bool border[8];
#pragma unroll
    for (uint32_t i = 0; i < 8; ++i) {
        px[i] = src[mapping[i]];
        border[i] = px[i].a != px[i].a;
        const float da = 1.0f - px[i].a;
        a2[i] = da*da;
    }
    t.err = 3.402823466e+38f; t.cw = 0; t.sel = 0;
#pragma unroll 1
    for (uint32_t cw = 0; cw < 8; ++cw) {
        float sr[4], sg[4], sb[4];
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {
            const float d = cw_delta(cw, k);
            sr[k] = clamp01(cr + d); sg[k] = clamp01(cg + d); sb[k] = clamp01(cb + d);
        }
        uint32_t sels = 0;
        float cw_err = 0.0f;
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) {
            float best = 3.402823466e+38f;
            uint32_t bsel = 0;
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
                float e;
                if (border[i]) e = 0.0f;                        // border texel
                else if (REC709) e = pixel_error_rec709(sr[k], sg[k], sb[k], px[i]);
                else {
                    const float dr = sr[k] - px[i].r, dg = sg[k] - px[i].g, db = sb[k] - px[i].b;
                    e = dr*dr + dg*dg + db*db + a2[i];
                }
                if (e < best) { best = e; bsel = k; }
            }
            sels |= bsel << (2*i);
            cw_err += best;
        }
        if (cw_err < t.err) { t.cw = cw; t.sel = sels; t.err = cw_err; }
    }
}

CFX_HD void try_half(const Px* src, const uint32_t* mapping, float cr, float cg, float cb, HalfTry& t, bool rec709 = false)
{
    if (rec709) try_half_t<true>(src, mapping, cr, cg, cb, t); else try_half_t<false>(src, mapping, cr, cg, cb, t);
}

struct Encoding { bool diff, flip; int r1, g1, b1, r2, g2, b2; uint32_t cw1, cw2, sel1, sel2; float err; };

CFX_HD void bend(int& c1, int& c2)
{
    const int d = c2 - c1;
    if (d > 3) { c1 += (d - 3)/2; c2 = c1 + 3; }
    else if (d < -4) { c1 += (d + 4)/2; c2 = c1 - 4; }
}

CFX_HD int away_from_edge(int v, int radius, int top) { return v < radius ? radius : (v > top - radius ? top - radius : v); }

// TryDifferentialHalf / TryIndividualHalf (EtcBlock4x4Encoding_ETC1.cpp:692-806, :885-998): the (2 radius + 1)^3 base colours
// around (r, g, b), red outermost, each with its best codeword; returns the index of the first best try.
struct TryRec { int r, g, b; uint32_t cw, sel; float err; };
constexpr int kMaxTrys = 27;

CFX_HD_NOINLINE int try_half_radius(const Px* src, const uint32_t* mapping, int r0, int g0, int b0, int radius, bool diff, TryRec* trys, int& count,
    bool rec709)
{
    int best = 0, n = 0;
    float best_err = 3.402823466e+38f;
#pragma unroll 1
    for (int r = r0 - radius; r <= r0 + radius; ++r)
#pragma unroll 1
        for (int g = g0 - radius; g <= g0 + radius; ++g)
#pragma unroll 1
            for (int b = b0 - radius; b <= b0 + radius; ++b) {
                // ConvertFromRGB5 / ConvertFromRGB4 on (unsigned char) components
                const uint32_t ur = static_cast<unsigned char>(r), ug = static_cast<unsigned char>(g), ub = static_cast<unsigned char>(b);
                const float fr = static_cast<float>(static_cast<unsigned char>(diff ? (ur << 3) + (ur >> 2) : (ur << 4) + ur))/255.0f;
                const float fg = static_cast<float>(static_cast<unsigned char>(diff ? (ug << 3) + (ug >> 2) : (ug << 4) + ug))/255.0f;
                const float fb = static_cast<float>(static_cast<unsigned char>(diff ? (ub << 3) + (ub >> 2) : (ub << 4) + ub))/255.0f;
                HalfTry t;
                try_half(src, mapping, fr, fg, fb, t, rec709);
                TryRec& o = trys[n];
                o.r = r; o.g = g; o.b = b; o.cw = t.cw; o.sel = t.sel; o.err = t.err;
                if (t.err < best_err) { best_err = t.err; best = n; }
                ++n;
            }
    count = n;
    return best;
}

struct BlockCtx {
    const Px* src;
    bool rec709;
    float avgL[3], avgR[3], avgT[3], avgB[3];
};
CFX_CONST uint32_t kMapL[8] = {0, 1, 2, 3, 4, 5, 6, 7}, kMapR[8] = {8, 9, 10, 11, 12, 13, 14, 15};
CFX_CONST uint32_t kMapT[8] = {0, 1, 4, 5, 8, 9, 12, 13}, kMapB[8] = {2, 3, 6, 7, 10, 11, 14, 15};

// Block4x4Encoding_ETC1::TryDifferential (:544-690; base colours from DifferentialTrys, EtcDifferentialTrys.cpp:41-150):
// both halves' tries, the best pair whose 5-bit deltas fit [-4, 3]; replaces `best` when strictly better.
// (not inlined: the later iterations call it up to 18 times per block)
CFX_HD_NOINLINE void try_differential(const BlockCtx& bc, bool flip, int radius, int gray1, int gray2, Encoding& best)
{
    const float* c1 = flip ? bc.avgT : bc.avgL;
    const float* c2 = flip ? bc.avgB : bc.avgR;
    const uint32_t* m1 = flip ? kMapT : kMapL;
    const uint32_t* m2 = flip ? kMapB : kMapR;
    int q1[3], q2[3];
    for (int c = 0; c < 3; ++c) {
        // QuantizeR5G5B5 then IntX(31): round(31*clamp(v)) -> expand -> *(1/255) -> round(*31)
        const uint32_t a5 = static_cast<uint32_t>(roundf(31.0f*clamp01(c1[c]))), b5 = static_cast<uint32_t>(roundf(31.0f*clamp01(c2[c])));
        const float fa = (1.0f/255.0f)*static_cast<float>((a5 << 3) + (a5 >> 2)), fb = (1.0f/255.0f)*static_cast<float>((b5 << 3) + (b5 >> 2));
        q1[c] = away_from_edge(static_cast<int>(roundf(fa*31.0f)) + gray1, radius, 31);
        q2[c] = away_from_edge(static_cast<int>(roundf(fb*31.0f)) + gray2, radius, 31);
        bend(q1[c], q2[c]);
    }
    TryRec t1[kMaxTrys], t2[kMaxTrys];
    int n1 = 0, n2 = 0;
    int i1 = try_half_radius(bc.src, m1, q1[0], q1[1], q1[2], radius, true, t1, n1, bc.rec709);
    int i2 = try_half_radius(bc.src, m2, q2[0], q2[1], q2[2], radius, true, t2, n2, bc.rec709);
    float err = 3.402823466e+38f;
    const int dr = t2[i2].r - t1[i1].r, dg = t2[i2].g - t1[i1].g, db = t2[i2].b - t1[i1].b;
    if (dr >= -4 && dr <= 3 && dg >= -4 && dg <= 3 && db >= -4 && db <= 3) err = t1[i1].err + t2[i2].err;
    else {
#pragma unroll 1
        for (int a = 0; a < n1; ++a)
#pragma unroll 1
            for (int b = 0; b < n2; ++b) {
                const int er = t2[b].r - t1[a].r, eg = t2[b].g - t1[a].g, eb = t2[b].b - t1[a].b;
                if (er <= 3 && er >= -4 && eg <= 3 && eg >= -4 && eb <= 3 && eb >= -4) {
                    const float e = t1[a].err + t2[b].err;
                    if (e < err) { err = e; i1 = a; i2 = b; }
                }
            }
    }
    if (err < best.err) {
        best.err = t1[i1].err + t2[i2].err; best.diff = true; best.flip = flip;
        best.r1 = t1[i1].r; best.g1 = t1[i1].g; best.b1 = t1[i1].b; best.r2 = t2[i2].r; best.g2 = t2[i2].g; best.b2 = t2[i2].b;
        best.cw1 = t1[i1].cw; best.cw2 = t2[i2].cw; best.sel1 = t1[i1].sel; best.sel2 = t2[i2].sel;
    }
}

// Block4x4Encoding_ETC1::TryIndividual (:807-883; IndividualTrys, EtcIndividualTrys.cpp:41-64): the best try of each half
CFX_HD_NOINLINE void try_individual(const BlockCtx& bc, bool flip, int radius, Encoding& best)
{
    const float* c1 = flip ? bc.avgT : bc.avgL;
    const float* c2 = flip ? bc.avgB : bc.avgR;
    const uint32_t* m1 = flip ? kMapT : kMapL;
    const uint32_t* m2 = flip ? kMapB : kMapR;
    int q1[3], q2[3];
    for (int c = 0; c < 3; ++c) {
        const uint32_t a4 = static_cast<uint32_t>(roundf(15.0f*clamp01(c1[c]))), b4 = static_cast<uint32_t>(roundf(15.0f*clamp01(c2[c])));
        const float fa = (1.0f/255.0f)*static_cast<float>((a4 << 4) + a4), fb = (1.0f/255.0f)*static_cast<float>((b4 << 4) + b4);
        q1[c] = away_from_edge(static_cast<int>(roundf(fa*15.0f)), radius, 15);
        q2[c] = away_from_edge(static_cast<int>(roundf(fb*15.0f)), radius, 15);
    }
    TryRec t1[kMaxTrys], t2[kMaxTrys];
    int n1 = 0, n2 = 0;
    const int i1 = try_half_radius(bc.src, m1, q1[0], q1[1], q1[2], radius, false, t1, n1, bc.rec709);
    const int i2 = try_half_radius(bc.src, m2, q2[0], q2[1], q2[2], radius, false, t2, n2, bc.rec709);
    const float err = t1[i1].err + t2[i2].err;
    if (err < best.err) {
        best.err = err; best.diff = false; best.flip = flip;
        best.r1 = t1[i1].r; best.g1 = t1[i1].g; best.b1 = t1[i1].b; best.r2 = t2[i2].r; best.g2 = t2[i2].g; best.b2 = t2[i2].b;
        best.cw1 = t1[i1].cw; best.cw2 = t2[i2].cw; best.sel1 = t1[i1].sel; best.sel2 = t2[i2].sel;
    }
}

// src: 16 texels in the reference's block order (column-major: pixel = x*4 + y).
// effort: etc2comp's, as EtcConverter maps Texture::Quality to it (lib/src/EtcConverter.cpp:34-51: 0, 20, 40, 70, 100).  One
// Etc::Image per block means the effort percentage of EtcImage.cpp:276-330 is all or nothing: up to 40 only encoding
// iteration 0 runs, at 70 and 100 the block is iterated until Block4x4Encoding_ETC1::PerformIteration (:232-306) says done.
CFX_HD uint2 encode_etc1_exact(const Px* src, float effort = 40.0f, bool rec709 = false)
{
    BlockCtx bc;
    bc.src = src; bc.rec709 = rec709;
    // CalculateSourceAverages (RGBX branch): quadrant sums, border texels count as (0,0,0)
    float ul[3], ll[3], ur[3], lr[3];
    {
        auto q = [&](int a, int b, int c, int d, float* o) {
            o[0] = ((src[a].r + src[b].r) + src[c].r) + src[d].r;
            o[1] = ((src[a].g + src[b].g) + src[c].g) + src[d].g;
            o[2] = ((src[a].b + src[b].b) + src[c].b) + src[d].b;
        };
        q(0, 1, 4, 5, ul); q(2, 3, 6, 7, ll); q(8, 9, 12, 13, ur); q(10, 11, 14, 15, lr);
    }
    bool partial = false;
    for (int i = 0; i < 16; ++i) partial = partial || src[i].a != src[i].a;
    if (rec709 && partial) {
        // a block with border texels is "translucent" under every metric but RGBX: alpha-weighted averages, NaN alpha
        // counts as 0 (CalculateSourceAverages, :413-520); the border texels are (0, 0, 0), so the sums above stand
        auto wq = [&](int a, int b, int c, int d) {
            const float w0 = src[a].a != src[a].a ? 0.0f : src[a].a, w1 = src[b].a != src[b].a ? 0.0f : src[b].a;
            const float w2 = src[c].a != src[c].a ? 0.0f : src[c].a, w3 = src[d].a != src[d].a ? 0.0f : src[d].a;
            return ((w0 + w1) + w2) + w3;
        };
        const float wul = wq(0, 1, 4, 5), wll = wq(2, 3, 6, 7), wur = wq(8, 9, 12, 13), wlr = wq(10, 11, 14, 15);
        const float wL = wul + wll, wR = wur + wlr, wT = wul + wur, wB = wll + wlr;
        for (int c = 0; c < 3; ++c) {
            if (wL > 0.0f) bc.avgL[c] = (ul[c] + ll[c])*(1.0f/wL);
            if (wR > 0.0f) bc.avgR[c] = (ur[c] + lr[c])*(1.0f/wR);
            if (wT > 0.0f) bc.avgT[c] = (ul[c] + ur[c])*(1.0f/wT);
            if (wB > 0.0f) bc.avgB[c] = (ll[c] + lr[c])*(1.0f/wB);
        }
        for (int c = 0; c < 3; ++c) {
            if (wL == 0.0f) bc.avgL[c] = bc.avgR[c];
            if (wR == 0.0f) bc.avgR[c] = bc.avgL[c];
            if (wT == 0.0f) bc.avgT[c] = bc.avgB[c];
            if (wB == 0.0f) bc.avgB[c] = bc.avgT[c];
        }
    } else
    for (int c = 0; c < 3; ++c) {
        bc.avgL[c] = (ul[c] + ll[c])*0.125f; bc.avgR[c] = (ur[c] + lr[c])*0.125f;
        bc.avgT[c] = (ul[c] + ur[c])*0.125f; bc.avgB[c] = (ll[c] + lr[c])*0.125f;
    }
    // CalculateMostLikelyFlip
    float eL = 0.0f, eR = 0.0f, eT = 0.0f, eB = 0.0f;
    for (uint32_t i = 0; i < 8; ++i) {
        const float l = gray_distance2(src[i], bc.avgL), r = gray_distance2(src[i + 8], bc.avgR);
        const float t = gray_distance2(src[kMapT[i]], bc.avgT), b = gray_distance2(src[kMapB[i]], bc.avgB);
        eL += l; eR += r; eT += t; eB += b;
    }
    const bool likely = (eT + eB) < (eL + eR);

    Encoding best;
    best.err = 3.402823466e+38f; best.diff = true; best.flip = false;
    best.r1 = best.g1 = best.b1 = best.r2 = best.g2 = best.b2 = 0; best.cw1 = best.cw2 = best.sel1 = best.sel2 = 0;
    // encoding iteration 0 (PerformFirstIteration): radius 0, SetDoneIfPerfect after every try
    try_differential(bc, likely, 0, 0, 0, best);
    if (best.err != 0.0f) try_individual(bc, likely, 0, best);
    if (best.err != 0.0f) try_differential(bc, !likely, 0, 0, 0, best);
    if (best.err != 0.0f) try_individual(bc, !likely, 0, best);
    // later iterations: radius 1 on both flips, then the "degenerate" tries (gray offsets on the base colours)
    if (effort > 40.0f) {
#pragma unroll 1
        for (int it = 1; it <= 8 && best.err != 0.0f; ++it) {
            bool done = false;
            switch (it) {
                case 1: try_differential(bc, likely, 1, 0, 0, best); break;
                case 2: try_individual(bc, likely, 1, best); done = effort <= 49.5f; break;
                case 3: try_differential(bc, !likely, 1, 0, 0, best); done = effort <= 59.5f; break;
                case 4: try_individual(bc, !likely, 1, best); done = effort <= 69.5f; break;
                case 5: case 6: case 7: case 8: {
                    const bool f = it == 6 ? !likely : likely;
                    const int m = it == 8 ? 4 : 2;
                    if (it == 7) {
                        try_differential(bc, f, 1, -2, -2, best); try_differential(bc, f, 1, -2, 2, best);
                        try_differential(bc, f, 1, 2, -2, best); try_differential(bc, f, 1, 2, 2, best);
                    } else {
                        try_differential(bc, f, 1, -m, 0, best); try_differential(bc, f, 1, m, 0, best);
                        try_differential(bc, f, 1, 0, m, best); try_differential(bc, f, 1, 0, -m, best);
                    }
                    done = it == 5 ? effort <= 79.5f : (it == 6 ? effort <= 89.5f : (it == 7 ? effort <= 99.5f : true));
                    break;
                }
            }
            if (done) break;
        }
    }
    const uint32_t* mapL = kMapL; const uint32_t* mapR = kMapR; const uint32_t* mapT = kMapT; const uint32_t* mapB = kMapB;

    // SetEncodingBits
    uint32_t hi = 0;
    if (best.diff) {
        hi |= (static_cast<uint32_t>(best.r1) << 27) | ((static_cast<uint32_t>(best.r2 - best.r1) & 7u) << 24);
        hi |= (static_cast<uint32_t>(best.g1) << 19) | ((static_cast<uint32_t>(best.g2 - best.g1) & 7u) << 16);
        hi |= (static_cast<uint32_t>(best.b1) << 11) | ((static_cast<uint32_t>(best.b2 - best.b1) & 7u) << 8);
    } else {
        hi |= ((static_cast<uint32_t>(best.r1) & 15u) << 28) | ((static_cast<uint32_t>(best.r2) & 15u) << 24);
        hi |= ((static_cast<uint32_t>(best.g1) & 15u) << 20) | ((static_cast<uint32_t>(best.g2) & 15u) << 16);
        hi |= ((static_cast<uint32_t>(best.b1) & 15u) << 12) | ((static_cast<uint32_t>(best.b2) & 15u) << 8);
    }
    hi |= (best.cw1 << 5) | (best.cw2 << 2) | ((best.diff ? 1u : 0u) << 1) | (best.flip ? 1u : 0u);
    const uint32_t* m1 = best.flip ? mapT : mapL;
    const uint32_t* m2 = best.flip ? mapB : mapR;
    uint32_t lo = 0;
    for (uint32_t i = 0; i < 8; ++i) {
        const uint32_t s1 = (best.sel1 >> (2*i)) & 3u, s2 = (best.sel2 >> (2*i)) & 3u;
        lo |= ((s1 >> 1) << (16u + m1[i])) | ((s1 & 1u) << m1[i]);
        lo |= ((s2 >> 1) << (16u + m2[i])) | ((s2 & 1u) << m2[i]);
    }
    return make_uint2(__byte_perm(hi, 0, 0x0123), __byte_perm(lo, 0, 0x0123));
}

} // namespace etc1x
} // namespace cfx
