// Byte-exact ETC1 for Quality::Lowest / Low / Normal in linear colour space: a restatement of what
// etc2comp computes when EtcConverter::process (lib/src/EtcConverter.cpp:120-152) encodes one block at
// effort <= 40 -- only encoding iteration 0 runs (lib/etc2comp/EtcLib/Etc/EtcImage.cpp:282,
// Block4x4Encoding_ETC1::PerformFirstIteration, EtcCodec/EtcBlock4x4Encoding_ETC1.cpp:311-338):
//   source averages of the four halves (:402-410), most likely flip from "gray line" distances
//   (:350-386, EtcBlock4x4Encoding_ETC1.h:138-152), then differential and individual tries at radius 0
//   for that flip and for the other one (:545-690, :807-900; base colours from
//   EtcDifferentialTrys.cpp:41-150 / EtcIndividualTrys.cpp:41-64), error metric RGBX
//   (EtcBlock4x4Encoding.cpp:144-154), strict '<' everywhere, SetDoneIfPerfect early exits.
// Byte parity needs the same IEEE single-precision operations in the same order and no fused
// multiply-add: etc.cu is compiled with -fmad=false.  ONE LANE OWNS ONE BLOCK.
#pragma once
#include "hostdev.h"

namespace cfx {
namespace etc1x {

struct Px { float r, g, b, a; };      // a = NaN marks a border texel (outside the image)

CFX_HD float clamp01(float v) { if (v < 0.0f) v = 0.0f; if (v > 1.0f) v = 1.0f; return v; }

// codeword table entries as the reference's compile-time floats k/255
CFX_HD float cw_delta(uint32_t cw, uint32_t sel)
{
    const float s[8] = {2.0f/255.0f, 5.0f/255.0f, 9.0f/255.0f, 13.0f/255.0f, 18.0f/255.0f, 24.0f/255.0f, 33.0f/255.0f, 47.0f/255.0f};
    const float l[8] = {8.0f/255.0f, 17.0f/255.0f, 29.0f/255.0f, 42.0f/255.0f, 60.0f/255.0f, 80.0f/255.0f, 106.0f/255.0f, 183.0f/255.0f};
    const float v = (sel & 1u) ? l[cw] : s[cw];
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
CFX_HD void try_half(const Px* src, const uint32_t* mapping, float cr, float cg, float cb, HalfTry& t)
{
    t.err = 3.402823466e+38f; t.cw = 0; t.sel = 0;
    for (uint32_t cw = 0; cw < 8; ++cw) {
        float sr[4], sg[4], sb[4];
        for (uint32_t k = 0; k < 4; ++k) {
            const float d = cw_delta(cw, k);
            sr[k] = clamp01(cr + d); sg[k] = clamp01(cg + d); sb[k] = clamp01(cb + d);
        }
        uint32_t sels = 0;
        float cw_err = 0.0f;
        for (uint32_t i = 0; i < 8; ++i) {
            const Px& s = src[mapping[i]];
            float best = 3.402823466e+38f;
            uint32_t bsel = 0;
            for (uint32_t k = 0; k < 4; ++k) {
                float e;
                if (s.a != s.a) e = 0.0f;                       // border texel
                else {
                    const float dr = sr[k] - s.r, dg = sg[k] - s.g, db = sb[k] - s.b, da = 1.0f - s.a;
                    e = dr*dr + dg*dg + db*db + da*da;
                }
                if (e < best) { best = e; bsel = k; }
            }
            sels |= bsel << (2*i);
            cw_err += best;
        }
        if (cw_err < t.err) { t.cw = cw; t.sel = sels; t.err = cw_err; }
    }
}

CFX_HD int quant_component(float v, float scale)   // Quantize..().Int..(scale): round(clamp(v)*scale), expanded, back to int
{
    return static_cast<int>(roundf(scale*clamp01(v)));
}

struct Encoding { bool diff, flip; int r1, g1, b1, r2, g2, b2; uint32_t cw1, cw2, sel1, sel2; float err; };

CFX_HD void bend(int& c1, int& c2)
{
    const int d = c2 - c1;
    if (d > 3) { c1 += (d - 3)/2; c2 = c1 + 3; }
    else if (d < -4) { c1 += (d + 4)/2; c2 = c1 - 4; }
}

// src: 16 texels in the reference's block order (column-major: pixel = x*4 + y)
CFX_HD uint2 encode_etc1_exact(const Px* src)
{
    const uint32_t mapL[8] = {0, 1, 2, 3, 4, 5, 6, 7}, mapR[8] = {8, 9, 10, 11, 12, 13, 14, 15};
    const uint32_t mapT[8] = {0, 1, 4, 5, 8, 9, 12, 13}, mapB[8] = {2, 3, 6, 7, 10, 11, 14, 15};
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
    float avgL[3], avgR[3], avgT[3], avgB[3];
    for (int c = 0; c < 3; ++c) {
        avgL[c] = (ul[c] + ll[c])*0.125f; avgR[c] = (ur[c] + lr[c])*0.125f;
        avgT[c] = (ul[c] + ur[c])*0.125f; avgB[c] = (ll[c] + lr[c])*0.125f;
    }
    // CalculateMostLikelyFlip
    float eL = 0.0f, eR = 0.0f, eT = 0.0f, eB = 0.0f;
    for (uint32_t i = 0; i < 8; ++i) {
        const float l = gray_distance2(src[i], avgL), r = gray_distance2(src[i + 8], avgR);
        const float t = gray_distance2(src[mapT[i]], avgT), b = gray_distance2(src[mapB[i]], avgB);
        eL += l; eR += r; eT += t; eB += b;
    }
    const bool likely_flip = (eT + eB) < (eL + eR);

    Encoding best;
    best.err = 3.402823466e+38f; best.diff = true; best.flip = false;
    best.r1 = best.g1 = best.b1 = best.r2 = best.g2 = best.b2 = 0; best.cw1 = best.cw2 = best.sel1 = best.sel2 = 0;
    for (int step = 0; step < 4; ++step) {
        const bool flip = step < 2 ? likely_flip : !likely_flip;
        const bool diff = (step & 1) == 0;
        const float* c1 = flip ? avgT : avgL;
        const float* c2 = flip ? avgB : avgR;
        const uint32_t* m1 = flip ? mapT : mapL;
        const uint32_t* m2 = flip ? mapB : mapR;
        int q1[3], q2[3];
        float f1[3], f2[3];
        if (diff) {
            for (int c = 0; c < 3; ++c) {
                // QuantizeR5G5B5 then IntX(31): round(31*clamp(v)) -> expand -> *(1/255) -> round(*31)
                const uint32_t a5 = static_cast<uint32_t>(roundf(31.0f*clamp01(c1[c]))), b5 = static_cast<uint32_t>(roundf(31.0f*clamp01(c2[c])));
                const float fa = (1.0f/255.0f)*static_cast<float>((a5 << 3) + (a5 >> 2)), fb = (1.0f/255.0f)*static_cast<float>((b5 << 3) + (b5 >> 2));
                q1[c] = min(max(static_cast<int>(roundf(fa*31.0f)), 0), 31);
                q2[c] = min(max(static_cast<int>(roundf(fb*31.0f)), 0), 31);
                bend(q1[c], q2[c]);
                f1[c] = static_cast<float>(static_cast<unsigned char>((q1[c] << 3) + (q1[c] >> 2)))/255.0f;
                f2[c] = static_cast<float>(static_cast<unsigned char>((q2[c] << 3) + (q2[c] >> 2)))/255.0f;
            }
        } else {
            for (int c = 0; c < 3; ++c) {
                const uint32_t a4 = static_cast<uint32_t>(roundf(15.0f*clamp01(c1[c]))), b4 = static_cast<uint32_t>(roundf(15.0f*clamp01(c2[c])));
                const float fa = (1.0f/255.0f)*static_cast<float>((a4 << 4) + a4), fb = (1.0f/255.0f)*static_cast<float>((b4 << 4) + b4);
                // IndividualTrys uses the same MoveAwayFromEdge (clamp to [0, 31]) as the differential path
                q1[c] = min(max(static_cast<int>(roundf(fa*15.0f)), 0), 31);
                q2[c] = min(max(static_cast<int>(roundf(fb*15.0f)), 0), 31);
                f1[c] = static_cast<float>(static_cast<unsigned char>((q1[c] << 4) + q1[c]))/255.0f;
                f2[c] = static_cast<float>(static_cast<unsigned char>((q2[c] << 4) + q2[c]))/255.0f;
            }
        }
        HalfTry t1, t2;
        try_half(src, m1, f1[0], f1[1], f1[2], t1);
        try_half(src, m2, f2[0], f2[1], f2[2], t2);
        const float err = t1.err + t2.err;
        if (err < best.err) {
            best.err = err; best.diff = diff; best.flip = flip;
            best.r1 = q1[0]; best.g1 = q1[1]; best.b1 = q1[2]; best.r2 = q2[0]; best.g2 = q2[1]; best.b2 = q2[2];
            best.cw1 = t1.cw; best.cw2 = t2.cw; best.sel1 = t1.sel; best.sel2 = t2.sel;
        }
        if (best.err == 0.0f) break;                                 // SetDoneIfPerfect
    }

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
