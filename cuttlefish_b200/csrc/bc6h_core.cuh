// Lane-local BC6H (unsigned half float) encoder: ONE LANE OWNS ONE BLOCK.  Our own search, not a
// port of Compressonator's CompressBlockBC6 (lib/compressonator/cmp_core/shaders/
// bc6_encode_kernel.cpp:4365-4411), which is what Bc6HConverter::compressBlock
// (lib/src/S3tcConverter.cpp:573-590) calls in the ISPC=0 configuration.
//
// Search: one-region modes 11-14 (10.10 direct, 11.9 / 12.8 / 16.4 delta) and, on the best of the
// 32 two-region shapes (line-fit residual), modes 10 (6.6.6.6 direct), 2 (7.6.6.6 delta) and
// 1 (10.5.5.5 delta).  Every fit is PCA end points -> quantise -> exact index search against the
// decoder's palette -> two rounds of least-squares end points; the mode with the smallest squared
// error in the decoder's pre-"finish" integer domain wins.
// Compiles for the device and, through hostdev.h, for tools/emu_bc6h.cpp.
#pragma once
#ifdef __CUDACC__
#include <cuda_fp16.h>
#endif
#include "bc7_tables.cuh"
#include "hostdev.h"

namespace cfx {
namespace bc6h {

CFX_CONST uint8_t kW4[16] = {0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64};
CFX_CONST uint8_t kW3[8] = {0, 9, 18, 27, 37, 46, 55, 64};

// texel channel c of texel t of the lane's block: xs[(t*3 + c)*32 + lane], values in the decoder's
// unquantised domain (half bits * 64 / 31, 0..65535)
CFX_HD float& px(float* xs, uint32_t lane, uint32_t t, uint32_t c) { return xs[(t*3u + c)*32u + lane]; }

// Error metric.  The search runs in the decoder's integer domain (half BITS, i.e. a piecewise-linear logarithm of the
// value), but the reference minimises -- and the parity tests measure -- the squared error of the decoded FLOAT values
// (Compressonator converts the block to float before it searches, bc6_encode_kernel.cpp).  So:
//  * the index search and every error that decides between fits, shapes and modes is the EXACT squared error of the
//    decoded values (half_value of the palette entry against half_value of the texel);
//  * the continuous steps (moments, principal axis, least-squares end points) stay in the integer domain, where the
//    palette is a straight line, with every texel channel weighted by its local slope^2 = 4^(e - e_max of the block),
//    e = the half's exponent -- floored at 4^-3 so that the shadows of a block with highlights keep a say in where
//    the line goes (their error is re-measured exactly afterwards).
// xs[(48 + t)*32 + lane] holds texel t's three exponent distances (5 bits each) and, at bit 15, the smallest of them
// (the texel's weight where one number per texel is needed: moments, principal axis).
CFX_HD float half_bits_value(uint32_t h)
{
#ifdef __CUDA_ARCH__
    return __half2float(__ushort_as_half(static_cast<unsigned short>(h)));
#else
    const int e = static_cast<int>(h >> 10) & 31, m = static_cast<int>(h & 1023u);
    return e ? ldexpf(static_cast<float>(1024 + m), e - 25) : ldexpf(static_cast<float>(m), -24);
#endif
}
// value the decoder's "finish" step gives an integer-domain colour p (unsigned: p*31/64 as half bits; signed: sign and
// |p|*31/32)
CFX_HD float value_of(int p, bool sg)
{
    if (!sg) return half_bits_value(static_cast<uint32_t>(max(p, 0)*31) >> 6);
    const float v = half_bits_value(static_cast<uint32_t>(abs(p)*31) >> 5);
    return p < 0 ? -v : v;
}
constexpr int kMaxWeightShift = 2, kMaxWeightShiftSigned = 8;
constexpr uint32_t kWordsPerLane = 64;     // 48 texel channels + 16 weight words
CFX_HD uint32_t& pw(float* xs, uint32_t lane, uint32_t t) { return reinterpret_cast<uint32_t*>(xs)[(48u + t)*32u + lane]; }
CFX_HD float w_of(uint32_t d) { return __uint_as_float((127u - 2u*(d & 31u)) << 23); }        // 4^-d
CFX_HD void prepare_weights(float* xs, uint32_t lane, bool sg)
{
    const float to_half = sg ? 31.0f/32.0f : 31.0f/64.0f;
    int emax = 1;
    for (uint32_t t = 0; t < 16; ++t)
#pragma unroll
        for (uint32_t c = 0; c < 3; ++c) emax = max(emax, __float2int_rn(fabsf(px(xs, lane, t, c))*to_half) >> 10);
    for (uint32_t t = 0; t < 16; ++t) {
        uint32_t word = 0, dmin = 31;
#pragma unroll
        for (uint32_t c = 0; c < 3; ++c) {
            const int e = max(__float2int_rn(fabsf(px(xs, lane, t, c))*to_half) >> 10, 1);
            const uint32_t d = static_cast<uint32_t>(min(emax - e, sg ? kMaxWeightShiftSigned : kMaxWeightShift));
            word |= d << (5u*c);
            dmin = min(dmin, d);
        }
        pw(xs, lane, t) = word | (dmin << 15);
    }
}

// sg: signed format (BC6H SF16, Texture::Type::Float -> SetSignedBC6, lib/src/S3tcConverter.cpp:566-570): end points are
// two's complement, the unquantised domain is +-0x7FFF and texels are sign * half magnitude * 32/31.
CFX_HD int unquantize(int x, int bits, bool sg)
{
    if (sg) {
        if (bits >= 16) return x;
        const bool neg = x < 0;
        const int ax = neg ? -x : x;
        int u;
        if (ax == 0) u = 0;
        else if (ax >= (1 << (bits - 1)) - 1) u = 0x7FFF;
        else u = ((ax << 15) + 0x4000) >> (bits - 1);
        return neg ? -u : u;
    }
    if (bits >= 15) return x;
    if (x == 0) return 0;
    if (x == (1 << bits) - 1) return 0xFFFF;
    return ((x << 15) + 0x4000) >> (bits - 1);
}

CFX_HD_NOINLINE int quantize(float u, int bits, bool sg)
{
    if (sg) {
        const int maxq = (1 << (bits - 1)) - 1;
        const bool neg = u < 0.0f;
        const float au = fminf(fabsf(u), 32767.0f);
        int best;
        if (bits >= 16) best = min(__float2int_rn(au), 32767);
        else {
            const int x = min(max(static_cast<int>(au*(1.0f/static_cast<float>(1 << (16 - bits)))), 0), maxq);
            best = x;
            float bd = fabsf(static_cast<float>(unquantize(x, bits, true)) - au);
            if (x > 0) { const float d = fabsf(static_cast<float>(unquantize(x - 1, bits, true)) - au); if (d < bd) { bd = d; best = x - 1; } }
            if (x < maxq) { const float d = fabsf(static_cast<float>(unquantize(x + 1, bits, true)) - au); if (d < bd) { bd = d; best = x + 1; } }
        }
        return neg ? -best : best;
    }
    const int maxq = (1 << bits) - 1;
    u = fminf(fmaxf(u, 0.0f), 65535.0f);
    if (bits >= 15) return min(max(__float2int_rn(u), 0), maxq);
    int x = min(max(static_cast<int>(u*(1.0f/static_cast<float>(1 << (16 - bits)))), 0), maxq);
    // nearest of x-1, x, x+1 under the real unquantiser (the end codes are special-cased by it)
    int best = x;
    float bd = fabsf(static_cast<float>(unquantize(x, bits, false)) - u);
    if (x > 0) { const float d = fabsf(static_cast<float>(unquantize(x - 1, bits, false)) - u); if (d < bd) { bd = d; best = x - 1; } }
    if (x < maxq) { const float d = fabsf(static_cast<float>(unquantize(x + 1, bits, false)) - u); if (d < bd) { bd = d; best = x + 1; } }
    return best;
}

struct SubsetFit {
    int q0[3], q1[3];       // quantised end points (wBits wide)
    float err;
    uint64_t idx;           // 4 bits per texel (only the subset's texels are meaningful)
};

// Exact index search for the texels of `mask` against the palette of (q0, q1); returns the error.
CFX_HD_NOINLINE float assign_indices(float* xs, uint32_t lane, uint32_t mask, const int* q0, const int* q1, int wbits, int ibits,
    uint64_t& idx_out, bool sg)
{
    const int n = 1 << ibits;
    int a[3], d[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        a[c] = unquantize(q0[c], wbits, sg);
        d[c] = unquantize(q1[c], wbits, sg) - a[c];
    }
    float err = 0.0f;
    uint64_t idx = idx_out;
    for (uint32_t t = 0; t < 16; ++t) {
        if (!((mask >> t) & 1u)) continue;
        const float x0 = px(xs, lane, t, 0), x1 = px(xs, lane, t, 1), x2 = px(xs, lane, t, 2);
        const uint32_t wd = pw(xs, lane, t);
        const float w0 = w_of(wd), w1 = w_of(wd >> 5), w2 = w_of(wd >> 10);
        const float f0 = value_of(__float2int_rn(x0), sg), f1 = value_of(__float2int_rn(x1), sg), f2 = value_of(__float2int_rn(x2), sg);
        // weighted projection on the palette line
        const float d0 = static_cast<float>(d[0]), d1 = static_cast<float>(d[1]), d2 = static_cast<float>(d[2]);
        const float len2 = w0*d0*d0 + w1*d1*d1 + w2*d2*d2;
        const float proj = len2 > 0.0f ? (w0*(x0 - static_cast<float>(a[0]))*d0 + w1*(x1 - static_cast<float>(a[1]))*d1 +
            w2*(x2 - static_cast<float>(a[2]))*d2)*static_cast<float>(n - 1)/len2 : 0.0f;
        const int k0 = min(max(__float2int_rn(proj), 0), n - 1);
        // the nearest index along the line and its neighbour on the side the texel lies on
        const int k1 = min(max(proj > static_cast<float>(k0) ? k0 + 1 : k0 - 1, 0), n - 1);
        float beste = 3.0e38f;
        int bestk = k0;
#pragma unroll
        for (int dk = 0; dk < 2; ++dk) {
            const int k = dk ? k1 : k0;
            const int w = ibits == 4 ? kW4[k] : kW3[k];
            float e = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int p = a[c] + ((d[c]*w + 32) >> 6);          // == (a*(64-w) + b*w + 32) >> 6
                const float df = value_of(p, sg) - (c == 0 ? f0 : (c == 1 ? f1 : f2));
                e += df*df;
            }
            if (e < beste) { beste = e; bestk = k; }
        }
        err += beste;
        idx = (idx & ~(15ull << (4*t))) | (static_cast<uint64_t>(bestk) << (4*t));
    }
    idx_out = idx;
    return err;
}

// Full fit of one subset: PCA -> quantise -> indices -> 2 x least squares.
CFX_HD_NOINLINE void fit_subset(float* xs, uint32_t lane, uint32_t mask, int wbits, int ibits, SubsetFit& f, bool sg)
{
    float n = 0.0f, m[3] = {0, 0, 0};
    for (uint32_t t = 0; t < 16; ++t) {
        if (!((mask >> t) & 1u)) continue;
        const float wt = w_of(pw(xs, lane, t) >> 15);
        n += wt; m[0] += wt*px(xs, lane, t, 0); m[1] += wt*px(xs, lane, t, 1); m[2] += wt*px(xs, lane, t, 2);
    }
    const float inv = n > 0.0f ? 1.0f/n : 0.0f;
    m[0] *= inv; m[1] *= inv; m[2] *= inv;
    float cv[6] = {0, 0, 0, 0, 0, 0};
    for (uint32_t t = 0; t < 16; ++t) {
        if (!((mask >> t) & 1u)) continue;
        const float wt = w_of(pw(xs, lane, t) >> 15);
        const float d0 = px(xs, lane, t, 0) - m[0], d1 = px(xs, lane, t, 1) - m[1], d2 = px(xs, lane, t, 2) - m[2];
        cv[0] += wt*d0*d0; cv[1] += wt*d0*d1; cv[2] += wt*d0*d2; cv[3] += wt*d1*d1; cv[4] += wt*d1*d2; cv[5] += wt*d2*d2;
    }
    float v[3] = {cv[0], cv[1], cv[2]};
    float best = cv[0];
    if (cv[3] > best) { best = cv[3]; v[0] = cv[1]; v[1] = cv[3]; v[2] = cv[4]; }
    if (cv[5] > best) { best = cv[5]; v[0] = cv[2]; v[1] = cv[4]; v[2] = cv[5]; }
    for (int it = 0; it < 5; ++it) {
        const float n2 = v[0]*v[0] + v[1]*v[1] + v[2]*v[2];
        const float s = n2 > 1e-30f ? rsqrtf(n2) : 0.0f;
        const float a0 = v[0]*s, a1 = v[1]*s, a2 = v[2]*s;
        v[0] = cv[0]*a0 + cv[1]*a1 + cv[2]*a2;
        v[1] = cv[1]*a0 + cv[3]*a1 + cv[4]*a2;
        v[2] = cv[2]*a0 + cv[4]*a1 + cv[5]*a2;
    }
    {
        const float n2 = v[0]*v[0] + v[1]*v[1] + v[2]*v[2];
        const float s = n2 > 1e-30f ? rsqrtf(n2) : 0.0f;
        v[0] *= s; v[1] *= s; v[2] *= s;
    }
    float tmin = 3.0e38f, tmax = -3.0e38f;
    for (uint32_t t = 0; t < 16; ++t) {
        if (!((mask >> t) & 1u)) continue;
        const float p = (px(xs, lane, t, 0) - m[0])*v[0] + (px(xs, lane, t, 1) - m[1])*v[1] + (px(xs, lane, t, 2) - m[2])*v[2];
        tmin = fminf(tmin, p); tmax = fmaxf(tmax, p);
    }
    if (!(tmax >= tmin)) { tmin = tmax = 0.0f; }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        f.q0[c] = quantize(m[c] + tmin*v[c], wbits, sg);
        f.q1[c] = quantize(m[c] + tmax*v[c], wbits, sg);
    }
    f.idx = 0;
    f.err = assign_indices(xs, lane, mask, f.q0, f.q1, wbits, ibits, f.idx, sg);
    for (int round = 0; round < 2 && f.err > 0.0f; ++round) {
        // weighted normal equations (one weight per texel: its brightest channel's)
        float A = 0, B = 0, C = 0, P[3] = {0, 0, 0}, Q[3] = {0, 0, 0};
        for (uint32_t t = 0; t < 16; ++t) {
            if (!((mask >> t) & 1u)) continue;
            const uint32_t k = static_cast<uint32_t>(f.idx >> (4*t)) & 15u;
            const float w = static_cast<float>(ibits == 4 ? kW4[k] : kW3[k])*(1.0f/64.0f), iw = 1.0f - w;
            const float wt = w_of(pw(xs, lane, t) >> 15), wiw = wt*iw, ww = wt*w;
            A += wiw*iw; B += wiw*w; C += ww*w;
#pragma unroll
            for (int c = 0; c < 3; ++c) { const float x = px(xs, lane, t, c); P[c] += wiw*x; Q[c] += ww*x; }
        }
        const float det = A*C - B*B;
        if (!(fabsf(det) > 1e-6f*(A + C)*(A + C))) break;      // (relative test: the weights span orders of magnitude)
        const float id = 1.0f/det;
        SubsetFit trial;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            trial.q0[c] = quantize((C*P[c] - B*Q[c])*id, wbits, sg);
            trial.q1[c] = quantize((A*Q[c] - B*P[c])*id, wbits, sg);
        }
        trial.idx = f.idx;
        trial.err = assign_indices(xs, lane, mask, trial.q0, trial.q1, wbits, ibits, trial.idx, sg);
        if (trial.err < f.err) f = trial; else break;
    }
}

// Line-fit residual of a two-region shape (sum over regions of trace - lambda_max).
CFX_HD float shape_score(float* xs, uint32_t lane, uint32_t m1, const float* sT, const float* cT, float wT)
{
    float n1 = 0, s1[3] = {0, 0, 0}, c1[6] = {0, 0, 0, 0, 0, 0};
    for (uint32_t t = 0; t < 16; ++t) {
        const float f = ((m1 >> t) & 1u) ? w_of(pw(xs, lane, t) >> 15) : 0.0f;
        const float x0 = px(xs, lane, t, 0), x1 = px(xs, lane, t, 1), x2 = px(xs, lane, t, 2);
        n1 += f; s1[0] += f*x0; s1[1] += f*x1; s1[2] += f*x2;
        c1[0] += f*x0*x0; c1[1] += f*x0*x1; c1[2] += f*x0*x2; c1[3] += f*x1*x1; c1[4] += f*x1*x2; c1[5] += f*x2*x2;
    }
    float score = 0.0f;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const float nn = s ? n1 : wT - n1;
        float sm[3], cc[6];
#pragma unroll
        for (int k = 0; k < 3; ++k) sm[k] = s ? s1[k] : sT[k] - s1[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) cc[k] = s ? c1[k] : cT[k] - c1[k];
        const float inv = nn > 0.0f ? 1.0f/nn : 0.0f;
        cc[0] -= sm[0]*sm[0]*inv; cc[1] -= sm[0]*sm[1]*inv; cc[2] -= sm[0]*sm[2]*inv;
        cc[3] -= sm[1]*sm[1]*inv; cc[4] -= sm[1]*sm[2]*inv; cc[5] -= sm[2]*sm[2]*inv;
        float v[3] = {cc[0], cc[1], cc[2]};
        float best = cc[0];
        if (cc[3] > best) { best = cc[3]; v[0] = cc[1]; v[1] = cc[3]; v[2] = cc[4]; }
        if (cc[5] > best) { best = cc[5]; v[0] = cc[2]; v[1] = cc[4]; v[2] = cc[5]; }
        float lam = 0.0f;
        for (int it = 0; it < 3; ++it) {
            const float n2 = v[0]*v[0] + v[1]*v[1] + v[2]*v[2];
            const float sc = n2 > 1e-30f ? rsqrtf(n2) : 0.0f;
            const float a0 = v[0]*sc, a1 = v[1]*sc, a2 = v[2]*sc;
            v[0] = cc[0]*a0 + cc[1]*a1 + cc[2]*a2;
            v[1] = cc[1]*a0 + cc[3]*a1 + cc[4]*a2;
            v[2] = cc[2]*a0 + cc[4]*a1 + cc[5]*a2;
            lam = a0*v[0] + a1*v[1] + a2*v[2];
        }
        score += fmaxf(cc[0] + cc[3] + cc[5] - lam, 0.0f);
    }
    return score;
}

struct Bits128 {
    uint64_t lo, hi;
    CFX_HD void put(uint32_t pos, uint32_t v, uint32_t n)
    {
        const uint64_t vv = static_cast<uint64_t>(v) & ((1ull << n) - 1ull);
        if (pos < 64) {
            lo |= vv << pos;
            if (pos + n > 64) hi |= vv >> (64u - pos);
        } else {
            hi |= vv << (pos - 64u);
        }
    }
};

CFX_HD bool delta_fits(int d, int bits) { const int lim = (1 << (bits - 1)) - 1; return d >= -lim && d <= lim; }

// Encode the lane's block; returns the 16 bytes.
CFX_HD uint4 encode_block(float* xs, uint32_t lane, uint32_t quality, bool sg = false)
{
    prepare_weights(xs, lane, sg);
    // ---- one-region modes: {mode bits, wBits, delta bits}
    const int one_mode[4] = {0x03, 0x07, 0x0B, 0x0F}, one_w[4] = {10, 11, 12, 16}, one_t[4] = {10, 9, 8, 4};
    float best_err = 3.0e38f;
    int best_kind = -1;             // 0..3 one-region, 4..6 two-region (mode 10, 2, 1)
    SubsetFit bf0, bf1;
    bf0.err = bf1.err = 0.0f; bf0.idx = bf1.idx = 0;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        SubsetFit f;
        fit_subset(xs, lane, 0xFFFFu, one_w[k], 4, f, sg);
        if (k > 0 && !(delta_fits(f.q1[0] - f.q0[0], one_t[k]) && delta_fits(f.q1[1] - f.q0[1], one_t[k]) &&
            delta_fits(f.q1[2] - f.q0[2], one_t[k]))) continue;
        if (f.err < best_err) { best_err = f.err; best_kind = k; bf0 = f; }
    }
    // ---- two-region modes on the best shape
    uint32_t best_shape = 0;
    if (best_err > 0.0f && quality >= 1) {
        float sT[3] = {0, 0, 0}, cT[6] = {0, 0, 0, 0, 0, 0}, wT = 0.0f;
        for (uint32_t t = 0; t < 16; ++t) {
            const float x0 = px(xs, lane, t, 0), x1 = px(xs, lane, t, 1), x2 = px(xs, lane, t, 2);
            const float wt = w_of(pw(xs, lane, t) >> 15);
            wT += wt; sT[0] += wt*x0; sT[1] += wt*x1; sT[2] += wt*x2;
            cT[0] += wt*x0*x0; cT[1] += wt*x0*x1; cT[2] += wt*x0*x2; cT[3] += wt*x1*x1; cT[4] += wt*x1*x2; cT[5] += wt*x2*x2;
        }
        // the two best shapes by line-fit residual
        float bs0 = 3.0e38f, bs1 = 3.0e38f;
        uint32_t sh0 = 0, sh1 = 0;
#pragma unroll 1
        for (uint32_t s = 0; s < 32; ++s) {
            const float sc = shape_score(xs, lane, kBc7Part2[s], sT, cT, wT);
            if (sc < bs0) { bs1 = bs0; sh1 = sh0; bs0 = sc; sh0 = s; }
            else if (sc < bs1) { bs1 = sc; sh1 = s; }
        }
        // two-region precisions tried: mode 10 (6.6.6.6 direct), 2 (7.6.6.6), 1 (10.5.5.5), 6 (9.5.5.5): the last one
        // catches the noisy blocks whose deltas overflow 5 bits at 10-bit precision
        const int two_w[4] = {6, 7, 10, 9}, two_t[4] = {6, 6, 5, 5};
        const uint32_t nshapes = quality >= 2 ? 2u : 1u;
#pragma unroll 1
        for (uint32_t si = 0; si < nshapes; ++si) {
            const uint32_t shape = si ? sh1 : sh0;
            const uint32_t m1 = kBc7Part2[shape], m0 = ~m1 & 0xFFFFu;
#pragma unroll 1
            for (int k = 0; k < 4; ++k) {
                SubsetFit f0, f1;
                fit_subset(xs, lane, m0, two_w[k], 3, f0, sg);
                fit_subset(xs, lane, m1, two_w[k], 3, f1, sg);
                if (k > 0) {
                    bool ok = true;
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        ok = ok && delta_fits(f0.q1[c] - f0.q0[c], two_t[k]) && delta_fits(f1.q0[c] - f0.q0[c], two_t[k]) &&
                            delta_fits(f1.q1[c] - f0.q0[c], two_t[k]) && delta_fits(f0.q0[c] - f0.q1[c], two_t[k]) &&
                            delta_fits(f1.q0[c] - f0.q1[c], two_t[k]) && delta_fits(f1.q1[c] - f0.q1[c], two_t[k]);
                    if (!ok) continue;
                }
                if (f0.err + f1.err < best_err) { best_err = f0.err + f1.err; best_kind = 4 + k; bf0 = f0; bf1 = f1; best_shape = shape; }
            }
        }
    }

    Bits128 b; b.lo = b.hi = 0;
    if (best_kind < 4) {
        const int k = best_kind < 0 ? 0 : best_kind;
        const int wb = one_w[k], tb = one_t[k];
        // anchor: index of texel 0 must have a clear MSB
        if ((bf0.idx & 15ull) >= 8) {
#pragma unroll
            for (int c = 0; c < 3; ++c) { const int tmp = bf0.q0[c]; bf0.q0[c] = bf0.q1[c]; bf0.q1[c] = tmp; }
            bf0.idx = 0xFFFFFFFFFFFFFFFFull - bf0.idx;      // 15 - idx per nibble
        }
        b.put(0, static_cast<uint32_t>(one_mode[k]), 5);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint32_t w = static_cast<uint32_t>(bf0.q0[c]);
            const uint32_t x = (k == 0 ? static_cast<uint32_t>(bf0.q1[c]) : static_cast<uint32_t>(bf0.q1[c] - bf0.q0[c])) & ((1u << tb) - 1u);
            b.put(5 + 10*c, w & 1023u, 10);
            b.put(35 + 10*c, x, tb);
            // high bits of w, mirrored, fill the rest of the 10-bit field up to bit 44 + 10c
            for (int bit = 10; bit < wb; ++bit) b.put(35 + 10*c + (9 - (bit - 10)), (w >> bit) & 1u, 1);
        }
        b.put(65, static_cast<uint32_t>(bf0.idx & 7ull), 3);
        for (uint32_t t = 1; t < 16; ++t) b.put(65 + 3 + 4*(t - 1), static_cast<uint32_t>(bf0.idx >> (4*t)) & 15u, 4);
    } else {
        const int k = best_kind - 4;
        const uint32_t m1 = kBc7Part2[best_shape];
        const uint32_t anchor1 = kBc7Anchor2[best_shape];
        if ((bf0.idx & 15ull) >= 4) {
#pragma unroll
            for (int c = 0; c < 3; ++c) { const int tmp = bf0.q0[c]; bf0.q0[c] = bf0.q1[c]; bf0.q1[c] = tmp; }
            bf0.idx = 0x7777777777777777ull - bf0.idx;
        }
        if (((bf1.idx >> (4*anchor1)) & 15ull) >= 4) {
#pragma unroll
            for (int c = 0; c < 3; ++c) { const int tmp = bf1.q0[c]; bf1.q0[c] = bf1.q1[c]; bf1.q1[c] = tmp; }
            bf1.idx = 0x7777777777777777ull - bf1.idx;
        }
        const int tb = k <= 1 ? 6 : 5;
        const uint32_t tm = (1u << tb) - 1u;
        uint32_t w[3], x[3], y[3], z[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            w[c] = static_cast<uint32_t>(bf0.q0[c]) & ((1u << (k == 0 ? 6 : (k == 1 ? 7 : (k == 2 ? 10 : 9)))) - 1u);
            if (k == 0) { x[c] = static_cast<uint32_t>(bf0.q1[c]) & 63u; y[c] = static_cast<uint32_t>(bf1.q0[c]) & 63u; z[c] = static_cast<uint32_t>(bf1.q1[c]) & 63u; }
            else {
                x[c] = static_cast<uint32_t>(bf0.q1[c] - bf0.q0[c]) & tm;
                y[c] = static_cast<uint32_t>(bf1.q0[c] - bf0.q0[c]) & tm;
                z[c] = static_cast<uint32_t>(bf1.q1[c] - bf0.q0[c]) & tm;
            }
        }
        auto bit = [](uint32_t v, int i) { return (v >> i) & 1u; };
        if (k == 0) {           // mode 10: 6.6.6.6
            b.put(0, 0x1E, 5);
            b.put(5, w[0], 6); b.put(15, w[1], 6); b.put(25, w[2], 6);
            b.put(35, x[0], 6); b.put(45, x[1], 6); b.put(55, x[2], 6);
            b.put(65, y[0], 6); b.put(71, z[0], 6);
            b.put(41, y[1] & 15u, 4); b.put(24, bit(y[1], 4), 1); b.put(21, bit(y[1], 5), 1);
            b.put(61, y[2] & 15u, 4); b.put(14, bit(y[2], 4), 1); b.put(22, bit(y[2], 5), 1);
            b.put(51, z[1] & 15u, 4); b.put(11, bit(z[1], 4), 1); b.put(31, bit(z[1], 5), 1);
            b.put(12, bit(z[2], 0), 1); b.put(13, bit(z[2], 1), 1); b.put(23, bit(z[2], 2), 1);
            b.put(32, bit(z[2], 3), 1); b.put(34, bit(z[2], 4), 1); b.put(33, bit(z[2], 5), 1);
        } else if (k == 1) {    // mode 2: 7.6.6.6
            b.put(0, 0x01, 2);
            b.put(5, w[0], 7); b.put(15, w[1], 7); b.put(25, w[2], 7);
            b.put(35, x[0], 6); b.put(45, x[1], 6); b.put(55, x[2], 6);
            b.put(65, y[0], 6); b.put(71, z[0], 6);
            b.put(41, y[1] & 15u, 4); b.put(24, bit(y[1], 4), 1); b.put(2, bit(y[1], 5), 1);
            b.put(51, z[1] & 15u, 4); b.put(3, bit(z[1], 4), 1); b.put(4, bit(z[1], 5), 1);
            b.put(61, y[2] & 15u, 4); b.put(14, bit(y[2], 4), 1); b.put(22, bit(y[2], 5), 1);
            b.put(12, bit(z[2], 0), 1); b.put(13, bit(z[2], 1), 1); b.put(23, bit(z[2], 2), 1);
            b.put(32, bit(z[2], 3), 1); b.put(34, bit(z[2], 4), 1); b.put(33, bit(z[2], 5), 1);
        } else if (k == 3) {    // mode 6: 9.5.5.5
            b.put(0, 0x0E, 5);
            b.put(5, w[0], 9); b.put(15, w[1], 9); b.put(25, w[2], 9);
            b.put(35, x[0], 5); b.put(45, x[1], 5); b.put(55, x[2], 5);
            b.put(65, y[0], 5); b.put(71, z[0], 5);
            b.put(41, y[1] & 15u, 4); b.put(24, bit(y[1], 4), 1);
            b.put(51, z[1] & 15u, 4); b.put(40, bit(z[1], 4), 1);
            b.put(61, y[2] & 15u, 4); b.put(14, bit(y[2], 4), 1);
            b.put(50, bit(z[2], 0), 1); b.put(60, bit(z[2], 1), 1); b.put(70, bit(z[2], 2), 1);
            b.put(76, bit(z[2], 3), 1); b.put(34, bit(z[2], 4), 1);
        } else {                // mode 1: 10.5.5.5
            b.put(0, 0x00, 2);
            b.put(5, w[0], 10); b.put(15, w[1], 10); b.put(25, w[2], 10);
            b.put(35, x[0], 5); b.put(45, x[1], 5); b.put(55, x[2], 5);
            b.put(65, y[0], 5); b.put(71, z[0], 5);
            b.put(41, y[1] & 15u, 4); b.put(2, bit(y[1], 4), 1);
            b.put(51, z[1] & 15u, 4); b.put(40, bit(z[1], 4), 1);
            b.put(61, y[2] & 15u, 4); b.put(3, bit(y[2], 4), 1);
            b.put(50, bit(z[2], 0), 1); b.put(60, bit(z[2], 1), 1); b.put(70, bit(z[2], 2), 1);
            b.put(76, bit(z[2], 3), 1); b.put(4, bit(z[2], 4), 1);
        }
        b.put(77, best_shape, 5);
        uint32_t pos = 82;
        for (uint32_t t = 0; t < 16; ++t) {
            const bool in1 = (m1 >> t) & 1u;
            const uint32_t k3 = static_cast<uint32_t>((in1 ? bf1.idx : bf0.idx) >> (4*t)) & 7u;
            const uint32_t nb = (t == 0 || t == anchor1) ? 2u : 3u;
            b.put(pos, k3, nb); pos += nb;
        }
    }
    return make_uint4(static_cast<uint32_t>(b.lo), static_cast<uint32_t>(b.lo >> 32), static_cast<uint32_t>(b.hi),
        static_cast<uint32_t>(b.hi >> 32));
}

} // namespace bc6h
} // namespace cfx
