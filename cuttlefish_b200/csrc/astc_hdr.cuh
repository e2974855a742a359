// ASTC HDR pieces (Texture::Type::UFloat -> astcenc profiles HDR / HDR_RGB_LDR_A, lib/src/AstcConverter.cpp:151-163):
// the LNS representation texels are searched in, and colour end point mode 11 ("HDR RGB direct", ASTC specification
// section "HDR Endpoint Mode 11"; the reference's decoder for it is hdr_rgb_unpack,
// lib/astc-encoder/Source/astcenc_color_unquantize.cpp:442-599, its encoder quantize_hdr_rgb in
// astcenc_color_quantize.cpp).  The decoder below is the format's definition; the ENCODER is our own: every sub-mode is
// tried (lane = sub-mode in the kernel), each candidate is quantised to the block's colour level and DECODED again,
// and the candidate whose decoded end points are closest to the least-squares targets wins -- no case analysis to get
// wrong, and quantisation damage to the packed bit fields is measured rather than reasoned about.
// Compiles for the device and, through hostdev.h, for host tools.
#pragma once
#include "astc_core.cuh"

namespace cfx {
namespace astc {

// float -> 16-bit LNS (the inverse of the specification's LNS -> FP16 decode; same piecewise mantissa map as
// astcenc's float_to_lns, lib/astc-encoder/Source/astcenc_vecmathlib.h:558-596).  Result in [0, 65535].
CFX_HD float lns_from_float(float a)
{
    if (!(a > 1.0f/67108864.0f)) return 0.0f;             // underflow, negatives, NaN
    if (a >= 65536.0f) return 65535.0f;
    int e;
    float x;
#ifdef __CUDA_ARCH__
    const float mant = frexpf(a, &e);
#else
    const float mant = std::frexp(a, &e);
#endif
    if (e < -13) { x = a*33554432.0f; e = 0; }
    else { x = (mant - 0.5f)*4096.0f; e += 14; }
    if (x < 384.0f) x *= 4.0f/3.0f;
    else if (x <= 1408.0f) x += 128.0f;
    else x = (x + 512.0f)*(4.0f/5.0f);
    return x + static_cast<float>(e)*2048.0f + 1.0f;
}

// Decode of the six (unquantised, 0..255) values of end point mode 11 into two 12-bit RGB end points.
CFX_HD void hdr_rgb_decode(const int* v, int* e0, int* e1)
{
    const int modeval = ((v[1] & 0x80) >> 7) | (((v[2] & 0x80) >> 7) << 1) | (((v[3] & 0x80) >> 7) << 2);
    const int majcomp = ((v[4] & 0x80) >> 7) | (((v[5] & 0x80) >> 7) << 1);
    if (majcomp == 3) {
        e0[0] = v[0] << 4; e0[1] = v[2] << 4; e0[2] = (v[4] & 0x7F) << 5;
        e1[0] = v[1] << 4; e1[1] = v[3] << 4; e1[2] = (v[5] & 0x7F) << 5;
        return;
    }
    int a = v[0] | ((v[1] & 0x40) << 2);
    int b0 = v[2] & 0x3F, b1 = v[3] & 0x3F, c = v[1] & 0x3F, d0 = v[4] & 0x7F, d1 = v[5] & 0x7F;
    const int dbits = (modeval & 1) ? 6 : ((modeval & 4) ? 5 : 7);          // 7 6 7 6 5 6 5 6
    const int bit0 = (v[2] >> 6) & 1, bit1 = (v[3] >> 6) & 1, bit2 = (v[4] >> 6) & 1, bit3 = (v[5] >> 6) & 1;
    const int bit4 = (v[4] >> 5) & 1, bit5 = (v[5] >> 5) & 1;
    const int oh = 1 << modeval;
    if (oh & 0xA4) a |= bit0 << 9;
    if (oh & 0x08) a |= bit2 << 9;
    if (oh & 0x50) a |= bit4 << 9;
    if (oh & 0x50) a |= bit5 << 10;
    if (oh & 0xA0) a |= bit1 << 10;
    if (oh & 0xC0) a |= bit2 << 11;
    if (oh & 0x04) c |= bit1 << 6;
    if (oh & 0xE8) c |= bit3 << 6;
    if (oh & 0x20) c |= bit2 << 7;
    if (oh & 0x5B) { b0 |= bit0 << 6; b1 |= bit1 << 6; }
    if (oh & 0x12) { b0 |= bit2 << 7; b1 |= bit3 << 7; }
    if (oh & 0xAF) { d0 |= bit4 << 5; d1 |= bit5 << 5; }
    if (oh & 0x05) { d0 |= bit2 << 6; d1 |= bit3 << 6; }
    const int sx = 32 - dbits;
    d0 = static_cast<int>(static_cast<uint32_t>(d0) << sx) >> sx;
    d1 = static_cast<int>(static_cast<uint32_t>(d1) << sx) >> sx;
    const int sh = (modeval >> 1) ^ 3;
    a <<= sh; b0 <<= sh; b1 <<= sh; c <<= sh; d0 *= (1 << sh); d1 *= (1 << sh);
    int r1 = a, g1 = a - b0, bl1 = a - b1, r0 = a - c, g0 = a - b0 - c - d0, bl0 = a - b1 - c - d1;
    r0 = min(max(r0, 0), 4095); g0 = min(max(g0, 0), 4095); bl0 = min(max(bl0, 0), 4095);
    r1 = min(max(r1, 0), 4095); g1 = min(max(g1, 0), 4095); bl1 = min(max(bl1, 0), 4095);
    if (majcomp == 1) { int t = r0; r0 = g0; g0 = t; t = r1; r1 = g1; g1 = t; }
    else if (majcomp == 2) { int t = r0; r0 = bl0; bl0 = t; t = r1; r1 = bl1; bl1 = t; }
    e0[0] = r0; e0[1] = g0; e0[2] = bl0; e1[0] = r1; e1[1] = g1; e1[2] = bl1;
}

// Widths of the fields of sub-mode m (0..7): a, c, b, d.
CFX_HD void hdr_rgb_field_bits(int m, int& abits, int& cbits, int& bbits, int& dbits)
{
    abits = 9 + (m >> 1);                                  // 9 9 10 10 11 11 12 12
    const int oh = 1 << m;
    cbits = 6 + ((oh & 0xEC) ? 1 : 0) + ((oh & 0x20) ? 1 : 0);      // 6 6 7 7 6 8 7 7
    bbits = 6 + ((oh & 0x5B) ? 1 : 0) + ((oh & 0x12) ? 1 : 0);      // 7 8 6 7 8 6 7 6
    dbits = (m & 1) ? 6 : ((m & 4) ? 5 : 7);
}

// One candidate: sub-mode m (0..7) or the flat fallback (m == 8) for the 12-bit targets t0 (end point 0) and t1.
// Writes the six values, UNQUANTISED domain 0..255 and not yet snapped to a colour level; false if the targets do
// not fit the sub-mode's fields.
CFX_HD bool hdr_rgb_candidate(int m, const float* t0, const float* t1, int* v)
{
    if (m == 8) {
        // flat: 8 bits for red and green, 7 for blue, major component marker 3
        v[0] = min(max(__float2int_rn(t0[0]*(1.0f/16.0f)), 0), 255); v[1] = min(max(__float2int_rn(t1[0]*(1.0f/16.0f)), 0), 255);
        v[2] = min(max(__float2int_rn(t0[1]*(1.0f/16.0f)), 0), 255); v[3] = min(max(__float2int_rn(t1[1]*(1.0f/16.0f)), 0), 255);
        v[4] = min(max(__float2int_rn(t0[2]*(1.0f/32.0f)), 0), 127) | 0x80; v[5] = min(max(__float2int_rn(t1[2]*(1.0f/32.0f)), 0), 127) | 0x80;
        return true;
    }
    // major component: the largest channel of end point 1 goes first
    int maj = 0;
    if (t1[1] > t1[0] && t1[1] >= t1[2]) maj = 1;
    else if (t1[2] > t1[0] && t1[2] > t1[1]) maj = 2;
    float p0[3] = {t0[0], t0[1], t0[2]}, p1[3] = {t1[0], t1[1], t1[2]};
    if (maj == 1) { float t = p0[0]; p0[0] = p0[1]; p0[1] = t; t = p1[0]; p1[0] = p1[1]; p1[1] = t; }
    if (maj == 2) { float t = p0[0]; p0[0] = p0[2]; p0[2] = t; t = p1[0]; p1[0] = p1[2]; p1[2] = t; }
    int abits, cbits, bbits, dbits;
    hdr_rgb_field_bits(m, abits, cbits, bbits, dbits);
    const int sh = 12 - abits;
    const float is = 1.0f/static_cast<float>(1 << sh);
    const int a = min(max(__float2int_rn(p1[0]*is), 0), (1 << abits) - 1);
    const float af = static_cast<float>(a << sh);
    const int c = __float2int_rn((af - p0[0])*is), b0 = __float2int_rn((af - p1[1])*is), b1 = __float2int_rn((af - p1[2])*is);
    if (c < 0 || c >= (1 << cbits) || b0 < 0 || b0 >= (1 << bbits) || b1 < 0 || b1 >= (1 << bbits)) return false;
    // green0 = a - b0 - c - d0  =>  d0 = (a - b0 - c) - green0
    const int d0 = __float2int_rn((static_cast<float>((a - b0 - c) << sh) - p0[1])*is);
    const int d1 = __float2int_rn((static_cast<float>((a - b1 - c) << sh) - p0[2])*is);
    const int dlim = 1 << (dbits - 1);
    if (d0 < -dlim || d0 >= dlim || d1 < -dlim || d1 >= dlim) return false;
    // scatter the fields: the exact inverse of hdr_rgb_decode's gather
    const int oh = 1 << m;
    int bit0 = 0, bit1 = 0, bit2 = 0, bit3 = 0, bit4 = 0, bit5 = 0;
    if (oh & 0xA4) bit0 = (a >> 9) & 1;
    if (oh & 0x08) bit2 = (a >> 9) & 1;
    if (oh & 0x50) bit4 = (a >> 9) & 1;
    if (oh & 0x50) bit5 = (a >> 10) & 1;
    if (oh & 0xA0) bit1 = (a >> 10) & 1;
    if (oh & 0xC0) bit2 = (a >> 11) & 1;
    if (oh & 0x04) bit1 = (c >> 6) & 1;
    if (oh & 0xE8) bit3 = (c >> 6) & 1;
    if (oh & 0x20) bit2 = (c >> 7) & 1;
    if (oh & 0x5B) { bit0 = (b0 >> 6) & 1; bit1 = (b1 >> 6) & 1; }
    if (oh & 0x12) { bit2 = (b0 >> 7) & 1; bit3 = (b1 >> 7) & 1; }
    if (oh & 0xAF) { bit4 = (d0 >> 5) & 1; bit5 = (d1 >> 5) & 1; }
    if (oh & 0x05) { bit2 = (d0 >> 6) & 1; bit3 = (d1 >> 6) & 1; }
    v[0] = a & 0xFF;
    v[1] = (c & 0x3F) | (((a >> 8) & 1) << 6) | ((m & 1) << 7);
    v[2] = (b0 & 0x3F) | (bit0 << 6) | (((m >> 1) & 1) << 7);
    v[3] = (b1 & 0x3F) | (bit1 << 6) | (((m >> 2) & 1) << 7);
    v[4] = (d0 & 0x1F) | (bit4 << 5) | (bit2 << 6) | ((maj & 1) << 7);
    v[5] = (d1 & 0x1F) | (bit5 << 5) | (bit3 << 6) | (((maj >> 1) & 1) << 7);
    return true;
}

// Snap the six values to colour level `cl`, decode them, and return the squared distance of the decoded end points
// from the targets (12-bit units); 3e38 if the candidate does not exist.
CFX_HD float hdr_rgb_try(const Ctx& ctx, uint32_t cl, int m, const float* t0, const float* t1, int* vq, int* e0, int* e1)
{
    int v[6];
    if (!hdr_rgb_candidate(m, t0, t1, v)) return 3.0e38f;
#pragma unroll
    for (int k = 0; k < 6; ++k) vq[k] = static_cast<int>(quant_color(ctx, cl, static_cast<float>(v[k])));
    hdr_rgb_decode(vq, e0, e1);
    float err = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float da = static_cast<float>(e0[k]) - t0[k], db = static_cast<float>(e1[k]) - t1[k];
        err += da*da + db*db;
    }
    return err;
}

// Void-extent block with FP16 colour (constant HDR blocks): h = four half bit patterns.
CFX_HD uint4 pack_void_extent_hdr(uint32_t hr, uint32_t hg, uint32_t hb, uint32_t ha)
{
    return make_uint4(0xFFFFFFFCu, 0xFFFFFFFFu, hr | (hg << 16), hb | (ha << 16));
}

} // namespace astc
} // namespace cfx
