// Lane-local BC1 colour-block encoder: ONE LANE OWNS ONE BLOCK.  Used for BC1_RGB, BC1_RGBA
// (punch-through), and the colour halves of BC2 and BC3.
//
// PARITY STATUS: this is our own search (PCA -> 565 quantisation -> exact index search -> least
// squares -> +-1 end-point descent, in four-colour and three-colour modes), held to "RGB PSNR >=
// reference - 0.1 dB".  It is NOT yet bit-identical to rgbcx::encode_bc1 level 9
// (lib/bc7enc_rdo/rgbcx.cpp:1692-1811, :2263-2604), which Bc1Converter::compressBlock
// (lib/src/S3tcConverter.cpp:263-270) calls: byte parity needs rgbcx's trained total-ordering tables
// (lib/bc7enc_rdo/rgbcx_table4.h) and its exact float op order; see DESIGN.md.
// Compiles for the device and, through hostdev.h, for tools/emu_bc1.cpp.
#pragma once
#include "bc1_tables.cuh"
#include "hostdev.h"

namespace cfx {
namespace bc1 {

enum { kAllow3 = 1, kAllowBlack = 2, kPunchThrough = 4 };

CFX_HD uint32_t expand565(uint32_t c)   // 565 -> 0x00BBGGRR with bit replication
{
    const uint32_t r = (c >> 11) & 31u, g = (c >> 5) & 63u, b = c & 31u;
    return ((r << 3) | (r >> 2)) | (((g << 2) | (g >> 4)) << 8) | (((b << 3) | (b >> 2)) << 16);
}

CFX_HD uint32_t quant565(float r, float g, float b)
{
    const int qr = min(max(__float2int_rn(r*(31.0f/255.0f)), 0), 31);
    const int qg = min(max(__float2int_rn(g*(63.0f/255.0f)), 0), 63);
    const int qb = min(max(__float2int_rn(b*(31.0f/255.0f)), 0), 31);
    return static_cast<uint32_t>((qr << 11) | (qg << 5) | qb);
}

// palette of (c0, c1): entries ordered along the line, pal[0] = c0 ... pal[n-1] = c1
CFX_HD void palette(uint32_t c0, uint32_t c1, bool four, uint32_t pal[4])
{
    const uint32_t a = expand565(c0), b = expand565(c1);
    pal[0] = a;
    if (four) {
        uint32_t p1 = 0, p2 = 0;
#pragma unroll
        for (int s = 0; s < 24; s += 8) {
            const uint32_t x = (a >> s) & 0xFFu, y = (b >> s) & 0xFFu;
            p1 |= ((2u*x + y)/3u) << s;
            p2 |= ((x + 2u*y)/3u) << s;
        }
        pal[1] = p1; pal[2] = p2; pal[3] = b;
    } else {
        uint32_t p1 = 0;
#pragma unroll
        for (int s = 0; s < 24; s += 8) p1 |= ((((a >> s) & 0xFFu) + ((b >> s) & 0xFFu))/2u) << s;
        pal[1] = p1; pal[2] = b; pal[3] = 0;
    }
}

CFX_HD uint32_t rgb_sse(uint32_t a, uint32_t b)
{
    const uint32_t d = __vabsdiffu4(a & 0x00FFFFFFu, b & 0x00FFFFFFu);
    return __dp4a(d, d, 0u);
}

// Best palette position per texel; texels in `skip` take position 3 of a three-colour palette
// (black / transparent) at no cost when punch-through, or at their real cost when black.
// pos: 2 bits per texel.  Returns the SSE.
CFX_HD uint32_t assign(const uint32_t* px, uint32_t c0, uint32_t c1, bool four, uint32_t skip, bool skip_is_free,
    uint32_t& pos)
{
    uint32_t pal[4];
    palette(c0, c1, four, pal);
    const int n = four ? 4 : 3;
    uint32_t err = 0, p = 0;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        if ((skip >> i) & 1u) {
            p |= 3u << (2*i);
            if (!skip_is_free) err += rgb_sse(px[i], 0u);
            continue;
        }
        uint32_t be = rgb_sse(px[i], pal[0]), bk = 0;
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            if (k >= n) break;
            const uint32_t e = rgb_sse(px[i], pal[k]);
            if (e < be) { be = e; bk = k; }
        }
        err += be;
        p |= bk << (2*i);
    }
    pos = p;
    return err;
}

struct Fit { uint32_t c0, c1, pos, err; };

// Fit the end points of one mode to the texels not in `skip`.
CFX_HD void fit_mode(const uint32_t* px, bool four, uint32_t skip, bool skip_is_free, int descent_rounds, Fit& f)
{
    float n = 0.0f, m[3] = {0, 0, 0};
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        if ((skip >> i) & 1u) continue;
        n += 1.0f; m[0] += static_cast<float>(px[i] & 0xFF); m[1] += static_cast<float>((px[i] >> 8) & 0xFF);
        m[2] += static_cast<float>((px[i] >> 16) & 0xFF);
    }
    if (n == 0.0f) { f.c0 = f.c1 = 0; f.err = assign(px, 0, 0, four, skip, skip_is_free, f.pos); return; }
    const float inv = 1.0f/n;
    m[0] *= inv; m[1] *= inv; m[2] *= inv;
    float cv[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        if ((skip >> i) & 1u) continue;
        const float d0 = static_cast<float>(px[i] & 0xFF) - m[0], d1 = static_cast<float>((px[i] >> 8) & 0xFF) - m[1],
            d2 = static_cast<float>((px[i] >> 16) & 0xFF) - m[2];
        cv[0] += d0*d0; cv[1] += d0*d1; cv[2] += d0*d2; cv[3] += d1*d1; cv[4] += d1*d2; cv[5] += d2*d2;
    }
    float v[3] = {cv[0], cv[1], cv[2]};
    float best = cv[0];
    if (cv[3] > best) { best = cv[3]; v[0] = cv[1]; v[1] = cv[3]; v[2] = cv[4]; }
    if (cv[5] > best) { best = cv[5]; v[0] = cv[2]; v[1] = cv[4]; v[2] = cv[5]; }
    for (int it = 0; it < 5; ++it) {
        const float n2 = v[0]*v[0] + v[1]*v[1] + v[2]*v[2];
        const float s = n2 > 1e-20f ? rsqrtf(n2) : 0.0f;
        const float a0 = v[0]*s, a1 = v[1]*s, a2 = v[2]*s;
        v[0] = cv[0]*a0 + cv[1]*a1 + cv[2]*a2;
        v[1] = cv[1]*a0 + cv[3]*a1 + cv[4]*a2;
        v[2] = cv[2]*a0 + cv[4]*a1 + cv[5]*a2;
    }
    {
        const float n2 = v[0]*v[0] + v[1]*v[1] + v[2]*v[2];
        const float s = n2 > 1e-20f ? rsqrtf(n2) : 0.0f;
        v[0] *= s; v[1] *= s; v[2] *= s;
    }
    float tmin = 3.0e38f, tmax = -3.0e38f;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        if ((skip >> i) & 1u) continue;
        const float t = (static_cast<float>(px[i] & 0xFF) - m[0])*v[0] + (static_cast<float>((px[i] >> 8) & 0xFF) - m[1])*v[1] +
            (static_cast<float>((px[i] >> 16) & 0xFF) - m[2])*v[2];
        tmin = fminf(tmin, t); tmax = fmaxf(tmax, t);
    }
    f.c0 = quant565(m[0] + tmin*v[0], m[1] + tmin*v[1], m[2] + tmin*v[2]);
    f.c1 = quant565(m[0] + tmax*v[0], m[1] + tmax*v[1], m[2] + tmax*v[2]);
    f.err = assign(px, f.c0, f.c1, four, skip, skip_is_free, f.pos);

    // least squares for the assigned positions
    const float wdiv = four ? (1.0f/3.0f) : 0.5f;
    for (int round = 0; round < 3 && f.err; ++round) {
        float A = 0, B = 0, C = 0, P[3] = {0, 0, 0}, Q[3] = {0, 0, 0};
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            if ((skip >> i) & 1u) continue;
            const float w = static_cast<float>((f.pos >> (2*i)) & 3u)*wdiv, iw = 1.0f - w;
            const float r = static_cast<float>(px[i] & 0xFF), g = static_cast<float>((px[i] >> 8) & 0xFF),
                b = static_cast<float>((px[i] >> 16) & 0xFF);
            A += iw*iw; B += iw*w; C += w*w;
            P[0] += iw*r; P[1] += iw*g; P[2] += iw*b; Q[0] += w*r; Q[1] += w*g; Q[2] += w*b;
        }
        const float det = A*C - B*B;
        if (fabsf(det) < 1e-6f) break;
        const float id = 1.0f/det;
        Fit t;
        t.c0 = quant565((C*P[0] - B*Q[0])*id, (C*P[1] - B*Q[1])*id, (C*P[2] - B*Q[2])*id);
        t.c1 = quant565((A*Q[0] - B*P[0])*id, (A*Q[1] - B*P[1])*id, (A*Q[2] - B*P[2])*id);
        t.err = assign(px, t.c0, t.c1, four, skip, skip_is_free, t.pos);
        if (t.err < f.err) f = t; else break;
    }
    // coordinate descent: +-1 on each of the six 5/6-bit end-point components
    for (int round = 0; round < descent_rounds && f.err; ++round) {
        bool improved = false;
#pragma unroll 1
        for (int k = 0; k < 12; ++k) {
            const int comp = k >> 1, delta = (k & 1) ? 1 : -1;
            const int shift = comp % 3 == 0 ? 11 : (comp % 3 == 1 ? 5 : 0), maxv = comp % 3 == 1 ? 63 : 31;
            uint32_t c = comp < 3 ? f.c0 : f.c1;
            const int val = static_cast<int>((c >> shift) & static_cast<uint32_t>(maxv)) + delta;
            if (val < 0 || val > maxv) continue;
            c = (c & ~(static_cast<uint32_t>(maxv) << shift)) | (static_cast<uint32_t>(val) << shift);
            Fit t;
            t.c0 = comp < 3 ? c : f.c0; t.c1 = comp < 3 ? f.c1 : c;
            t.err = assign(px, t.c0, t.c1, four, skip, skip_is_free, t.pos);
            if (t.err < f.err) { f = t; improved = true; }
        }
        if (!improved) break;
    }
}

// Packs a fit into the 8 colour bytes, honouring the c0 > c1 (four colours) / c0 <= c1 (three) rule.
CFX_HD uint2 pack(const Fit& f, bool four)
{
    uint32_t c0 = f.c0, c1 = f.c1, sel = 0;
    if (four) {
        // positions 0..3 along c0->c1 are selectors 0,2,3,1
        bool swap = c0 < c1;
        if (c0 == c1) {
            // degenerate: only selector 0 is safe (c0 <= c1 decodes as three colours)
            return make_uint2(c0 | (c1 << 16), 0u);
        }
        if (swap) { const uint32_t t = c0; c0 = c1; c1 = t; }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            uint32_t p = (f.pos >> (2*i)) & 3u;
            if (swap) p = 3u - p;
            const uint32_t s = p == 0 ? 0u : (p == 3 ? 1u : (p == 1 ? 2u : 3u));
            sel |= s << (2*i);
        }
    } else {
        // positions 0,1,2 along c0->c1 are selectors 0,2,1; position 3 = selector 3 (black / transparent)
        const bool swap = c0 > c1;
        if (swap) { const uint32_t t = c0; c0 = c1; c1 = t; }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            uint32_t p = (f.pos >> (2*i)) & 3u;
            uint32_t s;
            if (p == 3) s = 3;
            else { if (swap) p = 2u - p; s = p == 0 ? 0u : (p == 2 ? 1u : 2u); }
            sel |= s << (2*i);
        }
    }
    return make_uint2(c0 | (c1 << 16), sel);
}

// px: 16 RGBA8 texels.  flags: kAllow3 | kAllowBlack | kPunchThrough.  Returns the 8 bytes.
CFX_HD uint2 encode_color_block(const uint32_t* px, uint32_t flags, int descent_rounds)
{
    uint32_t transparent = 0, black = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if ((flags & kPunchThrough) && (px[i] >> 24) < 128u) transparent |= 1u << i;
        if (((px[i] | (px[i] >> 8) | (px[i] >> 16)) & 0xFFu) < 4u) black |= 1u << i;
    }
    if (transparent) {
        Fit f;
        fit_mode(px, false, transparent, true, descent_rounds, f);
        return pack(f, false);
    }
    bool solid = true;
#pragma unroll
    for (int i = 1; i < 16; ++i) solid = solid && ((px[i] ^ px[0]) & 0x00FFFFFFu) == 0;
    if (solid) {
        // every texel at the two-thirds point of the pair that reproduces the colour best
        const uint32_t r = kSolid5[px[0] & 0xFFu], g = kSolid6[(px[0] >> 8) & 0xFFu], b = kSolid5[(px[0] >> 16) & 0xFFu];
        Fit f;
        f.c0 = ((r & 0xFFu) << 11) | ((g & 0xFFu) << 5) | (b & 0xFFu);
        f.c1 = ((r >> 8) << 11) | ((g >> 8) << 5) | (b >> 8);
        f.pos = 0x55555555u; f.err = 0;
        return pack(f, true);
    }
    Fit best;
    fit_mode(px, true, 0, true, descent_rounds, best);
    bool four = true;
    if (best.err) {
        // near-flat blocks: the pair that reproduces the AVERAGE colour best, all texels free to choose
        uint32_t sr = 8, sg = 8, sb = 8;
#pragma unroll
        for (int i = 0; i < 16; ++i) { sr += px[i] & 0xFFu; sg += (px[i] >> 8) & 0xFFu; sb += (px[i] >> 16) & 0xFFu; }
        const uint32_t r = kSolid5[sr >> 4], g = kSolid6[sg >> 4], b = kSolid5[sb >> 4];
        Fit f;
        f.c0 = ((r & 0xFFu) << 11) | ((g & 0xFFu) << 5) | (b & 0xFFu);
        f.c1 = ((r >> 8) << 11) | ((g >> 8) << 5) | (b >> 8);
        f.err = assign(px, f.c0, f.c1, true, 0, true, f.pos);
        if (f.err < best.err) best = f;
    }
    if (best.err && (flags & kAllow3)) {
        Fit f;
        fit_mode(px, false, 0, true, descent_rounds, f);
        if (f.err < best.err) { best = f; four = false; }
    }
    if (best.err && (flags & kAllowBlack) && black) {
        Fit f;
        fit_mode(px, false, black, false, descent_rounds, f);
        if (f.err < best.err) { best = f; four = false; }
    }
    return pack(best, four);
}

// SSE of a packed colour block against the texels (decoder of palette(): thirds and halves by integer division).
// always4: BC2 / BC3 colour halves, which decode in four-colour mode whatever the order of c0 and c1.
CFX_HD uint32_t block_sse(const uint32_t* px, uint2 blk, bool always4)
{
    const uint32_t c0 = blk.x & 0xFFFFu, c1 = blk.x >> 16;
    uint32_t lin[4];
    const bool four = always4 || c0 > c1;
    palette(c0, c1, four, lin);
    // selector order: four colours c0, c1, 2/3, 1/3; three colours c0, c1, 1/2, black
    const uint32_t col[4] = {lin[0], four ? lin[3] : lin[2], lin[1], four ? lin[2] : 0u};
    uint32_t sse = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t k = (blk.y >> (2*i)) & 3u;
        sse += rgb_sse(px[i], k == 0 ? col[0] : (k == 1 ? col[1] : (k == 2 ? col[2] : col[3])));
    }
    return sse;
}

} // namespace bc1
} // namespace cfx
