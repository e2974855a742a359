// TEST INFRASTRUCTURE ONLY -- the CPU oracle for the block-encode hot path.
//
// Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.
//
// What this file is: a restatement, in our own words, of the ~150 lines of Cuttlefish
// glue that sit between Texture::convert() and the vendored block encoders
// (gather one block with edge clamp, quantise, map Quality -> encoder knobs, call the
// encoder, store 8/16 bytes row-major).  The block encoders themselves are NOT restated:
// oracle/Makefile compiles them from the sources where they lie under /root/reference
// (rgbcx, bc7enc, libsquish, Compressonator cmp_core bc4/5/6, etc2comp, astcenc 5.3.0 --
// the CUTTLEFISH_ISPC=0 configuration) into oracle/_ref/libcfref.so next to this file.
// So the oracle's arithmetic IS the reference's arithmetic; parity is pinned by running
// the reference's own encoders (the reference's tests pin nothing but output size).
//
// Reference lines followed (all relative to /root/reference):
//   job loop / thread pool ........ lib/src/Converter.cpp:538-583
//   4x4 gather + clamp ............ lib/src/S3tcConverter.cpp:242-255
//   float -> u8 ................... lib/src/S3tcConverter.cpp:97-111
//   quality -> rgbcx level ........ lib/src/S3tcConverter.cpp:66-71
//   quality -> BC4/5 radius ....... lib/src/S3tcConverter.cpp:80-95
//   quality -> cmp quality ........ lib/src/S3tcConverter.cpp:73-78
//   BC1/BC1A/BC2/BC3/BC4/BC5 ...... lib/src/S3tcConverter.cpp:263-490
//   BC6H (Compressonator) ......... lib/src/S3tcConverter.cpp:492-591, HalfFloat.h:96-134
//   BC7 (bc7enc) .................. lib/src/S3tcConverter.cpp:170-227,593-646
//   ETC/EAC ....................... lib/src/EtcConverter.cpp:30-152
//   ASTC .......................... lib/src/AstcConverter.cpp:134-230
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "rgbcx.h"
#include "bc7enc.h"
#include "bc7decomp.h"
#include "squish.h"
#include "cmp_core.h"
#include "astcenc.h"
#include "Etc.h"
#include "EtcImage.h"
#include "EtcBlock4x4.h"
#include "EtcBlock4x4EncodingBits.h"

extern "C" {

// Numeric values mirror cuttlefish::Texture::Format (lib/include/cuttlefish/Texture.h:59-130).
enum {
    F_BC1_RGB = 29, F_BC1_RGBA, F_BC2, F_BC3, F_BC4, F_BC5, F_BC6H, F_BC7,
    F_ETC1, F_ETC2_R8G8B8, F_ETC2_R8G8B8A1, F_ETC2_R8G8B8A8, F_EAC_R11, F_EAC_R11G11,
    F_ASTC_4x4, F_ASTC_5x4, F_ASTC_5x5, F_ASTC_6x5, F_ASTC_6x6, F_ASTC_8x5, F_ASTC_8x6,
    F_ASTC_8x8, F_ASTC_10x5, F_ASTC_10x6, F_ASTC_10x8, F_ASTC_10x10, F_ASTC_12x10, F_ASTC_12x12
};
enum { T_UNORM = 0, T_SNORM = 1, T_UINT = 2, T_INT = 3, T_UFLOAT = 4, T_FLOAT = 5 };
enum { A_NONE = 0, A_STANDARD = 1, A_PREMULT = 2, A_ENCODED = 3 };

struct cfref_desc {
    uint32_t format, type, quality, alpha_type, color_mask, color_space;
    uint32_t width, height;
};

} // extern "C"

namespace {

struct RGBAf { float r, g, b, a; };

const unsigned astcDims[14][2] = {{4,4},{5,4},{5,5},{6,5},{6,6},{8,5},{8,6},{8,8},{10,5},{10,6},
    {10,8},{10,10},{12,10},{12,12}};

bool blockInfo(uint32_t format, unsigned& bw, unsigned& bh, unsigned& bytes)
{
    bw = bh = 4;
    switch (format) {
        case F_BC1_RGB: case F_BC1_RGBA: case F_BC4: case F_ETC1: case F_ETC2_R8G8B8:
        case F_ETC2_R8G8B8A1: case F_EAC_R11:
            bytes = 8; return true;
        case F_BC2: case F_BC3: case F_BC5: case F_BC6H: case F_BC7: case F_ETC2_R8G8B8A8:
        case F_EAC_R11G11:
            bytes = 16; return true;
        default:
            if (format >= F_ASTC_4x4 && format <= F_ASTC_12x12) {
                bw = astcDims[format - F_ASTC_4x4][0];
                bh = astcDims[format - F_ASTC_4x4][1];
                bytes = 16;
                return true;
            }
            return false;
    }
}

inline float clampf(float v, float lo, float hi) { return std::min(std::max(v, lo), hi); }

inline uint8_t toU8(float v) { return static_cast<uint8_t>(std::round(clampf(v, 0.0f, 1.0f)*0xFF)); }

// f32 -> f16, round to nearest even, overflow to inf (== _mm_cvtps_ph(x, 0)).
uint16_t toHalf(float f)
{
    uint32_t x; std::memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t absx = x & 0x7FFFFFFFu;
    if (absx >= 0x7F800000u)
        return static_cast<uint16_t>(sign | 0x7C00u | ((absx > 0x7F800000u) ? (0x200u | ((absx >> 13) & 0x3FFu)) : 0u));
    if (absx >= 0x477FF000u) // rounds to >= 65520 -> inf
        return static_cast<uint16_t>(sign | 0x7C00u);
    if (absx < 0x33000001u) // <= 2^-25 -> 0 (2^-25 exactly ties to even = 0)
        return static_cast<uint16_t>(sign);
    int e = static_cast<int>(absx >> 23) - 127;
    uint32_t m = (absx & 0x7FFFFFu) | 0x800000u;
    int shift;
    uint32_t base;
    if (e < -14) { shift = 13 + (-14 - e); base = 0; }
    else { shift = 13; base = static_cast<uint32_t>(e + 15) << 10; m &= 0x7FFFFFu; }
    uint32_t q = m >> shift;
    uint32_t rem = m & ((1u << shift) - 1u);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u)))
        ++q;
    return static_cast<uint16_t>(sign | (base + q));
}

struct Job {
    const cfref_desc* d;
    const float* src; // RGBAF, row 0 = top
    size_t pitch;     // floats per row
    uint8_t* dst;
    unsigned bw, bh, bytes, jobsX, jobsY;
};

inline const RGBAf* rowPtr(const Job& j, unsigned y)
{
    return reinterpret_cast<const RGBAf*>(j.src + static_cast<size_t>(y)*j.pitch);
}

// ---- BCn -----------------------------------------------------------------------------------

struct S3tcState {
    uint32_t level, radius;
    float cmpQuality;
    void* cmpOptions = nullptr;
    bc7enc_compress_block_params bc7;
    int squishFlags;
};

void initS3tc(const cfref_desc& d, S3tcState& s)
{
    static bool once = (rgbcx::init(), bc7enc_compress_block_init(), true);
    (void)once;
    s.level = rgbcx::MIN_LEVEL + (rgbcx::MAX_LEVEL - rgbcx::MIN_LEVEL)*d.quality/4u;
    static const uint32_t radii[5] = {3, 3, 5, 16, 32};
    s.radius = radii[d.quality];
    s.cmpQuality = static_cast<float>(d.quality)/4.0f;
    bool keepSign = d.type == T_SNORM || d.type == T_FLOAT;
    if (d.format == F_BC4 && keepSign) { CreateOptionsBC4(&s.cmpOptions); SetQualityBC4(s.cmpOptions, s.cmpQuality); }
    if (d.format == F_BC5 && keepSign) { CreateOptionsBC5(&s.cmpOptions); SetQualityBC5(s.cmpOptions, s.cmpQuality); }
    if (d.format == F_BC6H) {
        CreateOptionsBC6(&s.cmpOptions);
        SetQualityBC6(s.cmpOptions, s.cmpQuality);
        SetSignedBC6(s.cmpOptions, keepSign);
    }
    s.squishFlags = squish::kDxt1;
    if (d.quality <= 1) s.squishFlags |= squish::kColourRangeFit;
    else if (d.quality == 4) s.squishFlags |= squish::kColourIterativeClusterFit;

    bc7enc_compress_block_params_init(&s.bc7);
    bool srgb = d.color_space == 1;
    switch (d.quality) {
        case 0: s.bc7.m_max_partitions = 0; s.bc7.m_uber_level = 0; s.bc7.m_try_least_squares = false;
            s.bc7.m_mode17_partition_estimation_filterbank = true;
            bc7enc_compress_block_params_init_linear_weights(&s.bc7); break;
        case 1: s.bc7.m_max_partitions = 16; s.bc7.m_uber_level = 0; s.bc7.m_try_least_squares = true;
            s.bc7.m_mode17_partition_estimation_filterbank = true;
            bc7enc_compress_block_params_init_linear_weights(&s.bc7); break;
        case 2: s.bc7.m_max_partitions = BC7ENC_MAX_PARTITIONS; s.bc7.m_uber_level = 1;
            s.bc7.m_try_least_squares = true; s.bc7.m_mode17_partition_estimation_filterbank = false;
            if (srgb) bc7enc_compress_block_params_init_perceptual_weights(&s.bc7);
            else bc7enc_compress_block_params_init_linear_weights(&s.bc7);
            break;
        default: s.bc7.m_max_partitions = BC7ENC_MAX_PARTITIONS; s.bc7.m_uber_level = 4;
            s.bc7.m_try_least_squares = true; s.bc7.m_mode17_partition_estimation_filterbank = false;
            if (srgb) bc7enc_compress_block_params_init_perceptual_weights(&s.bc7);
            else bc7enc_compress_block_params_init_linear_weights(&s.bc7);
            break;
    }
    for (int c = 0; c < 4; ++c)
        if (!(d.color_mask & (1u << c))) s.bc7.m_weights[c] = 0;
}

void freeS3tc(const cfref_desc& d, S3tcState& s)
{
    if (!s.cmpOptions) return;
    if (d.format == F_BC4) DestroyOptionsBC4(s.cmpOptions);
    else if (d.format == F_BC5) DestroyOptionsBC5(s.cmpOptions);
    else DestroyOptionsBC6(s.cmpOptions);
    s.cmpOptions = nullptr;
}

void s3tcBlock(const Job& j, const S3tcState& s, unsigned x, unsigned y)
{
    const cfref_desc& d = *j.d;
    RGBAf px[16];
    for (unsigned r = 0; r < 4; ++r) {
        const RGBAf* row = rowPtr(j, std::min(y*4 + r, d.height - 1));
        for (unsigned c = 0; c < 4; ++c)
            px[r*4 + c] = row[std::min(x*4 + c, d.width - 1)];
    }
    uint8_t* out = j.dst + (static_cast<size_t>(y)*j.jobsX + x)*j.bytes;
    bool keepSign = d.type == T_SNORM || d.type == T_FLOAT;
    uint8_t u8[16][4];
    if (d.format != F_BC6H && !((d.format == F_BC4 || d.format == F_BC5) && keepSign))
        for (int i = 0; i < 16; ++i) {
            u8[i][0] = toU8(px[i].r); u8[i][1] = toU8(px[i].g);
            u8[i][2] = toU8(px[i].b); u8[i][3] = toU8(px[i].a);
        }

    switch (d.format) {
        case F_BC1_RGB:
            rgbcx::encode_bc1(s.level, out, &u8[0][0], true, true, nullptr);
            break;
        case F_BC1_RGBA: {
            bool hasAlpha = false;
            for (int i = 0; i < 16; ++i) if (px[i].a < 0.5f) hasAlpha = true;
            if (hasAlpha) {
                bool srgb = d.color_space == 1;
                static const float lum[3] = {0.2126f, 0.7152f, 0.0722f};
                float w[3];
                for (int c = 0; c < 3; ++c)
                    w[c] = (d.color_mask & (1u << c)) ? (srgb ? lum[c] : 1.0f) : 0.0f;
                squish::Compress(&u8[0][0], out, s.squishFlags, w);
            } else
                rgbcx::encode_bc1(s.level, out, &u8[0][0], true, false, nullptr);
            break;
        }
        case F_BC2: {
            const float alphaScale = 15.0f/255.0f;
            for (int i = 0; i < 8; ++i) {
                uint8_t a0 = static_cast<uint8_t>(std::round(u8[i*2][3]*alphaScale));
                uint8_t a1 = static_cast<uint8_t>(std::round(u8[i*2 + 1][3]*alphaScale));
                out[i] = static_cast<uint8_t>(a0 | (a1 << 4));
            }
            rgbcx::encode_bc1(s.level, out + 8, &u8[0][0], false, false, nullptr);
            break;
        }
        case F_BC3:
            if (d.quality <= 1) rgbcx::encode_bc3(s.level, out, &u8[0][0]);
            else rgbcx::encode_bc3_hq(s.level, out, &u8[0][0], s.radius);
            break;
        case F_BC4:
            if (keepSign) {
                uint8_t v[16];
                for (int i = 0; i < 16; ++i)
                    v[i] = static_cast<uint8_t>(static_cast<int8_t>(std::round(clampf(px[i].r, -1.0f, 1.0f)*0x7F)));
                CompressBlockBC4S(reinterpret_cast<const char*>(v), 4, out, s.cmpOptions);
            } else {
                uint8_t v[16];
                for (int i = 0; i < 16; ++i) v[i] = u8[i][0];
                if (d.quality <= 1) rgbcx::encode_bc4(out, v, 1);
                else rgbcx::encode_bc4_hq(out, v, 1, s.radius);
            }
            break;
        case F_BC5:
            if (keepSign) {
                uint8_t v[2][16];
                for (int i = 0; i < 16; ++i) {
                    v[0][i] = static_cast<uint8_t>(static_cast<int8_t>(std::round(clampf(px[i].r, -1.0f, 1.0f)*0x7F)));
                    v[1][i] = static_cast<uint8_t>(static_cast<int8_t>(std::round(clampf(px[i].g, -1.0f, 1.0f)*0x7F)));
                }
                CompressBlockBC5S(reinterpret_cast<const char*>(v[0]), 4,
                    reinterpret_cast<const char*>(v[1]), 4, out, s.cmpOptions);
            } else {
                uint8_t v[16][2];
                for (int i = 0; i < 16; ++i) { v[i][0] = u8[i][0]; v[i][1] = u8[i][1]; }
                if (d.quality <= 1) rgbcx::encode_bc5(out, &v[0][0], 0, 1, 2);
                else rgbcx::encode_bc5_hq(out, &v[0][0], 0, 1, 2, s.radius);
            }
            break;
        case F_BC6H: {
            uint16_t h[16][3];
            for (int i = 0; i < 16; ++i) {
                h[i][0] = toHalf(px[i].r); h[i][1] = toHalf(px[i].g); h[i][2] = toHalf(px[i].b);
            }
            CompressBlockBC6(&h[0][0], 12, out, s.cmpOptions);
            break;
        }
        case F_BC7:
            bc7enc_compress_block(out, u8, &s.bc7);
            break;
    }
}

// ---- ETC -----------------------------------------------------------------------------------

struct EtcState { float effort; Etc::Image::Format format; Etc::ErrorMetric metric; };

void initEtc(const cfref_desc& d, EtcState& s)
{
    static const float efforts[5] = {0.0f, 20.0f, 40.0f, 70.0f, 100.0f};
    s.effort = efforts[d.quality];
    bool linear = d.color_space == 0;
    switch (d.format) {
        case F_ETC1: s.format = Etc::Image::Format::ETC1; s.metric = linear ? Etc::RGBX : Etc::REC709; break;
        case F_ETC2_R8G8B8: s.format = Etc::Image::Format::RGB8; s.metric = linear ? Etc::RGBX : Etc::REC709; break;
        case F_ETC2_R8G8B8A1: s.format = Etc::Image::Format::RGB8A1; s.metric = linear ? Etc::RGBA : Etc::REC709; break;
        case F_ETC2_R8G8B8A8: s.format = Etc::Image::Format::RGBA8; s.metric = linear ? Etc::RGBA : Etc::REC709; break;
        case F_EAC_R11: s.format = d.type == T_UNORM ? Etc::Image::Format::R11 : Etc::Image::Format::SIGNED_R11;
            s.metric = Etc::NUMERIC; break;
        default: s.format = d.type == T_UNORM ? Etc::Image::Format::RG11 : Etc::Image::Format::SIGNED_RG11;
            s.metric = Etc::NUMERIC; break;
    }
}

void etcBlock(const Job& j, const EtcState& s, unsigned x, unsigned y)
{
    const cfref_desc& d = *j.d;
    RGBAf px[16];
    unsigned limX = std::min((x + 1)*4, d.width), limY = std::min((y + 1)*4, d.height);
    unsigned n = 0;
    for (unsigned r = y*4; r < limY; ++r) {
        const RGBAf* row = rowPtr(j, r);
        for (unsigned c = x*4; c < limX; ++c) px[n++] = row[c];
    }
    if (s.format == Etc::Image::Format::SIGNED_R11 || s.format == Etc::Image::Format::SIGNED_RG11)
        for (unsigned i = 0; i < 16; ++i) { px[i].r = px[i].r*0.5f + 0.5f; px[i].g = px[i].g*0.5f + 0.5f; }
    Etc::Image img(reinterpret_cast<float*>(px), limX - x*4, limY - y*4, s.metric);
    img.Encode(s.format, s.metric, s.effort, 1, 1);
    std::memcpy(j.dst + (static_cast<size_t>(y)*j.jobsX + x)*j.bytes, img.GetEncodingBits(), j.bytes);
}

// ---- ASTC ----------------------------------------------------------------------------------

struct AstcState { astcenc_swizzle swz; astcenc_config cfg; };

bool initAstc(const cfref_desc& d, unsigned bw, unsigned bh, AstcState& s)
{
    bool mr = d.color_mask & 1, mg = d.color_mask & 2, mb = d.color_mask & 4, ma = d.color_mask & 8;
    s.swz.r = mr ? ASTCENC_SWZ_R : ASTCENC_SWZ_0;
    s.swz.g = mg ? ASTCENC_SWZ_G : ASTCENC_SWZ_0;
    s.swz.b = mb ? ASTCENC_SWZ_B : ASTCENC_SWZ_0;
    s.swz.a = ma ? (d.alpha_type == A_NONE ? ASTCENC_SWZ_1 : ASTCENC_SWZ_A) : ASTCENC_SWZ_0;
    astcenc_profile prof = ASTCENC_PRF_LDR;
    if (d.type == T_UFLOAT)
        prof = (d.alpha_type == A_NONE || d.alpha_type == A_PREMULT) ? ASTCENC_PRF_HDR_RGB_LDR_A : ASTCENC_PRF_HDR;
    unsigned flags = 0;
    if (d.alpha_type == A_STANDARD || d.alpha_type == A_PREMULT) flags |= ASTCENC_FLG_USE_ALPHA_WEIGHT;
    if (d.color_space == 1) flags |= ASTCENC_FLG_USE_PERCEPTUAL;
    static const float presets[5] = {ASTCENC_PRE_FASTEST, ASTCENC_PRE_FAST, ASTCENC_PRE_MEDIUM,
        ASTCENC_PRE_THOROUGH, ASTCENC_PRE_EXHAUSTIVE};
    return astcenc_config_init(prof, bw, bh, 1, presets[d.quality], flags, &s.cfg) == ASTCENC_SUCCESS;
}

void astcBlock(const Job& j, const AstcState& s, astcenc_context* ctx, unsigned x, unsigned y)
{
    const cfref_desc& d = *j.d;
    RGBAf px[144];
    void* rows[12];
    unsigned n = 0;
    for (unsigned r = 0; r < j.bh; ++r) {
        rows[r] = px + n;
        const RGBAf* row = rowPtr(j, std::min(y*j.bh + r, d.height - 1));
        for (unsigned c = 0; c < j.bw; ++c) px[n++] = row[std::min(x*j.bw + c, d.width - 1)];
    }
    astcenc_image img;
    img.dim_x = j.bw; img.dim_y = j.bh; img.dim_z = 1; img.data_type = ASTCENC_TYPE_F32; img.data = rows;
    astcenc_compress_image(ctx, &img, &s.swz, j.dst + (static_cast<size_t>(y)*j.jobsX + x)*16, 16, 0);
    astcenc_compress_reset(ctx);
}

template <typename PerThread>
void runJobs(unsigned jobs, unsigned threads, PerThread body)
{
    threads = std::max(1u, std::min(threads, jobs));
    std::atomic<unsigned> next(0);
    if (threads == 1) { body(next, jobs); return; }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; ++t) pool.emplace_back([&] { body(next, jobs); });
    for (auto& t : pool) t.join();
}

} // namespace

extern "C" {

size_t cfref_encoded_size(const cfref_desc* d)
{
    unsigned bw, bh, bytes;
    if (!blockInfo(d->format, bw, bh, bytes)) return 0;
    return static_cast<size_t>((d->width + bw - 1)/bw)*((d->height + bh - 1)/bh)*bytes;
}

// src: RGBAF texels, row 0 = top, pitch_floats floats between rows. threads 0 = all cores.
int cfref_encode(const cfref_desc* d, const float* src, size_t pitch_floats, uint8_t* dst, unsigned threads)
{
    Job j;
    j.d = d; j.src = src; j.pitch = pitch_floats; j.dst = dst;
    if (!blockInfo(d->format, j.bw, j.bh, j.bytes) || d->quality > 4 || !d->width || !d->height) return -1;
    j.jobsX = (d->width + j.bw - 1)/j.bw; j.jobsY = (d->height + j.bh - 1)/j.bh;
    unsigned jobs = j.jobsX*j.jobsY;
    if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());

    if (d->format <= F_BC7) {
        S3tcState s; initS3tc(*d, s);
        runJobs(jobs, threads, [&](std::atomic<unsigned>& next, unsigned n) {
            for (unsigned k; (k = next++) < n;) s3tcBlock(j, s, k % j.jobsX, k / j.jobsX);
        });
        freeS3tc(*d, s);
    } else if (d->format <= F_EAC_R11G11) {
        EtcState s; initEtc(*d, s);
        runJobs(jobs, threads, [&](std::atomic<unsigned>& next, unsigned n) {
            for (unsigned k; (k = next++) < n;) etcBlock(j, s, k % j.jobsX, k / j.jobsX);
        });
    } else {
        AstcState s;
        if (!initAstc(*d, j.bw, j.bh, s)) return -2;
        runJobs(jobs, threads, [&](std::atomic<unsigned>& next, unsigned n) {
            astcenc_context* ctx = nullptr;
            if (astcenc_context_alloc(&s.cfg, 1, &ctx) != ASTCENC_SUCCESS) return;
            for (unsigned k; (k = next++) < n;) astcBlock(j, s, ctx, k % j.jobsX, k / j.jobsX);
            astcenc_context_free(ctx);
        });
    }
    return 0;
}

// Decode packed blocks with the reference's in-tree decoders to RGBAF (row 0 = top, tight).
// LDR formats decode to u8/255; BC6H to float via half; ASTC via astcenc_decompress_image (F32).
static float halfToFloat(uint16_t h)
{
    uint32_t s = (h & 0x8000u) << 16, e = (h >> 10) & 31, m = h & 0x3FF, x;
    if (e == 0) {
        if (!m) x = s;
        else { int sh = 0; while (!(m & 0x400)) { m <<= 1; ++sh; } x = s | ((113 - sh) << 23) | ((m & 0x3FF) << 13); }
    } else if (e == 31) x = s | 0x7F800000u | (m << 13);
    else x = s | ((e + 112) << 23) | (m << 13);
    float f; std::memcpy(&f, &x, 4); return f;
}

int cfref_decode(const cfref_desc* d, const uint8_t* blocks, float* out)
{
    unsigned bw, bh, bytes;
    if (!blockInfo(d->format, bw, bh, bytes)) return -1;
    unsigned jx = (d->width + bw - 1)/bw, jy = (d->height + bh - 1)/bh;
    rgbcx::init();
    if (d->format >= F_ASTC_4x4) {
        AstcState s;
        cfref_desc dd = *d; dd.color_mask = 15;
        if (!initAstc(dd, bw, bh, s)) return -2;
        s.swz.r = ASTCENC_SWZ_R; s.swz.g = ASTCENC_SWZ_G; s.swz.b = ASTCENC_SWZ_B; s.swz.a = ASTCENC_SWZ_A;
        astcenc_context* ctx = nullptr;
        s.cfg.flags &= ~0u; // decode uses same config
        if (astcenc_context_alloc(&s.cfg, 1, &ctx) != ASTCENC_SUCCESS) return -3;
        astcenc_image img; img.dim_x = d->width; img.dim_y = d->height; img.dim_z = 1;
        img.data_type = ASTCENC_TYPE_F32; void* slice = out; img.data = &slice;
        astcenc_error e = astcenc_decompress_image(ctx, blocks, static_cast<size_t>(jx)*jy*16, &img, &s.swz, 0);
        astcenc_context_free(ctx);
        return e == ASTCENC_SUCCESS ? 0 : -4;
    }
    if (d->format >= F_ETC1) {
        Etc::Image::Format fmt; Etc::ErrorMetric metric = Etc::RGBA;
        EtcState s; cfref_desc dd = *d; dd.quality = 2; initEtc(dd, s); fmt = s.format;
        unsigned ew = jx*4, eh = jy*4;
        // The decode constructor dereferences its source image, so hand it an opaque black one.
        std::vector<float> dummy(static_cast<size_t>(ew)*eh*4, 0.0f);
        for (size_t i = 3; i < dummy.size(); i += 4) dummy[i] = 1.0f;
        Etc::Image srcImg(dummy.data(), ew, eh, metric);
        // ~Image() delete[]s the encoding bits it was given, so give it its own copy.
        unsigned char* owned = new unsigned char[static_cast<size_t>(jx)*jy*bytes];
        std::memcpy(owned, blocks, static_cast<size_t>(jx)*jy*bytes);
        Etc::Image img(fmt, ew, eh, owned, jx*jy*bytes, &srcImg, metric);
        Etc::Block4x4* blk = img.GetBlocks();
        for (unsigned by = 0; by < jy; ++by) for (unsigned bx = 0; bx < jx; ++bx) {
            Etc::Block4x4& b = blk[by*jx + bx];
            Etc::ColorFloatRGBA* dec = b.GetDecodedColors();
            float* alphas = b.GetDecodedAlphas();
            for (unsigned c = 0; c < 4; ++c) for (unsigned r = 0; r < 4; ++r) {
                unsigned X = bx*4 + c, Y = by*4 + r;
                if (X >= d->width || Y >= d->height) continue;
                float* o = out + (static_cast<size_t>(Y)*d->width + X)*4;
                o[0] = dec[c*4 + r].fR; o[1] = dec[c*4 + r].fG; o[2] = dec[c*4 + r].fB;
                o[3] = alphas ? alphas[c*4 + r] : 1.0f;
            }
        }
        return 0;
    }
    void* opt6 = nullptr;
    if (d->format == F_BC6H) { CreateOptionsBC6(&opt6); SetSignedBC6(opt6, d->type == T_FLOAT); }
    for (unsigned by = 0; by < jy; ++by) for (unsigned bx = 0; bx < jx; ++bx) {
        const uint8_t* b = blocks + (static_cast<size_t>(by)*jx + bx)*bytes;
        float px[16][4];
        uint8_t u[16][4];
        for (int i = 0; i < 16; ++i) { u[i][0] = u[i][1] = u[i][2] = 0; u[i][3] = 255; }
        bool ldr = true;
        switch (d->format) {
            case F_BC1_RGB: rgbcx::unpack_bc1(b, u, true); for (int i = 0; i < 16; ++i) u[i][3] = 255; break;
            case F_BC1_RGBA: rgbcx::unpack_bc1(b, u, true); break;
            case F_BC2:
                rgbcx::unpack_bc1(b + 8, u, false);
                for (int i = 0; i < 16; ++i) { unsigned a = (b[i >> 1] >> ((i & 1)*4)) & 15; u[i][3] = static_cast<uint8_t>(a*17); }
                break;
            case F_BC3: rgbcx::unpack_bc3(b, u); break;
            case F_BC4: rgbcx::unpack_bc4(b, &u[0][0], 4); break;
            case F_BC5: rgbcx::unpack_bc5(b, u, 0, 1, 4); break;
            case F_BC7: bc7decomp::unpack_bc7(b, reinterpret_cast<bc7decomp::color_rgba*>(u)); break;
            case F_BC6H: {
                uint16_t h[48];
                DecompressBlockBC6(b, h, opt6);
                for (int i = 0; i < 16; ++i) {
                    px[i][0] = halfToFloat(h[i*3]); px[i][1] = halfToFloat(h[i*3 + 1]);
                    px[i][2] = halfToFloat(h[i*3 + 2]); px[i][3] = 1.0f;
                }
                ldr = false;
                break;
            }
        }
        if (ldr) for (int i = 0; i < 16; ++i) for (int c = 0; c < 4; ++c) px[i][c] = u[i][c]/255.0f;
        for (unsigned r = 0; r < 4; ++r) for (unsigned c = 0; c < 4; ++c) {
            unsigned X = bx*4 + c, Y = by*4 + r;
            if (X >= d->width || Y >= d->height) continue;
            std::memcpy(out + (static_cast<size_t>(Y)*d->width + X)*4, px[r*4 + c], 16);
        }
    }
    if (opt6) DestroyOptionsBC6(opt6);
    return 0;
}

unsigned cfref_hardware_threads(void) { return std::max(1u, std::thread::hardware_concurrency()); }

} // extern "C"
