// BC4 (one 8-bit channel, 8 bytes) block encoder, one warp per block.
//
// Results are bit-identical to the reference's CPU encoders:
//   hq path   == rgbcx::encode_bc4_hq(dst, px, stride, search_rad, BC4_USE_ALL_MODES)
//                lib/bc7enc_rdo/rgbcx.cpp:2730-2884, palette rgbcx.h:391-423
//   fast path == rgbcx::encode_bc4   lib/bc7enc_rdo/rgbcx.cpp:2608-2728
//
// The reference walks  mode{8-value,6-value} x lo_delta x hi_delta  serially and keeps the first
// trial with the strictly smallest SSE.  Here the (2r+1)^2*2 trials are spread over the 32 lanes,
// each lane keeps its first minimum, and a lexicographic (SSE, trial index) warp reduction picks
// the same winner the serial loop would have kept.  All arithmetic is integer.
#pragma once
#include "common.cuh"

namespace cfx {

// SNORM blocks (Bc4Converter / Bc5Converter with Type::SNorm -> Compressonator's CompressBlockBC4S / BC5S,
// lib/src/S3tcConverter.cpp:400-429, :453-490; lib/compressonator/cmp_core/shaders/bc4_encode_kernel.cpp:202-229) are
// searched in a BIASED domain b = s + 128 (s = round(clamp(v,-1,1)*127) in [-127,127] -> b in [1,255]): interpolation
// is linear, so the unsigned search applies unchanged; only the 6-value mode's two constants differ (-1.0 and +1.0,
// i.e. b = 1 and 255) and end points stay >= 1.  Our own search (PSNR parity with the reference, not byte parity).
template <bool SIGNED = false>
__device__ __forceinline__ void bc4_palette(uint32_t e0, uint32_t e1, uint32_t& lo4, uint32_t& hi4)
{
    // bc4_block::get_block_values: 8 values when e0 > e1, else 6 values + {0,255}.
    uint32_t v2, v3, v4, v5, v6, v7;
    if (e0 > e1) {
        v2 = (e0*6 + e1)/7; v3 = (e0*5 + e1*2)/7; v4 = (e0*4 + e1*3)/7;
        v5 = (e0*3 + e1*4)/7; v6 = (e0*2 + e1*5)/7; v7 = (e0 + e1*6)/7;
    } else {
        v2 = (e0*4 + e1)/5; v3 = (e0*3 + e1*2)/5; v4 = (e0*2 + e1*3)/5; v5 = (e0 + e1*4)/5;
        v6 = SIGNED ? 1 : 0; v7 = 255;
    }
    lo4 = e0 | (e1 << 8) | (v2 << 16) | (v3 << 24);
    hi4 = v4 | (v5 << 8) | (v6 << 16) | (v7 << 24);
}

template <bool SIGNED = false>
__device__ __forceinline__ void bc4_trial_endpoints(uint32_t t, uint32_t n, int rad, uint32_t mn,
    uint32_t mx, uint32_t& e0, uint32_t& e1, bool& valid)
{
    uint32_t nn = n*n;
    uint32_t mode = t >= nn ? 1u : 0u;
    uint32_t rem = t - mode*nn;
    int lo_d = static_cast<int>(rem / n) - rad;
    int hi_d = static_cast<int>(rem % n) - rad;
    e0 = static_cast<uint32_t>(min(max(static_cast<int>(mx) + hi_d, SIGNED ? 1 : 0), 255));
    e1 = static_cast<uint32_t>(min(max(static_cast<int>(mn) + lo_d, SIGNED ? 1 : 0), 255));
    valid = e0 != e1;
    bool alpha6 = e0 <= e1;
    if ((mode == 0) ? alpha6 : !alpha6) { uint32_t tmp = e0; e0 = e1; e1 = tmp; }
}

// s_blk: RGBA8 texels of the block in shared memory, texel (row r, column c) at s_blk[r*row_stride + c] (block-major
// tiles: row_stride 4; row-major tiles as TMA writes them: row_stride = texels per tile row); chan: byte lane of the
// channel.  Returns the 8 block bytes as (lo, hi) words; identical in every lane.
template <bool SIGNED = false>
__device__ __forceinline__ uint2 bc4_encode_warp(const uint32_t* s_blk, uint32_t chan, uint32_t radius,
    bool hq, uint32_t row_stride = 4)
{
    auto texel = [&](uint32_t i) -> uint32_t { return s_blk[(i >> 2)*row_stride + (i & 3u)]; };
    // end point bytes as stored: SNORM blocks hold two's complement s = b - 128
    auto stored = [](uint32_t e) -> uint32_t { return SIGNED ? ((e - 128u) & 0xFFu) : e; };
    const uint32_t lane = lane_id();
    const uint32_t shift = chan*8;
    uint32_t rep[16];
    uint32_t mn = 255, mx = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        uint32_t v = (texel(i) >> shift) & 0xFFu;
        mn = min(mn, v); mx = max(mx, v);
        rep[i] = v*0x01010101u;
    }

    if (!hq && !SIGNED) {
        // encode_bc4: endpoints max/min, threshold selector assignment.
        if (mx == mn) return make_uint2(mx | (mn << 8), 0u);
        int delta = static_cast<int>(mx - mn);
        int bias = 4 - static_cast<int>(mn)*14;
        uint32_t sel = 0;
        if (lane < 16) {
            int v = static_cast<int>((texel(lane) >> shift) & 0xFFu);
            v = v*14 + bias;
            int cnt = (v >= delta*13) + (v >= delta*11) + (v >= delta*9) + (v >= delta*7) +
                (v >= delta*5) + (v >= delta*3) + (v >= delta);
            // s_tran: {1,7,6,5,4,3,2,0}
            sel = (0x02345671u >> (cnt*4)) & 7u;
        }
        uint64_t bits = static_cast<uint64_t>(sel) << (3*(lane & 15));
        if (lane >= 16) bits = 0;
        uint32_t lo = __reduce_or_sync(0xFFFFFFFFu, static_cast<uint32_t>(bits));
        uint32_t hi = __reduce_or_sync(0xFFFFFFFFu, static_cast<uint32_t>(bits >> 32));
        return make_uint2(mx | (mn << 8) | (lo << 16), (lo >> 16) | (hi << 16));
    }

    if (mx == mn) return make_uint2(stored(mn) | (stored(mn) << 8), 0u);

    const uint32_t n = 2*radius + 1;
    const uint32_t total = 2*n*n;
    uint32_t best_err = 0xFFFFFFFFu, best_t = 0xFFFFFFFFu;
    for (uint32_t t = lane; t < total; t += 32) {
        uint32_t e0, e1; bool valid;
        bc4_trial_endpoints<SIGNED>(t, n, static_cast<int>(radius), mn, mx, e0, e1, valid);
        if (!valid) continue;
        uint32_t lo4, hi4;
        bc4_palette<SIGNED>(e0, e1, lo4, hi4);
        uint32_t err = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            uint32_t m = __vminu4(__vabsdiffu4(lo4, rep[i]), __vabsdiffu4(hi4, rep[i]));
            m = __vminu4(m, m >> 16);
            m = min(m & 0xFFu, (m >> 8) & 0xFFu);
            err += m*m;
        }
        if (err < best_err) { best_err = err; best_t = t; }
    }
    uint32_t werr = __reduce_min_sync(0xFFFFFFFFu, best_err);
    uint32_t wt = __reduce_min_sync(0xFFFFFFFFu, best_err == werr ? best_t : 0xFFFFFFFFu);

    uint32_t e0, e1; bool valid;
    bc4_trial_endpoints<SIGNED>(wt, n, static_cast<int>(radius), mn, mx, e0, e1, valid);
    uint32_t lo4, hi4;
    bc4_palette<SIGNED>(e0, e1, lo4, hi4);
    uint32_t sel = 0;
    if (lane < 16) {
        uint32_t v = (texel(lane) >> shift) & 0xFFu;
        uint32_t bestd = 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t pv = ((j < 4 ? lo4 : hi4) >> ((j & 3)*8)) & 0xFFu;
            int d = static_cast<int>(pv) - static_cast<int>(v);
            uint32_t dd = static_cast<uint32_t>(d*d);
            if (dd < bestd) { bestd = dd; sel = j; }
        }
    }
    uint64_t bits = static_cast<uint64_t>(sel) << (3*(lane & 15));
    if (lane >= 16) bits = 0;
    uint32_t lo = __reduce_or_sync(0xFFFFFFFFu, static_cast<uint32_t>(bits));
    uint32_t hi = __reduce_or_sync(0xFFFFFFFFu, static_cast<uint32_t>(bits >> 32));
    return make_uint2(stored(e0) | (stored(e1) << 8) | (lo << 16), (lo >> 16) | (hi << 16));
}

} // namespace cfx
