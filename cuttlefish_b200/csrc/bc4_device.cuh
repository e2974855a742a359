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
__device__ __forceinline__ void bc4_trial_endpoints(uint32_t t, uint32_t n, uint32_t inv, int rad, uint32_t mn,
    uint32_t mx, uint32_t& e0, uint32_t& e1, bool& valid)
{
    uint32_t nn = n*n;
    uint32_t mode = t >= nn ? 1u : 0u;
    uint32_t rem = t - mode*nn;
    uint32_t lo_i = (rem*inv) >> 20;                                   // rem / n, inv = ceil(2^20 / n)
    int lo_d = static_cast<int>(lo_i) - rad;
    int hi_d = static_cast<int>(rem - lo_i*n) - rad;
    e0 = static_cast<uint32_t>(min(max(static_cast<int>(mx) + hi_d, SIGNED ? 1 : 0), 255));
    e1 = static_cast<uint32_t>(min(max(static_cast<int>(mn) + lo_d, SIGNED ? 1 : 0), 255));
    valid = e0 != e1;
    bool alpha6 = e0 <= e1;
    if ((mode == 0) ? alpha6 : !alpha6) { uint32_t tmp = e0; e0 = e1; e1 = tmp; }
}

// Words of shared memory bc4_encode_warp needs per warp for its threshold table (hq path).
constexpr uint32_t kBc4TableWords = 512;

// Ascending palette of a trial: q[0..7].  8-value mode: E_lo, the six sevenths, E_hi; 6-value mode: the low constant,
// E_lo, the four fifths, E_hi, 255 (bc4_block::get_block_values, rgbcx.h:391-423, sorted).  The floor divisions by 7
// and 5 are multiplications by ceil(2^16/7) and ceil(2^16/5) (exact for numerators <= 1785).
template <bool SIGNED>
__device__ __forceinline__ void bc4_sorted_palette(uint32_t mode, uint32_t elo, uint32_t ehi, uint32_t (&q)[8])
{
    const uint32_t diff = ehi - elo;
    if (mode == 0) {
        const uint32_t dm = diff*9363u, bm = elo*(7u*9363u);
        q[0] = elo; q[7] = ehi;
#pragma unroll
        for (uint32_t j = 1; j <= 6; ++j) q[j] = (j*dm + bm) >> 16;
    } else {
        const uint32_t dm = diff*13108u, bm = elo*(5u*13108u);
        q[0] = SIGNED ? 1u : 0u; q[1] = elo; q[6] = ehi; q[7] = 255u;
#pragma unroll
        for (uint32_t j = 1; j <= 4; ++j) q[1 + j] = (j*dm + bm) >> 16;
    }
}

// s_blk: RGBA8 texels of the block in shared memory, texel (row r, column c) at s_blk[r*row_stride + c] (block-major
// tiles: row_stride 4; row-major tiles as TMA writes them: row_stride = texels per tile row); chan: byte lane of the
// channel; s_tab: kBc4TableWords words of shared memory owned by this warp.  Returns the 8 block bytes as (lo, hi)
// words; identical in every lane.
//
// hq path.  The reference evaluates every trial's SSE = sum_i min_j (pal_j - v_i)^2 texel by texel (16 x 8 distances).
// The same integer falls out of the SORTED palette q_0 <= ... <= q_7 and two prefix functions of the block,
// N(x) = #{v_i <= x} and D(x) = 2 sum_{v_i <= x} v_i: a texel belongs to q_c when it lies in (tau_{c-1}, tau_c] with
// tau_c = floor((q_c + q_{c+1})/2) (ties are equidistant, so they do not change the SSE), and summing
// n_c q_c^2 - 2 q_c S1_c over the clusters by parts gives
//     SSE - sum v^2 = sum_{c<7} (q_c - q_{c+1}) (N(tau_c) (q_c + q_{c+1}) - D(tau_c)) + 16 q_7^2 - D(255) q_7,
// i.e. 7 table look-ups per trial.  The table (indexed by s = q_c + q_{c+1}, 512 entries holding the inner term
// N(s >> 1) s - D(s >> 1) itself) is built once per block with 16 shared-memory atomics and a warp scan.  Trials keep the reference's order (mode, lo_delta, hi_delta)
// and the lexicographic (SSE, trial) minimum is the serial loop's first minimum; the selectors are then computed for
// the winner only, with the reference's first-smallest-index tie rule.
// ceil(2^20 / n) for n = 2 radius + 1: r / n == (r*inv) >> 20 for r < n*n <= 4225.  One division per kernel, not per block.
__device__ __forceinline__ uint32_t bc4_trial_inv(uint32_t radius)
{
    const uint32_t n = 2*radius + 1;
    return ((1u << 20) + n - 1)/n;
}

template <bool SIGNED = false>
__device__ __forceinline__ uint2 bc4_encode_warp(const uint32_t* s_blk, uint32_t chan, uint32_t radius, uint32_t inv,
    bool hq, uint32_t* s_tab, uint32_t row_stride = 4)
{
    auto texel = [&](uint32_t i) -> uint32_t { return s_blk[(i >> 2)*row_stride + (i & 3u)]; };
    // end point bytes as stored: SNORM blocks hold two's complement s = b - 128
    auto stored = [](uint32_t e) -> uint32_t { return SIGNED ? ((e - 128u) & 0xFFu) : e; };
    const uint32_t lane = lane_id();
    const uint32_t shift = chan*8;
    const uint32_t v = (texel(lane & 15u) >> shift) & 0xFFu;       // lanes 16..31 mirror 0..15
    const uint32_t mn = __reduce_min_sync(0xFFFFFFFFu, v), mx = __reduce_max_sync(0xFFFFFFFFu, v);

    if (!hq && !SIGNED) {
        // encode_bc4: endpoints max/min, threshold selector assignment.
        if (mx == mn) return make_uint2(mx | (mn << 8), 0u);
        int delta = static_cast<int>(mx - mn);
        int bias = 4 - static_cast<int>(mn)*14;
        uint32_t sel = 0;
        if (lane < 16) {
            int x = static_cast<int>(v)*14 + bias;
            int cnt = (x >= delta*13) + (x >= delta*11) + (x >= delta*9) + (x >= delta*7) +
                (x >= delta*5) + (x >= delta*3) + (x >= delta);
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

    // ---- threshold table: D(x) | N(x) << 16 for x = 0..255 by histogram + scan, expanded to N s - D for s = 0..511 ----
    uint32_t dtot;
    {
        uint4* z = reinterpret_cast<uint4*>(s_tab) + lane*2;           // histogram lives in the first 256 words
        z[0] = make_uint4(0, 0, 0, 0); z[1] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        if (lane < 16) atomicAdd(&s_tab[v], (1u << 16) | (2u*v));
        __syncwarp();
        uint4 a = z[0], b = z[1];
        a.y += a.x; a.z += a.y; a.w += a.z; b.x += a.w; b.y += b.x; b.z += b.y; b.w += b.z;
        uint32_t run = b.w;                                            // inclusive scan of the lanes' totals
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, run, d);
            if (lane >= static_cast<uint32_t>(d)) run += o;
        }
        const uint32_t base = run - b.w;
        dtot = __shfl_sync(0xFFFFFFFFu, run, 31) & 0xFFFFu;
        a.x += base; a.y += base; a.z += base; a.w += base; b.x += base; b.y += base; b.z += base; b.w += base;
        __syncwarp();                                                  // every lane has read its histogram words
        // entry s (= q_c + q_{c+1}, tau = s >> 1) holds the whole inner term N(tau) s - D(tau)
        uint4* o4 = reinterpret_cast<uint4*>(s_tab) + lane*4;
        const uint32_t s0 = lane*16u;
        auto pair = [](uint32_t w, uint32_t s, uint32_t& even, uint32_t& odd) {
            const uint32_t cnt = w >> 16;
            even = cnt*s - (w & 0xFFFFu); odd = even + cnt;
        };
        uint4 o;
        pair(a.x, s0,      o.x, o.y); pair(a.y, s0 + 2,  o.z, o.w); o4[0] = o;
        pair(a.z, s0 + 4,  o.x, o.y); pair(a.w, s0 + 6,  o.z, o.w); o4[1] = o;
        pair(b.x, s0 + 8,  o.x, o.y); pair(b.y, s0 + 10, o.z, o.w); o4[2] = o;
        pair(b.z, s0 + 12, o.x, o.y); pair(b.w, s0 + 14, o.z, o.w); o4[3] = o;
        __syncwarp();
    }

    const uint32_t n = 2*radius + 1, nn = n*n;
    const int lowest = SIGNED ? 1 : 0;
    // both modes of an end point pair in one pass; the lane's first minimum in the reference's order (mode, lo, hi)
    // is mode 0's unless mode 1 is strictly better
    int best_err = 0x7FFFFFFF, best_err1 = 0x7FFFFFFF; uint32_t best_t = 0xFFFFFFFFu, best_t1 = 0xFFFFFFFFu;
    for (uint32_t r = lane; r < nn; r += 32) {
        const uint32_t lo_i = (r*inv) >> 20, hi_i = r - lo_i*n;
        const int a = min(max(static_cast<int>(mx + hi_i) - static_cast<int>(radius), lowest), 255);
        const int b = min(max(static_cast<int>(mn + lo_i) - static_cast<int>(radius), lowest), 255);
        if (a == b) continue;
        const uint32_t elo = static_cast<uint32_t>(min(a, b)), ehi = static_cast<uint32_t>(max(a, b));
#pragma unroll
        for (uint32_t mode = 0; mode < 2; ++mode) {
            uint32_t q[8];
            bc4_sorted_palette<SIGNED>(mode, elo, ehi, q);
            int acc = static_cast<int>(q[7]*(16u*q[7] - dtot));
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                const int inner = static_cast<int>(s_tab[q[c] + q[c + 1]]);
                acc += (static_cast<int>(q[c]) - static_cast<int>(q[c + 1]))*inner;
            }
            if (mode == 0) { if (acc < best_err) { best_err = acc; best_t = r; } }
            else if (acc < best_err1) { best_err1 = acc; best_t1 = nn + r; }
        }
    }
    if (best_err1 < best_err) { best_err = best_err1; best_t = best_t1; }
    const int werr = __reduce_min_sync(0xFFFFFFFFu, best_err);
    const uint32_t wt = __reduce_min_sync(0xFFFFFFFFu, best_err == werr ? best_t : 0xFFFFFFFFu);

    uint32_t e0, e1; bool valid;
    bc4_trial_endpoints<SIGNED>(wt, n, inv, static_cast<int>(radius), mn, mx, e0, e1, valid);
    uint32_t lo4, hi4;
    bc4_palette<SIGNED>(e0, e1, lo4, hi4);
    uint32_t sel = 0;
    if (lane < 16) {
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
    __syncwarp();                                                      // the table is free for the next call
    return make_uint2(stored(e0) | (stored(e1) << 8) | (lo << 16), (lo >> 16) | (hi << 16));
}

} // namespace cfx
