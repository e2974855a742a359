// ASTC format tables for one block footprint, generated on the host from the ASTC specification's
// formulas (block-mode layout, weight/colour unquantisation, BISE trit/quint packing, the partition
// hash, bilinear weight infill) and handed to the kernels as one flat blob + an offset header.
//
// These are FORMAT constants (what any ASTC decoder must agree on), written from the spec -- the
// reference builds the equivalent tables in astcenc's init_block_size_descriptor
// (lib/astc-encoder/Source/astcenc_block_sizes.cpp:1165), astcenc_quantization.cpp and
// astcenc_integer_sequence.cpp.  tools/check_astc_tables.py cross-checks them against those
// sources when /root/reference is mounted; parity tests decode our blocks with astcenc's decoder.
//
// Plain C++ (no CUDA): used by astc.cu's host side and by the host emulator tools/emu_astc.cpp.
#pragma once
#include <stdint.h>

#include "astc_mode_prior.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

namespace cfx {
namespace astc {

constexpr int kMaxTexels = 64;      // footprints up to 8x8 / 10x6
constexpr int kWeightLevels = 12;   // 2 3 4 5 6 8 10 12 16 20 24 32
constexpr int kColorLevels = 17;    // 6 8 10 12 16 20 24 32 40 48 64 80 96 128 160 192 256

// Candidates evaluated per slot of each type (1 / 2 / 3 subsets, dual plane) by Texture::Quality
// (Lowest .. Highest; AstcConverter maps them to astcenc's fastest .. exhaustive presets,
// lib/src/AstcConverter.cpp:174-195).
static const uint32_t kPlanCounts[5][4] = {{24, 16, 0, 16}, {40, 32, 16, 32}, {64, 64, 32, 64}, {128, 128, 64, 128},
    {4096, 4096, 4096, 4096}};

struct Quant { uint16_t n; uint8_t bits, trits, quints; };

static const Quant kWeightQuant[kWeightLevels] = {
    {2, 1, 0, 0}, {3, 0, 1, 0}, {4, 2, 0, 0}, {5, 0, 0, 1}, {6, 1, 1, 0}, {8, 3, 0, 0},
    {10, 1, 0, 1}, {12, 2, 1, 0}, {16, 4, 0, 0}, {20, 2, 0, 1}, {24, 3, 1, 0}, {32, 5, 0, 0}};
static const Quant kColorQuant[kColorLevels] = {
    {6, 1, 1, 0}, {8, 3, 0, 0}, {10, 1, 0, 1}, {12, 2, 1, 0}, {16, 4, 0, 0}, {20, 2, 0, 1},
    {24, 3, 1, 0}, {32, 5, 0, 0}, {40, 3, 0, 1}, {48, 4, 1, 0}, {64, 6, 0, 0}, {80, 4, 0, 1},
    {96, 5, 1, 0}, {128, 7, 0, 0}, {160, 5, 0, 1}, {192, 6, 1, 0}, {256, 8, 0, 0}};

inline int ise_bits(int n, const Quant& q)
{
    return n*q.bits + (q.trits ? (8*n + 4)/5 : 0) + (q.quints ? (7*n + 2)/3 : 0);
}

// ---- unquantisation (ASTC spec, "Endpoint Unquantization" / "Weight Unquantization") ----------
inline int unquant_color(int v, const Quant& q)
{
    if (!q.trits && !q.quints) {
        int r = v << (8 - q.bits);
        for (int s = q.bits; s < 8; s += q.bits) r |= v << (8 - q.bits) >> s;
        return r & 0xFF;
    }
    const int m = v & ((1 << q.bits) - 1), D = v >> q.bits;
    const int a = m & 1, b = (m >> 1) & 1, c = (m >> 2) & 1, d = (m >> 3) & 1, e = (m >> 4) & 1, f = (m >> 5) & 1;
    const int A = a ? 0x1FF : 0;
    int B = 0, C = 0;
    if (q.trits) {
        switch (q.bits) {
            case 1: C = 204; B = 0; break;
            case 2: C = 93; B = (b << 8) | (b << 4) | (b << 2) | (b << 1); break;
            case 3: C = 44; B = (c << 8) | (b << 7) | (c << 3) | (b << 2) | (c << 1) | b; break;
            case 4: C = 22; B = (d << 8) | (c << 7) | (b << 6) | (d << 2) | (c << 1) | b; break;
            case 5: C = 11; B = (e << 8) | (d << 7) | (c << 6) | (b << 5) | (e << 1) | d; break;
            default: C = 5; B = (f << 8) | (e << 7) | (d << 6) | (c << 5) | (b << 4) | f; break;
        }
    } else {
        switch (q.bits) {
            case 1: C = 113; B = 0; break;
            case 2: C = 54; B = (b << 8) | (b << 3) | (b << 2); break;
            case 3: C = 26; B = (c << 8) | (b << 7) | (c << 2) | (b << 1) | c; break;
            case 4: C = 13; B = (d << 8) | (c << 7) | (b << 6) | (d << 1) | c; break;
            default: C = 6; B = (e << 8) | (d << 7) | (c << 6) | (b << 5) | e; break;
        }
    }
    int T = D*C + B;
    T ^= A;
    return (A & 0x80) | (T >> 2);
}

inline int unquant_weight(int v, const Quant& q)
{
    int r;
    if (!q.trits && !q.quints) {
        r = 0;
        // replicate to 6 bits
        int bits = q.bits, val = v << (6 - bits);
        r = val;
        for (int s = bits; s < 6; s += bits) r |= val >> s;
        r &= 63;
    } else if (q.bits == 0) {
        static const int t3[3] = {0, 32, 63}, t5[5] = {0, 16, 32, 47, 63};
        r = q.trits ? t3[v] : t5[v];
    } else {
        const int m = v & ((1 << q.bits) - 1), D = v >> q.bits;
        const int a = m & 1, b = (m >> 1) & 1, c = (m >> 2) & 1;
        const int A = a ? 0x7F : 0;
        int B = 0, C = 0;
        if (q.trits) {
            switch (q.bits) {
                case 1: C = 50; B = 0; break;
                case 2: C = 23; B = (b << 6) | (b << 2) | b; break;
                default: C = 11; B = (c << 6) | (b << 5) | (c << 1) | b; break;
            }
        } else {
            switch (q.bits) {
                case 1: C = 28; B = 0; break;
                default: C = 13; B = (b << 6) | (b << 1); break;
            }
        }
        int T = D*C + B;
        T ^= A;
        r = (A & 0x20) | (T >> 2);
    }
    if (r > 32) r += 1;
    return r;
}

// ---- BISE trit / quint block decoding (spec "Integer Sequence Encoding"), inverted for packing --
inline void decode_trits(int T, int t[5])
{
    int C;
    if (((T >> 2) & 7) == 7) { C = (((T >> 5) & 7) << 2) | (T & 3); t[4] = t[3] = 2; }
    else {
        C = T & 0x1F;
        if (((T >> 5) & 3) == 3) { t[4] = 2; t[3] = (T >> 7) & 1; }
        else { t[4] = (T >> 7) & 1; t[3] = (T >> 5) & 3; }
    }
    if ((C & 3) == 3) { t[2] = 2; t[1] = (C >> 4) & 1; t[0] = (((C >> 3) & 1) << 1) | (((C >> 2) & 1) & ~((C >> 3) & 1)); }
    else if (((C >> 2) & 3) == 3) { t[2] = 2; t[1] = 2; t[0] = C & 3; }
    else { t[2] = (C >> 4) & 1; t[1] = (C >> 2) & 3; t[0] = (((C >> 1) & 1) << 1) | ((C & 1) & ~((C >> 1) & 1)); }
}

inline void decode_quints(int Q, int q[3])
{
    if (((Q >> 1) & 3) == 3 && ((Q >> 5) & 3) == 0) {
        q[2] = ((Q & 1) << 2) | ((((Q >> 4) & 1) & ~(Q & 1)) << 1) | (((Q >> 3) & 1) & ~(Q & 1));
        q[1] = q[0] = 4;
    } else {
        int C;
        if (((Q >> 1) & 3) == 3) { q[2] = 4; C = (((Q >> 3) & 3) << 3) | ((~(Q >> 5) & 3) << 1) | (Q & 1); }
        else { q[2] = (Q >> 5) & 3; C = Q & 0x1F; }
        if ((C & 7) == 5) { q[1] = 4; q[0] = (C >> 3) & 3; }
        else { q[1] = (C >> 3) & 3; q[0] = C & 7; }
    }
}

// ---- partition hash (spec "Partition Pattern Generation") ------------------------------------
inline uint32_t hash52(uint32_t p)
{
    p ^= p >> 15; p -= p << 17; p += p << 7; p += p << 4; p ^= p >> 5; p += p << 16;
    p ^= p >> 7; p ^= p >> 3; p ^= p << 6; p ^= p >> 17;
    return p;
}

inline int select_partition(int seed, int x, int y, int z, int pc, bool small_block)
{
    if (small_block) { x <<= 1; y <<= 1; z <<= 1; }
    seed += (pc - 1)*1024;
    uint32_t rnum = hash52(static_cast<uint32_t>(seed));
    uint8_t s[12];
    s[0] = rnum & 0xF; s[1] = (rnum >> 4) & 0xF; s[2] = (rnum >> 8) & 0xF; s[3] = (rnum >> 12) & 0xF;
    s[4] = (rnum >> 16) & 0xF; s[5] = (rnum >> 20) & 0xF; s[6] = (rnum >> 24) & 0xF; s[7] = (rnum >> 28) & 0xF;
    s[8] = (rnum >> 18) & 0xF; s[9] = (rnum >> 22) & 0xF; s[10] = (rnum >> 26) & 0xF;
    s[11] = ((rnum >> 30) | (rnum << 2)) & 0xF;
    for (int i = 0; i < 12; ++i) s[i] = static_cast<uint8_t>(s[i]*s[i]);
    int sh1, sh2;
    if (seed & 1) { sh1 = (seed & 2) ? 4 : 5; sh2 = (pc == 3) ? 6 : 5; }
    else { sh1 = (pc == 3) ? 6 : 5; sh2 = (seed & 2) ? 4 : 5; }
    const int sh3 = (seed & 0x10) ? sh1 : sh2;
    s[0] >>= sh1; s[1] >>= sh2; s[2] >>= sh1; s[3] >>= sh2; s[4] >>= sh1; s[5] >>= sh2; s[6] >>= sh1; s[7] >>= sh2;
    s[8] >>= sh3; s[9] >>= sh3; s[10] >>= sh3; s[11] >>= sh3;
    int a = s[0]*x + s[1]*y + s[10]*z + (rnum >> 14);
    int b = s[2]*x + s[3]*y + s[11]*z + (rnum >> 10);
    int c = s[4]*x + s[5]*y + s[8]*z + (rnum >> 6);
    int d = s[6]*x + s[7]*y + s[9]*z + (rnum >> 2);
    a &= 0x3F; b &= 0x3F; c &= 0x3F; d &= 0x3F;
    if (pc < 4) d = 0;
    if (pc < 3) c = 0;
    if (a >= b && a >= c && a >= d) return 0;
    if (b >= c && b >= d) return 1;
    if (c >= d) return 2;
    return 3;
}

// ---- block modes (spec "Block Mode") ----------------------------------------------------------
// Decodes an 11-bit block mode; returns false for reserved / void-extent encodings.
inline bool decode_block_mode(int mode, int& W, int& H, int& level, int& dual)
{
    int R, Hp, D;
    const int b = mode;
    if ((b & 3) != 0) {
        R = ((b >> 4) & 1) | ((b & 3) << 1);
        const int A = (b >> 5) & 3, B = (b >> 7) & 3;
        D = (b >> 10) & 1; Hp = (b >> 9) & 1;
        switch ((b >> 2) & 3) {
            case 0: W = B + 4; H = A + 2; break;
            case 1: W = B + 8; H = A + 2; break;
            case 2: W = A + 2; H = B + 8; break;
            default:
                if (B & 2) { W = (B & 1) + 2; H = A + 2; }
                else { W = A + 2; H = (B & 1) + 6; }
                break;
        }
    } else {
        if (((b >> 2) & 3) == 0) return false;              // reserved
        R = ((b >> 4) & 1) | (((b >> 2) & 3) << 1);
        const int A = (b >> 5) & 3;
        D = (b >> 10) & 1; Hp = (b >> 9) & 1;
        switch ((b >> 7) & 3) {
            case 0: W = 12; H = A + 2; break;
            case 1: W = A + 2; H = 12; break;
            case 3:
                if ((b >> 5) & 2) return false;              // void-extent / reserved
                if ((b >> 5) & 1) { W = 10; H = 6; } else { W = 6; H = 10; }
                break;
            default:
                W = A + 6; H = ((b >> 9) & 3) + 6; D = 0; Hp = 0;
                break;
        }
    }
    if (R < 2) return false;
    level = (R - 2) + 6*Hp;
    dual = D;
    return true;
}

// ---- the blob ---------------------------------------------------------------------------------
struct GridInfo { uint8_t w, h, nw, pad; };
struct ModeInfo { uint16_t mode_bits; uint8_t grid, level, wbits, nw, dual, pad; };

// All offsets are bytes from the blob start.
struct AstcTab {
    uint32_t bw, bh, texels;
    uint32_t n_grids, n_modes1, n_modes2;      // single-plane modes first, dual-plane after them
    uint32_t off_grids, off_modes;
    uint32_t off_infill;        // [grid][texel] uint2 {4 x weight index, 4 x factor (sum 16)}
    uint32_t off_csr_start;     // [grid][kMaxTexels + 2] uint16
    uint32_t off_csr_ent;       // [grid][4*texels] uint16 {texel | factor << 8}
    uint32_t off_wnorm;         // [grid][kMaxTexels] float 1 / (sum of factors of that weight)
    uint32_t off_wq_val;        // [level][32] u8 sorted unquantised weight values (0..64)
    uint32_t off_wq_enc;        // [level][32] u8 encoded integer of rank k
    uint32_t off_cq_near;       // [level][256] u8 rank of the nearest representable value
    uint32_t off_cq_val;        // [level][256] u8 value of rank k
    uint32_t off_cq_enc;        // [level][256] u8 encoded integer of rank k
    uint32_t off_clevel;        // [n_ints/2 (0..9)][128] u8 colour level for `bits` available, 0xFF = none
    uint32_t off_trit_enc;      // [243] u8
    uint32_t off_quint_enc;     // [125] u8
    uint32_t off_part2;         // [1024] uint64 texel mask of subset 1 (0 = unusable seed)
    uint32_t off_part3;         // [1024][2] uint64 masks of subsets 1 and 2 (0,0 = unusable)
    uint32_t n_part2, n_part3;  // usable seeds
    uint32_t off_cand[4];       // per slot type (1 / 2 / 3 subsets, dual plane): uint16 mode indices, likeliest first
    uint32_t n_cand[4];
    uint32_t off_cand_q[5][4];  // per quality: the first kPlanCounts[q][type] of those, re-ordered so that modes
    uint32_t n_cand_q[5][4];    // sharing a weight grid are adjacent (one decimation per grid)
    uint32_t blob_bytes;
};

struct Built {
    AstcTab tab;
    std::vector<uint8_t> blob;
};

inline Built build_tables(int bw, int bh)
{
    Built out;
    AstcTab& t = out.tab;
    std::memset(&t, 0, sizeof(t));
    t.bw = bw; t.bh = bh; t.texels = bw*bh;
    const int T = bw*bh;
    std::vector<uint8_t>& blob = out.blob;
    auto reserve = [&](size_t bytes, size_t align) {
        size_t off = (blob.size() + align - 1)/align*align;
        blob.resize(off + bytes, 0);
        return static_cast<uint32_t>(off);
    };

    // grids and modes
    std::vector<GridInfo> grids;
    std::map<std::pair<int, int>, int> grid_index;
    std::map<std::tuple<int, int, int, int>, int> seen;
    std::vector<ModeInfo> modes[2];
    for (int m = 0; m < 2048; ++m) {
        int W, H, level, dual;
        if (!decode_block_mode(m, W, H, level, dual)) continue;
        if (W > bw || H > bh) continue;
        const int nw = W*H;
        if (nw*(1 + dual) > 64) continue;
        const int wbits = ise_bits(nw*(1 + dual), kWeightQuant[level]);
        if (wbits < 24 || wbits > 96) continue;
        if (seen.count(std::make_tuple(W, H, level, dual))) continue;
        seen[std::make_tuple(W, H, level, dual)] = m;
        auto key = std::make_pair(W, H);
        if (!grid_index.count(key)) {
            grid_index[key] = static_cast<int>(grids.size());
            grids.push_back(GridInfo{static_cast<uint8_t>(W), static_cast<uint8_t>(H), static_cast<uint8_t>(nw), 0});
        }
        ModeInfo mi;
        mi.mode_bits = static_cast<uint16_t>(m); mi.grid = static_cast<uint8_t>(grid_index[key]);
        mi.level = static_cast<uint8_t>(level); mi.wbits = static_cast<uint8_t>(wbits);
        mi.nw = static_cast<uint8_t>(nw); mi.dual = static_cast<uint8_t>(dual); mi.pad = 0;
        modes[dual].push_back(mi);
    }
    for (int d = 0; d < 2; ++d)
        std::stable_sort(modes[d].begin(), modes[d].end(), [](const ModeInfo& a, const ModeInfo& b) {
            if (a.nw != b.nw) return a.nw > b.nw;
            if (a.grid != b.grid) return a.grid < b.grid;
            return a.level > b.level;
        });
    t.n_grids = static_cast<uint32_t>(grids.size());
    t.n_modes1 = static_cast<uint32_t>(modes[0].size());
    t.n_modes2 = static_cast<uint32_t>(modes[1].size());
    t.off_grids = reserve(grids.size()*sizeof(GridInfo), 4);
    std::memcpy(&blob[t.off_grids], grids.data(), grids.size()*sizeof(GridInfo));
    t.off_modes = reserve((modes[0].size() + modes[1].size())*sizeof(ModeInfo), 8);
    std::memcpy(&blob[t.off_modes], modes[0].data(), modes[0].size()*sizeof(ModeInfo));
    std::memcpy(&blob[t.off_modes + modes[0].size()*sizeof(ModeInfo)], modes[1].data(), modes[1].size()*sizeof(ModeInfo));

    // infill (spec "Weight Infill") and its transpose
    const int G = static_cast<int>(grids.size());
    t.off_infill = reserve(static_cast<size_t>(G)*T*8, 8);
    t.off_csr_start = reserve(static_cast<size_t>(G)*(kMaxTexels + 2)*2, 4);
    t.off_csr_ent = reserve(static_cast<size_t>(G)*4*T*2, 4);
    t.off_wnorm = reserve(static_cast<size_t>(G)*kMaxTexels*4, 4);
    for (int g = 0; g < G; ++g) {
        const int W = grids[g].w, H = grids[g].h;
        std::vector<std::vector<std::pair<int, int>>> per_weight(W*H);
        for (int y = 0; y < bh; ++y)
            for (int x = 0; x < bw; ++x) {
                const int tex = y*bw + x;
                const int Ds = (1024 + bw/2)/(bw - 1), Dt = (1024 + bh/2)/(bh - 1);
                const int cs = Ds*x, ct = Dt*y;
                const int gs = (cs*(W - 1) + 32) >> 6, gt = (ct*(H - 1) + 32) >> 6;
                const int js = gs >> 4, fs = gs & 15, jt = gt >> 4, ft = gt & 15;
                const int w11 = (fs*ft + 8) >> 4;
                const int w10 = ft - w11, w01 = fs - w11, w00 = 16 - fs - ft + w11;
                const int v0 = js + jt*W;
                int idx[4] = {v0, v0 + 1, v0 + W, v0 + W + 1};
                int fac[4] = {w00, w01, w10, w11};
                uint8_t* dst = &blob[t.off_infill + (static_cast<size_t>(g)*T + tex)*8];
                for (int k = 0; k < 4; ++k) {
                    if (fac[k] == 0) idx[k] = v0;        // keep indices in range when the factor is 0
                    dst[k] = static_cast<uint8_t>(idx[k]);
                    dst[4 + k] = static_cast<uint8_t>(fac[k]);
                    if (fac[k]) per_weight[idx[k]].push_back(std::make_pair(tex, fac[k]));
                }
            }
        uint16_t* start = reinterpret_cast<uint16_t*>(&blob[t.off_csr_start + static_cast<size_t>(g)*(kMaxTexels + 2)*2]);
        uint16_t* ent = reinterpret_cast<uint16_t*>(&blob[t.off_csr_ent + static_cast<size_t>(g)*4*T*2]);
        float* norm = reinterpret_cast<float*>(&blob[t.off_wnorm + static_cast<size_t>(g)*kMaxTexels*4]);
        int pos = 0;
        for (int j = 0; j < W*H; ++j) {
            start[j] = static_cast<uint16_t>(pos);
            int sum = 0;
            for (auto& e : per_weight[j]) { ent[pos++] = static_cast<uint16_t>(e.first | (e.second << 8)); sum += e.second; }
            norm[j] = sum ? 1.0f/static_cast<float>(sum) : 0.0f;
        }
        start[W*H] = static_cast<uint16_t>(pos);
    }

    // weight quantisation tables
    t.off_wq_val = reserve(kWeightLevels*32, 4);
    t.off_wq_enc = reserve(kWeightLevels*32, 4);
    for (int l = 0; l < kWeightLevels; ++l) {
        std::vector<std::pair<int, int>> v;
        for (int e = 0; e < kWeightQuant[l].n; ++e) v.push_back(std::make_pair(unquant_weight(e, kWeightQuant[l]), e));
        std::sort(v.begin(), v.end());
        for (size_t k = 0; k < v.size(); ++k) {
            blob[t.off_wq_val + l*32 + k] = static_cast<uint8_t>(v[k].first);
            blob[t.off_wq_enc + l*32 + k] = static_cast<uint8_t>(v[k].second);
        }
    }
    // colour quantisation tables
    t.off_cq_near = reserve(kColorLevels*256, 4);
    t.off_cq_val = reserve(kColorLevels*256, 4);
    t.off_cq_enc = reserve(kColorLevels*256, 4);
    for (int l = 0; l < kColorLevels; ++l) {
        std::vector<std::pair<int, int>> v;
        for (int e = 0; e < kColorQuant[l].n; ++e) v.push_back(std::make_pair(unquant_color(e, kColorQuant[l]), e));
        std::sort(v.begin(), v.end());
        for (size_t k = 0; k < v.size(); ++k) {
            blob[t.off_cq_val + l*256 + k] = static_cast<uint8_t>(v[k].first);
            blob[t.off_cq_enc + l*256 + k] = static_cast<uint8_t>(v[k].second);
        }
        for (int x = 0; x < 256; ++x) {
            int best = 0;
            for (size_t k = 1; k < v.size(); ++k)
                if (std::abs(v[k].first - x) < std::abs(v[best].first - x)) best = static_cast<int>(k);
            blob[t.off_cq_near + l*256 + x] = static_cast<uint8_t>(best);
        }
    }
    // colour level for (number of endpoint integers, bits available)
    t.off_clevel = reserve(10*128, 4);
    for (int half = 0; half < 10; ++half)
        for (int bits = 0; bits < 128; ++bits) {
            int best = 0xFF;
            for (int l = 0; l < kColorLevels; ++l)
                if (half > 0 && ise_bits(half*2, kColorQuant[l]) <= bits) best = l;
            blob[t.off_clevel + half*128 + bits] = static_cast<uint8_t>(best);
        }
    // trit / quint packing
    t.off_trit_enc = reserve(243, 4);
    t.off_quint_enc = reserve(125, 4);
    for (int T8 = 255; T8 >= 0; --T8) {
        int tr[5];
        decode_trits(T8, tr);
        blob[t.off_trit_enc + tr[0] + 3*tr[1] + 9*tr[2] + 27*tr[3] + 81*tr[4]] = static_cast<uint8_t>(T8);
    }
    for (int Q7 = 127; Q7 >= 0; --Q7) {
        int qu[3];
        decode_quints(Q7, qu);
        if (qu[0] > 4 || qu[1] > 4 || qu[2] > 4) continue;
        blob[t.off_quint_enc + qu[0] + 5*qu[1] + 25*qu[2]] = static_cast<uint8_t>(Q7);
    }
    // partitions
    t.off_part2 = reserve(1024*8, 8);
    t.off_part3 = reserve(1024*16, 8);
    {
        const bool small_block = T < 31;
        std::vector<uint64_t> seen2;
        std::vector<std::pair<uint64_t, uint64_t>> seen3;
        uint64_t* p2 = reinterpret_cast<uint64_t*>(&blob[t.off_part2]);
        uint64_t* p3 = reinterpret_cast<uint64_t*>(&blob[t.off_part3]);
        const uint64_t full = T == 64 ? ~0ull : ((1ull << T) - 1);
        for (int seed = 0; seed < 1024; ++seed) {
            uint64_t m1 = 0;
            for (int i = 0; i < T; ++i)
                if (select_partition(seed, i % bw, i/bw, 0, 2, small_block) == 1) m1 |= 1ull << i;
            bool ok = m1 != 0 && m1 != full;
            for (uint64_t s : seen2) if (s == m1 || s == (m1 ^ full)) ok = false;
            if (ok) { seen2.push_back(m1); p2[seed] = m1; t.n_part2++; }
            uint64_t a = 0, b = 0;
            for (int i = 0; i < T; ++i) {
                int p = select_partition(seed, i % bw, i/bw, 0, 3, small_block);
                if (p == 1) a |= 1ull << i;
                if (p == 2) b |= 1ull << i;
            }
            const uint64_t c = full ^ a ^ b;
            ok = a != 0 && b != 0 && c != 0;
            uint64_t key[3] = {a, b, c};
            std::sort(key, key + 3);
            for (auto& s : seen3) if (s.first == key[0] && s.second == key[1]) ok = false;
            if (ok) { seen3.push_back(std::make_pair(key[0], key[1])); p3[seed*2] = a; p3[seed*2 + 1] = b; t.n_part3++; }
        }
    }
    // candidate order per slot type: the footprint's prior (astc_mode_prior.h) first, then the rest
    {
        const ModeInfo* all = reinterpret_cast<const ModeInfo*>(&blob[t.off_modes]);
        const GridInfo* gr = reinterpret_cast<const GridInfo*>(&blob[t.off_grids]);
        (void)gr;
        for (int type = 0; type < 4; ++type) {
            const uint32_t first = type == 3 ? t.n_modes1 : 0, count = type == 3 ? t.n_modes2 : t.n_modes1;
            std::vector<uint16_t> order;
            std::vector<bool> used(t.n_modes1 + t.n_modes2, false);
            for (const ModePrior* pr = kModePriors; pr->modes; ++pr) {
                if (pr->bw != bw || pr->bh != bh || pr->type != type) continue;
                for (const uint16_t* e = pr->modes; *e != 0xFFFF; ++e)
                    for (uint32_t i = first; i < first + count; ++i)
                        if (!used[i] && gr[all[i].grid].w == (*e & 15) && gr[all[i].grid].h == ((*e >> 4) & 15) &&
                            all[i].level == ((*e >> 8) & 15)) { used[i] = true; order.push_back(static_cast<uint16_t>(i)); }
            }
            for (uint32_t i = first; i < first + count; ++i) if (!used[i]) order.push_back(static_cast<uint16_t>(i));
            t.n_cand[type] = static_cast<uint32_t>(order.size());
            t.off_cand[type] = reserve(order.size()*2, 4);
            std::memcpy(&blob[t.off_cand[type]], order.data(), order.size()*2);
            all = reinterpret_cast<const ModeInfo*>(&blob[t.off_modes]);       // reserve() may have moved the blob
            gr = reinterpret_cast<const GridInfo*>(&blob[t.off_grids]);
            for (int q = 0; q < 5; ++q) {
                const size_t n = std::min<size_t>(order.size(), kPlanCounts[q][type]);
                std::vector<uint16_t> sub(order.begin(), order.begin() + n);
                // densest grids first: they set a low error early, which lets coarse grids be skipped
                std::stable_sort(sub.begin(), sub.end(), [&](uint16_t a, uint16_t b) {
                    if (all[a].nw != all[b].nw) return all[a].nw > all[b].nw;
                    if (all[a].grid != all[b].grid) return all[a].grid < all[b].grid;
                    return all[a].level > all[b].level;
                });
                t.n_cand_q[q][type] = static_cast<uint32_t>(n);
                t.off_cand_q[q][type] = reserve(n*2 + 2, 4);
                if (n) std::memcpy(&blob[t.off_cand_q[q][type]], sub.data(), n*2);
                all = reinterpret_cast<const ModeInfo*>(&blob[t.off_modes]);
                gr = reinterpret_cast<const GridInfo*>(&blob[t.off_grids]);
            }
        }
    }
    t.blob_bytes = static_cast<uint32_t>(blob.size());
    return out;
}

} // namespace astc
} // namespace cfx
