// Extra per-footprint tables of the two-phase ASTC search (astc3.cu), appended to the blob that
// astc_tables.hpp builds.  For every weight grid g of the footprint, with P_g the bilinear infill
// matrix (texels x grid weights, ASTC spec "Weight Infill"):
//   M_g = pinv(P_g)      least-squares decimation: ideal texel weights -> ideal grid weights
//   R_g = I - P_g M_g    what decimation to grid g cannot represent
//   ce_gj = sum_i P_ij^2 what clamping grid weight j by one unit costs: the least-squares grid weights of bimodal
//                        content (text, hard edges) overshoot [0, 1] and are clamped; R t is orthogonal to range(P), so
//                        the clamped fit loses exactly |R t|^2 + |P o|^2 ~ |R t|^2 + sum_j ce_gj o_j^2, o = M t - clamp(M t)
// Both are stored as fp16 in the register layout of the B operand of mma.sync.m16n8k16, so that a warp
// gets decimated weights / decimation residuals of all its partition hypotheses ("slot planes", the
// A operand rows) from a handful of tensor-core instructions.
//
// What these tables replace in the reference: astcenc's per-decimation-mode ideal-weight computation and
// its error estimate, compute_ideal_weights_for_decimation / compute_error_of_weight_set_1plane
// (lib/astc-encoder/Source/astcenc_ideal_endpoints_and_weights.cpp:845, :688).
//
// Plain C++ (no CUDA): used by astc3.cu's host side and by the host tools.
#pragma once
#include "astc_tables.hpp"

#include <cmath>

namespace cfx {
namespace astc {

constexpr int kMaxGrids3 = 96;
constexpr int kMaxTexels3 = 144;    // up to 12x12
constexpr int kSlotTypes3 = 12;     // colour level / estimate list classes, see Astc3Tab::off_modecl
constexpr int kRows3 = 16;          // slot planes (A operand rows): 9 first planes, 4 second planes, 3 spare

struct Astc3Tab {
    uint32_t NT, KS;                // texel n-tiles (8 wide), texel k-steps (16 deep)
    uint32_t off_rfrag_idx;         // [n_grids] uint32: byte offset of grid's R fragments, 0 = full-resolution grid (R = 0)
    uint32_t off_mfrag_idx;         // [n_grids] uint32: byte offset of grid's M fragments
    uint32_t off_kappa;             // [n_grids][kMaxTexels3] float: kappa_gi = sum_j P_ij^2
    uint32_t off_ksum;              // [n_grids] float: sum_i kappa_gi
    uint32_t off_colenergy;         // [n_grids][64] float: ce_gj = sum_i P_ij^2 (0 beyond the grid's weights)
    uint32_t off_modecl;            // [2 alpha][8 slot type][n_modes1 + n_modes2] u8 colour level, 0xFF = does not fit
                                    // slot types: 0..2 = 1..3 subsets, 3 = dual plane, 4 / 5 / 6 = one / two / three subsets with luminance end points,
                                    // 7 / 8 / 9 = one / two / three subsets with RGB base + scale end points (CEM 6; one
                                    // subset of a block with alpha: CEM 10 = base + scale + alpha pair),
                                    // 10 = two subsets, MIXED end point modes, opaque: one subset base + scale (CEM 6), one RGB (CEM 8),
                                    // 11 = two subsets, mixed, block with alpha: one subset RGB (CEM 8, alpha 255), one RGBA (CEM 12)
    uint32_t n_modes;               // n_modes1 + n_modes2
    uint32_t off_rstream;           // R fragments of the decimated grids, contiguous in off_dec_list order (+1 pad tile)
    uint32_t off_dec_list;          // [n_dec] u8 grid index
    uint32_t n_dec;                 // number of decimated grids (nw < texels)
    uint32_t MW;                    // 64-bit words per texel mask
    uint32_t off_part2w;            // [1024][MW] uint64 texel mask of subset 1 (all zero = unusable seed)
    uint32_t off_part3w;            // [1024][2][MW] uint64 masks of subsets 1 and 2
    uint32_t n_seed2, n_seed3;      // usable (distinct, non-degenerate) seeds
    uint32_t off_seed2, off_seed3;  // [n_seed] uint16 seed numbers, ascending
    uint32_t off_part2c;            // [n_seed2][MW] the same masks, compacted in off_seed2 order
    uint32_t off_part3c;            // [n_seed3][2][MW]
    uint32_t off_est[2][kSlotTypes3];         // [alpha][slot type] -> uint4 list of the modes that fit:
    uint32_t n_est[2][kSlotTypes3];           //   {f32 rest, f32 kc = cvar(colour level), mode | grid << 16 | cl << 24, min(level, 5) | f16 a << 16}: see below
};

inline uint16_t f32_to_f16_bits(float f)
{
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7FFFFFFFu;
    if (x >= 0x47800000u) return static_cast<uint16_t>(sign | 0x7BFFu);            // clamp to max finite
    if (x < 0x38800000u) {                                                         // subnormal / zero
        if (x < 0x33000000u) return static_cast<uint16_t>(sign);
        const int shift = 113 - static_cast<int>(x >> 23);
        const uint32_t mant = (x & 0x7FFFFFu) | 0x800000u;
        uint32_t h = mant >> (shift + 13);
        const uint32_t rem = mant & ((1u << (shift + 13)) - 1u), half = 1u << (shift + 12);
        if (rem > half || (rem == half && (h & 1u))) ++h;
        return static_cast<uint16_t>(sign | h);
    }
    uint32_t h = ((x - 0x38000000u) >> 13);
    const uint32_t rem = x & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
    return static_cast<uint16_t>(sign | h);
}

// Solves (P^T P) M = P^T in double precision (Gauss-Jordan with partial pivoting).
inline void pinv_infill(const std::vector<double>& P, int T, int nw, std::vector<double>& M)
{
    std::vector<double> A(static_cast<size_t>(nw)*nw, 0.0);
    M.assign(static_cast<size_t>(nw)*T, 0.0);
    for (int a = 0; a < nw; ++a) {
        for (int b = 0; b < nw; ++b) {
            double s = 0;
            for (int i = 0; i < T; ++i) s += P[i*nw + a]*P[i*nw + b];
            A[a*nw + b] = s + (a == b ? 1e-9 : 0.0);
        }
        for (int i = 0; i < T; ++i) M[a*T + i] = P[i*nw + a];
    }
    for (int col = 0; col < nw; ++col) {
        int piv = col;
        for (int r = col + 1; r < nw; ++r) if (std::fabs(A[r*nw + col]) > std::fabs(A[piv*nw + col])) piv = r;
        if (piv != col) {
            for (int k = 0; k < nw; ++k) std::swap(A[piv*nw + k], A[col*nw + k]);
            for (int k = 0; k < T; ++k) std::swap(M[piv*T + k], M[col*T + k]);
        }
        const double inv = 1.0/A[col*nw + col];
        for (int k = 0; k < nw; ++k) A[col*nw + k] *= inv;
        for (int k = 0; k < T; ++k) M[col*T + k] *= inv;
        for (int r = 0; r < nw; ++r) {
            if (r == col) continue;
            const double f = A[r*nw + col];
            if (f == 0.0) continue;
            for (int k = 0; k < nw; ++k) A[r*nw + k] -= f*A[col*nw + k];
            for (int k = 0; k < T; ++k) M[r*T + k] -= f*M[col*T + k];
        }
    }
}

// Appends the tables to b.blob and returns their offsets.
inline Astc3Tab build_tables3(Built& b)
{
    Astc3Tab t3;
    std::memset(&t3, 0, sizeof(t3));
    AstcTab& t = b.tab;
    std::vector<uint8_t>& blob = b.blob;
    const int T = static_cast<int>(t.texels);
    const int G = static_cast<int>(t.n_grids);
    auto reserve = [&](size_t bytes, size_t align) {
        size_t off = (blob.size() + align - 1)/align*align;
        blob.resize(off + bytes, 0);
        return static_cast<uint32_t>(off);
    };
    t3.NT = static_cast<uint32_t>((T + 7)/8);
    t3.KS = static_cast<uint32_t>((T + 15)/16);
    t3.n_modes = t.n_modes1 + t.n_modes2;
    t3.off_rfrag_idx = reserve(static_cast<size_t>(G)*4, 4);
    t3.off_mfrag_idx = reserve(static_cast<size_t>(G)*4, 4);
    t3.off_kappa = reserve(static_cast<size_t>(G)*kMaxTexels3*4, 4);
    t3.off_ksum = reserve(static_cast<size_t>(G)*4, 4);
    t3.off_colenergy = reserve(static_cast<size_t>(G)*64*4, 16);
    const int KS = static_cast<int>(t3.KS), NT = static_cast<int>(t3.NT);
    for (int g = 0; g < G; ++g) {
        const GridInfo gi = reinterpret_cast<const GridInfo*>(&blob[t.off_grids])[g];
        const int nw = gi.nw;
        std::vector<double> P(static_cast<size_t>(T)*nw, 0.0), M;
        for (int i = 0; i < T; ++i) {
            const uint8_t* inf = &blob[t.off_infill + (static_cast<size_t>(g)*T + i)*8];
            for (int k = 0; k < 4; ++k) P[i*nw + inf[k]] += inf[4 + k]/16.0;
        }
        double ksum = 0;
        for (int i = 0; i < T; ++i) {
            double kap = 0;
            for (int j = 0; j < nw; ++j) kap += P[i*nw + j]*P[i*nw + j];
            reinterpret_cast<float*>(&blob[t3.off_kappa])[g*kMaxTexels3 + i] = static_cast<float>(kap);
            ksum += kap;
        }
        reinterpret_cast<float*>(&blob[t3.off_ksum])[g] = static_cast<float>(ksum);
        for (int j = 0; j < nw && j < 64; ++j) {
            double ce = 0;
            for (int i = 0; i < T; ++i) ce += P[i*nw + j]*P[i*nw + j];
            reinterpret_cast<float*>(&blob[t3.off_colenergy])[g*64 + j] = static_cast<float>(ce);
        }
        pinv_infill(P, T, nw, M);
        // M fragments: B[k = texel][n = grid weight] = M[n][k]
        const int NTW = (nw + 7)/8;
        {
            const uint32_t off = reserve(static_cast<size_t>(NTW)*KS*32*8, 16);
            reinterpret_cast<uint32_t*>(&blob[t3.off_mfrag_idx])[g] = off;
            uint16_t* dst = reinterpret_cast<uint16_t*>(&blob[off]);
            for (int nt = 0; nt < NTW; ++nt)
                for (int ks = 0; ks < KS; ++ks)
                    for (int lane = 0; lane < 32; ++lane) {
                        const int n = nt*8 + (lane >> 2), k0 = ks*16 + 2*(lane & 3);
                        uint16_t* d = dst + ((static_cast<size_t>(nt)*KS + ks)*32 + lane)*4;
                        const int ks4[4] = {k0, k0 + 1, k0 + 8, k0 + 9};
                        for (int e = 0; e < 4; ++e)
                            d[e] = (n < nw && ks4[e] < T) ? f32_to_f16_bits(static_cast<float>(M[n*T + ks4[e]])) : 0;
                    }
        }
        // R fragments: B[k = texel][n = texel i] = R[i][k]
        if (nw < T) {
            std::vector<double> R(static_cast<size_t>(T)*T);
            for (int i = 0; i < T; ++i)
                for (int k = 0; k < T; ++k) {
                    double s = i == k ? 1.0 : 0.0;
                    for (int j = 0; j < nw; ++j) s -= P[i*nw + j]*M[j*T + k];
                    R[i*T + k] = s;
                }
            const uint32_t off = reserve(static_cast<size_t>(NT)*KS*32*8, 16);
            reinterpret_cast<uint32_t*>(&blob[t3.off_rfrag_idx])[g] = off;
            uint16_t* dst = reinterpret_cast<uint16_t*>(&blob[off]);
            for (int nt = 0; nt < NT; ++nt)
                for (int ks = 0; ks < KS; ++ks)
                    for (int lane = 0; lane < 32; ++lane) {
                        const int n = nt*8 + (lane >> 2), k0 = ks*16 + 2*(lane & 3);
                        uint16_t* d = dst + ((static_cast<size_t>(nt)*KS + ks)*32 + lane)*4;
                        const int ks4[4] = {k0, k0 + 1, k0 + 8, k0 + 9};
                        for (int e = 0; e < 4; ++e)
                            d[e] = (n < T && ks4[e] < T) ? f32_to_f16_bits(static_cast<float>(R[n*T + ks4[e]])) : 0;
                    }
        }
    }
    // colour level of every (alpha, slot type, mode)
    t3.off_modecl = reserve(static_cast<size_t>(2)*kSlotTypes3*t3.n_modes, 4);
    for (int alpha = 0; alpha < 2; ++alpha)
        for (int type = 0; type < kSlotTypes3; ++type)
            for (uint32_t mi = 0; mi < t3.n_modes; ++mi) {
                const ModeInfo m = reinterpret_cast<const ModeInfo*>(&blob[t.off_modes])[mi];
                const int pc = (type == 1 || type == 5 || type == 8 || type >= 10) ? 2 : ((type == 2 || type == 6 || type == 9) ? 3 : 1);
                // luminance end points (CEM 0: L0 L1; with alpha CEM 4: L0 L1 A0 A1)
                // base + scale (CEM 6: R G B s; with alpha CEM 10: R G B s A0 A1)
                int n_ints = type >= 7 ? pc*(alpha ? 6 : 4) : (type >= 4 ? pc*(alpha ? 4 : 2) : pc*(alpha ? 8 : 6));
                if (type == 10) n_ints = 10;
                if (type == 11) n_ints = 14;
                // mixed modes spill 3*pc - 4 bits of their end point mode field to just below the weights
                const int extra = type >= 10 ? 3*pc - 4 : 0;
                const int avail = 128 - static_cast<int>(m.wbits) - (pc == 1 ? 17 : 29) - (type == 3 ? 2 : 0) - extra;
                uint8_t cl = 0xFF;
                const bool usable = !((type == 10 && alpha) || (type == 11 && !alpha));     // types 8, 9 with alpha: CEM 10 on every subset
                if (usable && (m.dual != 0) == (type == 3) && n_ints <= 18 && avail >= 0) cl = blob[t.off_clevel + (n_ints >> 1)*128 + avail];
                blob[t3.off_modecl + (static_cast<size_t>(alpha)*kSlotTypes3 + type)*t3.n_modes + mi] = cl;
            }
    // the decimated grids' R fragments again as one contiguous stream (phase 1a walks it with a one-tile prefetch)
    {
        std::vector<uint8_t> dec;
        for (int g = 0; g < G; ++g)
            if (reinterpret_cast<const uint32_t*>(&blob[t3.off_rfrag_idx])[g]) dec.push_back(static_cast<uint8_t>(g));
        t3.n_dec = static_cast<uint32_t>(dec.size());
        t3.off_dec_list = reserve(dec.size() + 4, 4);
        if (!dec.empty()) std::memcpy(&blob[t3.off_dec_list], dec.data(), dec.size());
        const size_t per = static_cast<size_t>(NT)*KS*32*8;
        t3.off_rstream = reserve((dec.size() + 1)*per, 16);
        for (size_t k = 0; k < dec.size(); ++k) {
            const uint32_t src = reinterpret_cast<const uint32_t*>(&blob[t3.off_rfrag_idx])[dec[k]];
            std::memcpy(&blob[t3.off_rstream + k*per], &blob[src], per);
        }
    }
    // partition masks, any footprint (spec "Partition Pattern Generation"; duplicates and degenerate seeds dropped,
    // like astcenc's init_partition_tables, lib/astc-encoder/Source/astcenc_partition_tables.cpp)
    {
        const int MW = (T + 63)/64;
        t3.MW = static_cast<uint32_t>(MW);
        t3.off_part2w = reserve(static_cast<size_t>(1024)*MW*8, 8);
        t3.off_part3w = reserve(static_cast<size_t>(1024)*2*MW*8, 8);
        const bool small_block = T < 31;
        const int bw = static_cast<int>(t.bw);
        std::vector<std::vector<uint64_t>> seen2, seen3;
        std::vector<uint64_t> full(MW, 0);
        for (int i = 0; i < T; ++i) full[i >> 6] |= 1ull << (i & 63);
        for (int seed = 0; seed < 1024; ++seed) {
            std::vector<uint64_t> m1(MW, 0), comp(MW, 0);
            for (int i = 0; i < T; ++i)
                if (select_partition(seed, i % bw, i/bw, 0, 2, small_block) == 1) m1[i >> 6] |= 1ull << (i & 63);
            for (int w = 0; w < MW; ++w) comp[w] = m1[w] ^ full[w];
            bool ok = m1 != std::vector<uint64_t>(MW, 0) && m1 != full;
            for (auto& sd : seen2) if (sd == m1 || sd == comp) ok = false;
            if (ok) {
                seen2.push_back(m1);
                std::memcpy(&blob[t3.off_part2w + static_cast<size_t>(seed)*MW*8], m1.data(), MW*8);
            }
            std::vector<uint64_t> a(MW, 0), b2(MW, 0), c(MW, 0);
            for (int i = 0; i < T; ++i) {
                const int pp = select_partition(seed, i % bw, i/bw, 0, 3, small_block);
                if (pp == 1) a[i >> 6] |= 1ull << (i & 63);
                if (pp == 2) b2[i >> 6] |= 1ull << (i & 63);
            }
            for (int w = 0; w < MW; ++w) c[w] = full[w] ^ a[w] ^ b2[w];
            const std::vector<uint64_t> zero(MW, 0);
            ok = a != zero && b2 != zero && c != zero;
            std::vector<std::vector<uint64_t>> key = {a, b2, c};
            std::sort(key.begin(), key.end());
            std::vector<uint64_t> flat;
            for (auto& k : key) flat.insert(flat.end(), k.begin(), k.end());
            for (auto& sd : seen3) if (sd == flat) ok = false;
            if (ok) {
                seen3.push_back(flat);
                std::memcpy(&blob[t3.off_part3w + (static_cast<size_t>(seed)*2)*MW*8], a.data(), MW*8);
                std::memcpy(&blob[t3.off_part3w + (static_cast<size_t>(seed)*2 + 1)*MW*8], b2.data(), MW*8);
            }
        }
    }
    // compacted copies for the seed scan (every lane of the warp then has a usable seed in every iteration)
    {
        const int MW = static_cast<int>(t3.MW);
        std::vector<uint16_t> s2, s3;
        std::vector<uint64_t> m2, m3;
        for (int seed = 0; seed < 1024; ++seed) {
            const uint64_t* a = reinterpret_cast<const uint64_t*>(&blob[t3.off_part2w + static_cast<size_t>(seed)*MW*8]);
            bool any = false;
            for (int w = 0; w < MW; ++w) any |= a[w] != 0;
            if (any) { s2.push_back(static_cast<uint16_t>(seed)); m2.insert(m2.end(), a, a + MW); }
            const uint64_t* b3 = reinterpret_cast<const uint64_t*>(&blob[t3.off_part3w + static_cast<size_t>(seed)*2*MW*8]);
            any = false;
            for (int w = 0; w < MW; ++w) any |= b3[w] != 0;
            if (any) { s3.push_back(static_cast<uint16_t>(seed)); m3.insert(m3.end(), b3, b3 + 2*MW); }
        }
        t3.n_seed2 = static_cast<uint32_t>(s2.size()); t3.n_seed3 = static_cast<uint32_t>(s3.size());
        t3.off_seed2 = reserve(s2.size()*2 + 2, 4); t3.off_seed3 = reserve(s3.size()*2 + 2, 4);
        t3.off_part2c = reserve(m2.size()*8 + 8, 8); t3.off_part3c = reserve(m3.size()*8 + 8, 8);
        if (!s2.empty()) { std::memcpy(&blob[t3.off_seed2], s2.data(), s2.size()*2); std::memcpy(&blob[t3.off_part2c], m2.data(), m2.size()*8); }
        if (!s3.empty()) { std::memcpy(&blob[t3.off_seed3], s3.data(), s3.size()*2); std::memcpy(&blob[t3.off_part3c], m3.data(), m3.size()*8); }
    }
    // estimate lists: per (alpha, slot type) the modes that fit, with their model terms
    for (int alpha = 0; alpha < 2; ++alpha)
        for (int type = 0; type < kSlotTypes3; ++type) {
            std::vector<uint32_t> ent;
            for (uint32_t mi = 0; mi < t3.n_modes; ++mi) {
                const uint8_t cl = blob[t3.off_modecl + (static_cast<size_t>(alpha)*kSlotTypes3 + type)*t3.n_modes + mi];
                if (cl == 0xFF) continue;
                const ModeInfo m = reinterpret_cast<const ModeInfo*>(&blob[t.off_modes])[mi];
                const float n1 = static_cast<float>(kWeightQuant[m.level].n - 1);
                const float kq = (1.0f/(n1*n1))*(1.0f/12.0f)*(1.0f - 0.75f/n1);
                const float step = 255.0f/static_cast<float>(kColorQuant[cl].n - 1);
                const float kc = step*step*(1.0f/18.0f);
                // weight-quantisation term = a*measured(level) + rest: coarse levels blend the loss measured on the slot's
                // ideal weights (weight a = grid weights / texels: it holds where the texel weights themselves are
                // quantised, decimated grid weights are blends of them; footprints above 64 texels have no
                // full-resolution grid and take it as it is) with the uniform model kq; fine levels are model only
                float a = 0.0f;
                if (m.level < 6) a = T > 64 ? 1.0f : static_cast<float>(m.nw)/static_cast<float>(T);
                const float rest = (1.0f - a)*kq;
                uint32_t w[4];
                std::memcpy(&w[0], &rest, 4); std::memcpy(&w[1], &kc, 4);
                w[2] = mi | (static_cast<uint32_t>(m.grid) << 16) | (static_cast<uint32_t>(cl) << 24);
                w[3] = (m.level < 6 ? m.level : 5u) | (static_cast<uint32_t>(f32_to_f16_bits(a)) << 16);
                ent.insert(ent.end(), w, w + 4);
            }
            t3.n_est[alpha][type] = static_cast<uint32_t>(ent.size()/4);
            t3.off_est[alpha][type] = reserve(ent.size()*4 + 16, 16);
            if (!ent.empty()) std::memcpy(&blob[t3.off_est[alpha][type]], ent.data(), ent.size()*4);
        }
    t.blob_bytes = static_cast<uint32_t>(blob.size());
    return t3;
}

} // namespace astc
} // namespace cfx
