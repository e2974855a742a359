// DEVELOPER TOOL: host build of the lane-local BC7 search (cuttlefish_b200/csrc/bc7_core.cuh) with
// the warp-level steps (ranking, argmin) done by plain loops, so encoder quality can be studied
// without a GPU.  Not part of libcfx.so and never a fallback.
//   g++ -O2 -shared -fPIC -o tools/_build/libemu_bc7.so tools/emu_bc7.cpp
#include "../cuttlefish_b200/csrc/bc7_core.cuh"

#include <algorithm>
#include <vector>

using namespace cfx;
using namespace cfx::bc7;

extern "C" int emu_bc7_encode(const uint8_t* rgba, uint32_t w, uint32_t h, uint8_t* out, int G,
    uint32_t color_mask, uint32_t* dbg /* per block: mode, shape, variant, err */,
    const uint16_t* cand_opaque, const uint16_t* cand_alpha)
{
    uint32_t bxn = (w + 3)/4, byn = (h + 3)/4;
    uint32_t chmask = 0;
    for (int c = 0; c < 4; ++c) if (color_mask & (1u << c)) chmask |= 0xFFu << (8*c);
    for (uint32_t by = 0; by < byn; ++by)
        for (uint32_t bx = 0; bx < bxn; ++bx) {
            uint32_t px[16];
            float4 pxf[16];
            uint32_t amin = 255;
            for (int i = 0; i < 16; ++i) {
                uint32_t x = std::min(bx*4 + (i & 3), w - 1), y = std::min(by*4 + (i >> 2), h - 1);
                uint32_t v;
                memcpy(&v, rgba + (size_t(y)*w + x)*4, 4);
                px[i] = v;
                pxf[i] = make_float4(float(v & 0xFF), float((v >> 8) & 0xFF), float((v >> 16) & 0xFF), float(v >> 24));
                amin = std::min(amin, v >> 24);
            }
            bool has_alpha = amin < 255 && (color_mask & 8u);
            float sT[4] = {0, 0, 0, 0}, cT[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = 0; i < 16; ++i) {
                float4 x = pxf[i];
                sT[0] += x.x; sT[1] += x.y; sT[2] += x.z; sT[3] += x.w;
                cT[0] += x.x*x.x; cT[1] += x.x*x.y; cT[2] += x.x*x.z; cT[3] += x.x*x.w;
                cT[4] += x.y*x.y; cT[5] += x.y*x.z; cT[6] += x.y*x.w;
                cT[7] += x.z*x.z; cT[8] += x.z*x.w; cT[9] += x.w*x.w;
            }
            uint32_t keys[64];
            for (uint32_t s = 0; s < 64; ++s) keys[s] = score_shape(pxf, sT, cT, s);
            std::sort(keys, keys + 64);
            uint32_t best_key = 0xFFFFFFFFu;
            uint4 best_blk = make_uint4(0, 0, 0, 0);
            uint32_t bm = 0, bs = 0, bv = 0, be = 0;
            for (uint32_t sub = 0; sub < uint32_t(G); ++sub) {
                uint32_t desc = has_alpha ? cand_alpha[sub] : cand_opaque[sub];
                uint32_t rank = cand_rank(desc);
                uint32_t shape = rank == 0xFFFFFFFFu ? 0 : (keys[rank] & 63u);
                uint32_t mode = cand_mode(desc), variant = cand_variant(desc);
                uint32_t m1 = mode == 6 ? 0u : kBc7Part2[shape];
                Fit fit;
                if (cand_is_dual(desc)) fit_dual(px, mode, cand_rotation(desc), variant & 1u, cand_rounds(desc), chmask, fit);
                else fit_candidate(pxf, px, mode, m1, variant, cand_rounds(desc), chmask, fit);
                uint32_t total = fit.err[0] + fit.err[1];
                uint32_t key = (std::min(total, 0x03FFFFFFu) << 5) | sub;
                if (key < best_key) {
                    best_key = key;
                    best_blk = cand_is_dual(desc) ? pack_dual(mode, cand_rotation(desc), variant & 1u, fit) : pack_block(mode, shape, m1, fit);
                    bm = mode; bs = shape; bv = variant; be = total;
                }
            }
            size_t bi = size_t(by)*bxn + bx;
            memcpy(out + bi*16, &best_blk, 16);
            if (dbg) { dbg[bi*4] = bm; dbg[bi*4 + 1] = bs; dbg[bi*4 + 2] = bv; dbg[bi*4 + 3] = be; }
        }
    return 0;
}
