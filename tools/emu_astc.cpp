// DEVELOPER TOOL: host build of the lane-local ASTC search (cuttlefish_b200/csrc/astc_core.cuh) with
// the warp replaced by a loop over 32 "lanes".  Not part of libcfx.so and never a fallback.
#include "../cuttlefish_b200/csrc/astc_core.cuh"

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace cfx;
using namespace cfx::astc;

static Built g_built;
static int g_bw = 0, g_bh = 0;

extern "C" int emu_astc_encode(const float* rgba /* 0..1 floats */, uint32_t w, uint32_t h, uint32_t bw, uint32_t bh,
    uint8_t* out, uint32_t slots, uint32_t refine, uint32_t quality, uint32_t* dbg /* per block: slot, mode index, err, pc */)
{
    if (g_bw != (int)bw || g_bh != (int)bh) { g_built = build_tables(bw, bh); g_bw = bw; g_bh = bh; }
    Ctx c; c.blob = g_built.blob.data(); c.tab = g_built.tab;
    const uint32_t T = bw*bh;
    const uint32_t bxn = (w + bw - 1)/bw, byn = (h + bh - 1)/bh;
    static BlockState st;
    std::vector<uint8_t> u_scr(64*32*4), w_scr(128*32*4);
    for (uint32_t by = 0; by < byn; ++by)
        for (uint32_t bx = 0; bx < bxn; ++bx) {
            bool constant = true, has_alpha = false;
            for (uint32_t i = 0; i < T; ++i) {
                uint32_t x = std::min(bx*bw + i % bw, w - 1), y = std::min(by*bh + i/bw, h - 1);
                const float* p = rgba + (size_t(y)*w + x)*4;
                auto cl = [](float v) { return std::min(std::max(v, 0.0f), 1.0f)*255.0f; };
                st.cf[i] = make_float4(cl(p[0]), cl(p[1]), cl(p[2]), cl(p[3]));
                if (st.cf[i].w != 255.0f) has_alpha = true;
                if (memcmp(&st.cf[i], &st.cf[0], 16) != 0) constant = false;
            }
            st.has_alpha = has_alpha;
            size_t bi = size_t(by)*bxn + bx;
            uint4 blk;
            if (constant) {
                blk = pack_void_extent(st.cf[0]);
                if (dbg) { dbg[bi*4] = 9; dbg[bi*4 + 1] = 0; dbg[bi*4 + 2] = 0; dbg[bi*4 + 3] = 0; }
                memcpy(out + bi*16, &blk, 16);
                continue;
            }
            for (uint32_t i = 0; i < kSlots; ++i) st.slots[i].valid = 0;
            for (uint32_t lane = 0; lane < 32; ++lane) step_init(c, st, lane);
            if (slots > 1) {
                for (uint32_t lane = 0; lane < 32; ++lane) step_rank(c, st, lane);
                for (uint32_t lane = 0; lane < 32; ++lane) step_score(c, st, lane);
                for (uint32_t lane = 0; lane < 32; ++lane) step_slots(c, st, lane);
            }
            float best_err[32];
            uint32_t best_mode[32], best_slot[32];
            for (uint32_t lane = 0; lane < 32; ++lane) { best_err[lane] = 3.0e38f; best_mode[lane] = 0; best_slot[lane] = 0; }
            Plan plan = make_plan(quality, c.tab);
            if (slots < plan.slots) plan.slots = slots;
            if (const char* ev = getenv("EMU_COUNTS")) {
                unsigned a, b, cc, d;
                if (sscanf(ev, "%u,%u,%u,%u", &a, &b, &cc, &d) == 4) { plan.n_cand[0] = a; plan.n_cand[1] = b; plan.n_cand[2] = cc; plan.n_cand[3] = d; }
            }
            for (uint32_t s = 0; s < plan.slots; ++s) {
                if (!st.slots[s].valid) continue;
                const uint32_t type = slot_type(s);
                for (uint32_t base = 0; base < plan.n_cand[type]; base += 32)
                    for (uint32_t lane = 0; lane < 32; ++lane) {
                        if (base + lane >= plan.n_cand[type]) continue;
                        const uint32_t mi = tab_u16(c, c.tab.off_cand[type] + (base + lane)*2u);
                        Enc e;
                        evaluate(c, st.cf, st.slots[s], tab_mode(c, mi), has_alpha, u_scr.data(), w_scr.data(), lane, -1, e);
                        if (e.err < best_err[lane]) { best_err[lane] = e.err; best_mode[lane] = mi; best_slot[lane] = s; }
                    }
            }
            float win = 3.0e38f;
            uint32_t wl = 0;
            Enc encs[32];
            for (uint32_t lane = 0; lane < 32; ++lane) {
                if (best_err[lane] >= 3.0e38f) continue;
                evaluate(c, st.cf, st.slots[best_slot[lane]], tab_mode(c, best_mode[lane]), has_alpha, u_scr.data(),
                    w_scr.data(), lane, (int)plan.refine, encs[lane]);
                if (encs[lane].err < win) { win = encs[lane].err; wl = lane; }
            }
            blk = pack_block(c, st.slots[best_slot[wl]], tab_mode(c, best_mode[wl]), encs[wl], has_alpha, u_scr.data(), wl);
            memcpy(out + bi*16, &blk, 16);
            if (dbg) {
                dbg[bi*4] = best_slot[wl]; dbg[bi*4 + 1] = best_mode[wl]; dbg[bi*4 + 2] = (uint32_t)win;
                dbg[bi*4 + 3] = st.slots[best_slot[wl]].pc;
            }
        }
    return 0;
}

extern "C" int emu_astc_mode_info(uint32_t bw, uint32_t bh, uint32_t index, uint32_t* out)
{
    if (g_bw != (int)bw || g_bh != (int)bh) { g_built = build_tables(bw, bh); g_bw = bw; g_bh = bh; }
    Ctx c; c.blob = g_built.blob.data(); c.tab = g_built.tab;
    if (index >= c.tab.n_modes1 + c.tab.n_modes2) return -1;
    ModeInfo m = tab_mode(c, index);
    const GridInfo* g = reinterpret_cast<const GridInfo*>(c.blob + c.tab.off_grids) + m.grid;
    out[0] = g->w; out[1] = g->h; out[2] = kWqN[m.level]; out[3] = m.wbits;
    return 0;
}

#ifdef CFX_COUNT_OPS
extern "C" void emu_astc_ops(unsigned long long* out) { for (int i = 0; i < 8; ++i) out[i] = g_ops[i]; }
#endif
