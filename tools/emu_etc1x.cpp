// DEVELOPER TOOL: host build of the byte-exact ETC1 restatement (csrc/etc1_exact.cuh).  Driven by tools/emu_etc1x.py.
#include "../cuttlefish_b200/csrc/etc1_exact.cuh"
#include <cmath>
using namespace cfx;
extern "C" int emu_etc1x_encode(const float* rgba, uint32_t w, uint32_t h, uint8_t* out, float effort, int rec709)
{
    uint32_t bxn = (w + 3)/4, byn = (h + 3)/4;
    for (uint32_t by = 0; by < byn; ++by)
        for (uint32_t bx = 0; bx < bxn; ++bx) {
            etc1x::Px src[16];
            for (uint32_t x = 0; x < 4; ++x)
                for (uint32_t y = 0; y < 4; ++y) {
                    etc1x::Px& p = src[x*4 + y];
                    const uint32_t sx = bx*4 + x, sy = by*4 + y;
                    if (sx >= w || sy >= h) { p.r = p.g = p.b = 0.0f; p.a = NAN; continue; }
                    const float* s = rgba + (size_t(sy)*w + sx)*4;
                    p.r = etc1x::clamp01(s[0]); p.g = etc1x::clamp01(s[1]); p.b = etc1x::clamp01(s[2]); p.a = 1.0f;   // formats without alpha: source alpha = 1 (EtcBlock4x4.cpp:333-338)
                }
            uint2 b = etc1x::encode_etc1_exact(src, effort, rec709 != 0);
            memcpy(out + (size_t(by)*bxn + bx)*8, &b, 8);
        }
    return 0;
}
