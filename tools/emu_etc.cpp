// DEVELOPER TOOL: host build of the lane-local ETC encoder (cuttlefish_b200/csrc/etc_core.cuh).
#include "../cuttlefish_b200/csrc/etc_core.cuh"
#include <vector>
using namespace cfx;
extern "C" int emu_etc_encode(const float* rgba, uint32_t w, uint32_t h, uint8_t* out, uint32_t format, int rounds)
{
    uint32_t bxn = (w + 3)/4, byn = (h + 3)/4;
    std::vector<float> xs(16*4*32);
    const uint32_t bytes = format == 40 ? 16 : 8;
    for (uint32_t by = 0; by < byn; ++by)
        for (uint32_t bx = 0; bx < bxn; ++bx) {
            for (uint32_t t = 0; t < 16; ++t) {
                uint32_t x = std::min(bx*4 + (t & 3), w - 1), y = std::min(by*4 + (t >> 2), h - 1);
                for (uint32_t c = 0; c < 4; ++c)
                    etc::px(xs.data(), 0, t, c) = std::min(std::max(rgba[(size_t(y)*w + x)*4 + c], 0.0f), 1.0f)*255.0f;
            }
            uint8_t* dst = out + (size_t(by)*bxn + bx)*bytes;
            if (format == 40) { uint2 a = etc::encode_eac_alpha(xs.data(), 0, 2); memcpy(dst, &a, 8); dst += 8; }
            uint2 c = format == 39 ? etc::encode_color_a1(xs.data(), 0, rounds) : etc::encode_color(xs.data(), 0, format != 37, rounds);
            memcpy(dst, &c, 8);
        }
    return 0;
}
