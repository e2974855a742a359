// DEVELOPER TOOL: host build of the lane-local BC6H encoder (cuttlefish_b200/csrc/bc6h_core.cuh).
#include "../cuttlefish_b200/csrc/bc6h_core.cuh"
#include <vector>
using namespace cfx;

extern "C" int emu_bc6h_encode(const uint16_t* rgba16f, uint32_t w, uint32_t h, uint8_t* out, uint32_t quality, uint32_t is_signed)
{
    uint32_t bxn = (w + 3)/4, byn = (h + 3)/4;
    std::vector<float> xs(bc6h::kWordsPerLane*32);
    for (uint32_t by = 0; by < byn; ++by)
        for (uint32_t bx = 0; bx < bxn; ++bx) {
            for (uint32_t t = 0; t < 16; ++t) {
                uint32_t x = std::min(bx*4 + (t & 3), w - 1), y = std::min(by*4 + (t >> 2), h - 1);
                for (uint32_t c = 0; c < 3; ++c) {
                    uint32_t hb = rgba16f[(size_t(y)*w + x)*4 + c];
                    float v = (hb & 0x8000u) ? 0.0f : float(std::min(hb, 0x7BFFu))*(64.0f/31.0f);
                    if (is_signed) { v = float(std::min(hb & 0x7FFFu, 0x7BFFu))*(32.0f/31.0f); if (hb & 0x8000u) v = -v; }
                    bc6h::px(xs.data(), 0, t, c) = v;
                }
            }
            uint4 blk = bc6h::encode_block(xs.data(), 0, quality, is_signed != 0);
            memcpy(out + (size_t(by)*bxn + bx)*16, &blk, 16);
        }
    return 0;
}
