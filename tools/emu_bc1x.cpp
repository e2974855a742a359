// DEVELOPER TOOL: host build of the byte-exact BC1 restatement (csrc/bc1_exact.cuh; rgbcx levels 0 / 4 / 9 =
// Texture::Quality Lowest / Low / Normal).  Driven by tools/emu_bc1x.py.
#include "../cuttlefish_b200/csrc/bc1_exact.cuh"
using namespace cfx;
extern "C" int emu_bc1x_encode(const uint8_t* rgba, uint32_t w, uint32_t h, uint8_t* out, int allow3, int allow_black, int quality)
{
    uint32_t bxn = (w + 3)/4, byn = (h + 3)/4;
    for (uint32_t by = 0; by < byn; ++by)
        for (uint32_t bx = 0; bx < bxn; ++bx) {
            uint32_t px[16];
            for (int i = 0; i < 16; ++i) {
                uint32_t x = std::min(bx*4 + (i & 3), w - 1), y = std::min(by*4 + (i >> 2), h - 1);
                memcpy(&px[i], rgba + (size_t(y)*w + x)*4, 4);
            }
            uint2 b = rgbcx9::encode_bc1_exact(px, static_cast<uint32_t>(quality), allow3 != 0, allow_black != 0);
            memcpy(out + (size_t(by)*bxn + bx)*8, &b, 8);
        }
    return 0;
}
