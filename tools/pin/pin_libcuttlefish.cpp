// DEVELOPER TOOL (this container only): drives the reference's REAL public API -- cuttlefish::Image (FreeImage-backed,
// bottom-up storage) -> Texture::setImage -> Texture::convert -> Texture::data -- from a full libcuttlefish.so built out of
// tree from /root/reference (tools/pin/pin_libcuttlefish.py), to pin oracle/cfref.cpp + the stub-based glue against it.
//   pin_libcuttlefish <in.f32> <w> <h> <format> <type> <quality> <srgb> <out.bin>
// in.f32: w*h RGBA float32, row 0 = top.  format / type / quality: numeric values of Texture::Format / Type / Quality.
#include <cuttlefish/Image.h>
#include <cuttlefish/Texture.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace cuttlefish;
int main(int argc, char** argv)
{
	if (argc != 9) return 2;
	const unsigned w = std::atoi(argv[2]), h = std::atoi(argv[3]);
	std::vector<float> px(size_t(w)*h*4);
	FILE* f = std::fopen(argv[1], "rb");
	if (!f || std::fread(px.data(), sizeof(float), px.size(), f) != px.size()) return 3;
	std::fclose(f);
	const ColorSpace cs = std::atoi(argv[7]) ? ColorSpace::sRGB : ColorSpace::Linear;
	Image image;
	if (!image.initialize(Image::Format::RGBAF, w, h, cs)) return 4;
	for (unsigned y = 0; y < h; ++y) std::memcpy(image.scanline(y), &px[size_t(y)*w*4], size_t(w)*16);
	Texture texture(Texture::Dimension::Dim2D, w, h, 0, 1, cs);
	if (!texture.setImage(image)) return 5;
	if (!texture.convert(static_cast<Texture::Format>(std::atoi(argv[4])), static_cast<Texture::Type>(std::atoi(argv[5])),
			static_cast<Texture::Quality>(std::atoi(argv[6]))))
		return 6;
	f = std::fopen(argv[8], "wb");
	std::fwrite(texture.data(), 1, texture.dataSize(), f);
	std::fclose(f);
	return 0;
}
