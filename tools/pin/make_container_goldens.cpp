// DEVELOPER TOOL (this container only): the reference's real Texture::save() for DDS and KTX, driven through the public
// API of a full libcuttlefish.so built out of tree (see pin_libcuttlefish.py), to produce the golden container headers
// and two whole golden files under tests/golden/containers/ (tools/pin/make_container_goldens.py).
//   make_container_goldens <in.f32> <w> <h> <format> <type> <quality> <srgb> <alpha> <mips: 0 = none, 1 = full chain> <filetype 1=DDS 2=KTX> <out>
#include <cuttlefish/Image.h>
#include <cuttlefish/Texture.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace cuttlefish;
int main(int argc, char** argv)
{
	if (argc != 12) return 2;
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
	if (std::atoi(argv[9]) && !texture.generateMipmaps(Image::ResizeFilter::CatmullRom)) return 7;
	if (!texture.convert(static_cast<Texture::Format>(std::atoi(argv[4])), static_cast<Texture::Type>(std::atoi(argv[5])),
			static_cast<Texture::Quality>(std::atoi(argv[6])), static_cast<Texture::Alpha>(std::atoi(argv[8]))))
		return 6;
	Texture::SaveResult r = texture.save(argv[11], static_cast<Texture::FileType>(std::atoi(argv[10])));
	return r == Texture::SaveResult::Success ? 0 : (r == Texture::SaveResult::Unsupported ? 10 : 11);
}
