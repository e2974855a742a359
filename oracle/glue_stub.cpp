// TEST INFRASTRUCTURE: lets the reference's REAL converter glue (lib/src/Converter.cpp,
// S3tcConverter.cpp, EtcConverter.cpp, AstcConverter.cpp, StandardConverter.cpp, HalfFloat.cpp,
// Shared.cpp -- compiled from /root/reference by oracle/Makefile, never copied) run without FreeImage:
// it defines only the handful of cuttlefish::Image / cuttlefish::Texture members those files touch,
// over a plain float buffer, and exports cfglue_encode(), which is what Texture::convert() does on
// this path (lib/src/Texture.cpp:1536-1561): set format/type/alpha/mask, call Converter::convert.
// tests/test_oracle_cpu.py uses it to pin oracle/cfref.cpp (our glue restatement) byte for byte.
#include <cuttlefish/Image.h>
#include <cuttlefish/Texture.h>

#include "Converter.h"

#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

// cuttlefish::Image stores its rows bottom-up (FreeImage, lib/src/Image.cpp:340-343); cfglue_set_bottom_up(1) makes this
// stand-in do the same, so that the adapter's CFX_FLAG_BOTTOM_UP path is exercised the way the real Image would.
static bool g_bottom_up = false;

namespace cuttlefish
{

struct Image::Impl
{
	Format format;
	ColorSpace colorSpace;
	unsigned int width, height;
	std::vector<float> data;     // RGBAF rows, top-down (scanline(y) = row y from the top, lib/src/Image.cpp:1092-1098)
};

Image::Image() {}
Image::Image(Format format, unsigned int width, unsigned int height, ColorSpace colorSpace)
	: m_impl(new Impl{format, colorSpace, width, height, std::vector<float>(std::size_t(width)*height*4)})
{
}
Image::~Image() {}
Image::Image(Image&& other) noexcept = default;
Image& Image::operator=(Image&& other) noexcept = default;
bool Image::isValid() const {return m_impl != nullptr;}
Image::operator bool() const {return m_impl != nullptr;}
Image::Format Image::format() const {return m_impl->format;}
ColorSpace Image::colorSpace() const {return m_impl->colorSpace;}
unsigned int Image::width() const {return m_impl->width;}
unsigned int Image::height() const {return m_impl->height;}
void* Image::scanline(unsigned int y) {return m_impl->data.data() + std::size_t(g_bottom_up ? m_impl->height - 1 - y : y)*m_impl->width*4;}
const void* Image::scanline(unsigned int y) const {return m_impl->data.data() + std::size_t(g_bottom_up ? m_impl->height - 1 - y : y)*m_impl->width*4;}
void Image::reset() {m_impl.reset();}

struct Texture::Impl
{
	Format format;
	Type type;
	Alpha alphaType;
	ColorMask colorMask;
	ColorSpace colorSpace;
};

Texture::Texture() : m_impl(new Impl{}) {}
ColorSpace Texture::colorSpace() const {return m_impl->colorSpace;}
// Texture::setImage records the image's colour space (lib/src/Texture.cpp:1255-1283); that is all this stub keeps
bool Texture::setImage(const Image& image, unsigned int, unsigned int)
{
	m_impl->colorSpace = image.colorSpace();
	return true;
}
Texture::~Texture() {}
Texture::Format Texture::format() const {return m_impl->format;}
Texture::Type Texture::type() const {return m_impl->type;}
Texture::Alpha Texture::alphaType() const {return m_impl->alphaType;}
Texture::ColorMask Texture::colorMask() const {return m_impl->colorMask;}

// The state-setting half of Texture::convert (lib/src/Texture.cpp:1544-1548); the converting half is
// Converter::convert, called from cfglue_encode below on the surface it builds.
bool Texture::convert(Format format, Type type, Quality, Alpha alphaType, ColorMask colorMask, unsigned int)
{
	m_impl->format = format;
	m_impl->type = type;
	m_impl->alphaType = alphaType;
	m_impl->colorMask = colorMask;
	return true;
}

} // namespace cuttlefish

using namespace cuttlefish;

struct cfglue_desc { uint32_t format, type, quality, alpha_type, color_mask, color_space, width, height; };

extern "C" void cfglue_set_bottom_up(int on) {g_bottom_up = on != 0;}

// convert_seconds (may be null) receives the wall time of the Converter::convert() call alone -- what Texture::convert()
// costs once the RGBAF image exists -- without the set-up copy this stand-in makes.
extern "C" int cfglue_encode_timed(const cfglue_desc* d, const float* src, size_t pitch_floats, uint8_t* dst, size_t dst_size,
	unsigned threads, double* convert_seconds)
{
	Texture texture;
	texture.convert(static_cast<Texture::Format>(d->format), static_cast<Texture::Type>(d->type),
		static_cast<Texture::Quality>(d->quality), static_cast<Texture::Alpha>(d->alpha_type),
		Texture::ColorMask((d->color_mask & 1) != 0, (d->color_mask & 2) != 0, (d->color_mask & 4) != 0, (d->color_mask & 8) != 0), 1);
	Converter::MipImageList images(1);
	images[0].resize(1);
	images[0][0].emplace_back(Image::Format::RGBAF, d->width, d->height, d->color_space ? ColorSpace::sRGB : ColorSpace::Linear);
	Image& img = images[0][0][0];
	for (unsigned int y = 0; y < d->height; ++y)
		std::memcpy(img.scanline(y), src + std::size_t(y)*pitch_floats, std::size_t(d->width)*4*sizeof(float));
	texture.setImage(img);
	Converter::MipTextureList out;
	if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
	const auto t0 = std::chrono::steady_clock::now();
	const bool ok = Converter::convert(texture, images, out, static_cast<Texture::Quality>(d->quality), threads);
	if (convert_seconds)
		*convert_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (!ok)
		return -2;
	const std::vector<uint8_t>& data = out[0][0][0];
	if (data.size() > dst_size) return -1;
	std::memcpy(dst, data.data(), data.size());
	return static_cast<int>(data.size());
}

extern "C" int cfglue_encode(const cfglue_desc* d, const float* src, size_t pitch_floats, uint8_t* dst, size_t dst_size, unsigned threads)
{
	return cfglue_encode_timed(d, src, pitch_floats, dst, dst_size, threads, nullptr);
}
