// See CudaConverter.h.
#include "CudaConverter.h"

#include <cfx.h>

#include <atomic>
#include <cstddef>
#include <cstdint>

namespace cuttlefish
{

static std::atomic<unsigned int> g_surfaces(0);

std::unique_ptr<Converter> CudaConverter::create(const Texture& texture, const Image& image,
	Texture::Quality quality, unsigned int)
{
	// The enum values of cfx.h are the numeric values of Texture::Format/Type/Quality/Alpha
	// (lib/include/cuttlefish/Texture.h:59-188): cast, never map.
	cfx_surface_desc desc = {};
	desc.format = static_cast<uint32_t>(texture.format());
	desc.type = static_cast<uint32_t>(texture.type());
	if (!cfx_format_supported(desc.format, desc.type))
		return nullptr;

	desc.quality = static_cast<uint32_t>(quality);
	desc.alpha_type = static_cast<uint32_t>(texture.alphaType());
	Texture::ColorMask mask = texture.colorMask();
	desc.color_mask = (mask.r ? 1u : 0u) | (mask.g ? 2u : 0u) | (mask.b ? 4u : 0u) | (mask.a ? 8u : 0u);
	desc.color_space = image.colorSpace() == ColorSpace::sRGB ? 1u : 0u;
	desc.width = image.width();
	desc.height = image.height();
	desc.src_format = CFX_SRC_RGBA32F;        // Converter images are always Image::Format::RGBAF (Converter.h:52-56)

	// Image keeps its rows bottom-up (FreeImage; lib/src/Image.cpp:340-343) and scanline(y) counts from the top
	// (lib/src/Image.cpp:1092-1098). Hand the pixels over as they lie: the lowest address is whichever of the first
	// and last scanline comes first in memory, and CFX_FLAG_BOTTOM_UP tells the library which way the rows run.
	const std::uint8_t* first = static_cast<const std::uint8_t*>(image.scanline(0));
	const std::uint8_t* src = first;
	std::size_t pitch = std::size_t(desc.width)*sizeof(float)*4;
	if (desc.height > 1)
	{
		const std::uint8_t* second = static_cast<const std::uint8_t*>(image.scanline(1));
		if (second < first)
		{
			pitch = std::size_t(first - second);
			src = static_cast<const std::uint8_t*>(image.scanline(desc.height - 1));
			desc.flags |= CFX_FLAG_BOTTOM_UP;
		}
		else
			pitch = std::size_t(second - first);
	}
	desc.src_row_pitch = pitch;

	std::unique_ptr<CudaConverter> converter(new CudaConverter(image));
	converter->data().resize(cfx_encoded_size(&desc));
	// cfx_encode() fans the surface out over the device pool (cfx_init_devices(0) = every GPU of the box); without an
	// explicit init it uses the current CUDA device.
	if (cfx_encode(&desc, src, converter->data().data(), converter->data().size()) != CFX_OK)
		return nullptr;
	++g_surfaces;
	return std::unique_ptr<Converter>(converter.release());
}

void CudaConverter::process(unsigned int, unsigned int, ThreadData*)
{
	// the blocks were produced in create()
}

} // namespace cuttlefish

extern "C" unsigned int cfx_adapter_surfaces_encoded(void)
{
	return cuttlefish::g_surfaces.load();
}
