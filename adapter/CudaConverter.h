// Reference-side adapter: a cuttlefish::Converter that hands a whole surface to libcfx.so (include/cfx.h).
//
// This is the file a Cuttlefish maintainer adds to lib/src/ (with CudaConverter.cpp), compiled when
// CUTTLEFISH_HAS_CUDA is set. It is written against the reference's own internal interface,
// lib/src/Converter.h:31-76, and follows the pattern the reference already uses for an external whole-image
// encoder: PvrtcConverter reports jobsX() == jobsY() == 1 (lib/src/PvrtcConverter.h:36-37,
// lib/src/PvrtcConverter.cpp:50-129) and Converter::convert() runs a one-job converter inline on the calling
// thread (lib/src/Converter.cpp:548-554).
//
// The hook is one statement at the top of createConverter() (lib/src/Converter.cpp:32-37):
//     if (auto gpu = CudaConverter::create(texture, image, quality, threadCount)) return gpu;
// create() returns nullptr -- and createConverter() falls through to the stock CPU converters -- when the
// (format, type) pair has no GPU encoder, when no sm_100 device is usable, or when the encode fails; that
// fallback lives in Cuttlefish, libcfx.so has none.
//
// oracle/Makefile compiles this file together with the reference's real Converter.cpp (the hook injected into a
// temporary copy) into oracle/_ref/libcfglue_cuda.so; tests/test_adapter_gpu.py drives Converter::convert()
// through it and requires the bytes cfx_encode() gives directly.
#pragma once

#include "Converter.h"

namespace cuttlefish
{

class CudaConverter : public Converter
{
public:
	// Encodes the surface right away (so that a failure can still fall back) and returns the finished converter.
	static std::unique_ptr<Converter> create(const Texture& texture, const Image& image,
		Texture::Quality quality, unsigned int threadCount);

	unsigned int jobsX() const override {return 1;}
	unsigned int jobsY() const override {return 1;}
	void process(unsigned int x, unsigned int y, ThreadData* threadData) override;

private:
	explicit CudaConverter(const Image& image) : Converter(image) {}
};

} // namespace cuttlefish

// How many surfaces went through the GPU since the library was loaded (tests use it to prove the path taken).
extern "C" unsigned int cfx_adapter_surfaces_encoded(void);
