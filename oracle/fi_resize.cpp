// TEST INFRASTRUCTURE (oracle): the reference's Image::resize() for RGBAF images, run through the REAL
// FreeImage_Rescale compiled from /root/reference/lib/FreeImage/Source (Resize.cpp, Rescale.cpp, BitmapAccess.cpp,
// PixelAccess.cpp, where they lie), without the other 536 FreeImage translation units.
//
// The few FreeImage symbols those four files reference but never reach for an RGBAF rescale (metadata tags,
// 8/24/32-bit conversions, tone mapping, the message callback) are stubbed to abort(). cfref_resize_rgbaf()
// restates the ~20 lines of lib/src/Image.cpp:1324-1379 around the call: the image lives in a bottom-up
// FIT_RGBAF bitmap (Image::scanline(y) is FreeImage row height-1-y, Image.cpp:986-994), the ResizeFilter ->
// FREE_IMAGE_FILTER map of :1347-1365, and for sRGB images the changeColorSpace() round trip of :1337-1344 /
// :1666-1710 with Color.h:224-242's double-precision transfer functions.
#include "FreeImage.h"
#include "Utilities.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define FI_STUB(name) { fprintf(stderr, "oracle/fi_resize.cpp: unexpected call to " name "\n"); abort(); }

struct FITAG;
FIBITMAP* DLL_CALLCONV FreeImage_ConvertTo8Bits(FIBITMAP*) FI_STUB("FreeImage_ConvertTo8Bits")
FIBITMAP* DLL_CALLCONV FreeImage_ConvertTo24Bits(FIBITMAP*) FI_STUB("FreeImage_ConvertTo24Bits")
FIBITMAP* DLL_CALLCONV FreeImage_ConvertTo32Bits(FIBITMAP*) FI_STUB("FreeImage_ConvertTo32Bits")
FIBITMAP* DLL_CALLCONV FreeImage_ConvertToGreyscale(FIBITMAP*) FI_STUB("FreeImage_ConvertToGreyscale")
FIBITMAP* DLL_CALLCONV FreeImage_ConvertToRGBF(FIBITMAP*) FI_STUB("FreeImage_ConvertToRGBF")
FIBITMAP* DLL_CALLCONV FreeImage_ConvertToStandardType(FIBITMAP*, BOOL) FI_STUB("FreeImage_ConvertToStandardType")
FIBITMAP* DLL_CALLCONV FreeImage_Copy(FIBITMAP*, int, int, int, int) FI_STUB("FreeImage_Copy")
FIBITMAP* DLL_CALLCONV FreeImage_ToneMapping(FIBITMAP*, FREE_IMAGE_TMO, double, double) FI_STUB("FreeImage_ToneMapping")
FITAG* DLL_CALLCONV FreeImage_CreateTag() FI_STUB("FreeImage_CreateTag")
void DLL_CALLCONV FreeImage_DeleteTag(FITAG*) FI_STUB("FreeImage_DeleteTag")
FITAG* DLL_CALLCONV FreeImage_CloneTag(FITAG*) FI_STUB("FreeImage_CloneTag")
const char* DLL_CALLCONV FreeImage_GetTagKey(FITAG*) FI_STUB("FreeImage_GetTagKey")
FREE_IMAGE_MDTYPE DLL_CALLCONV FreeImage_GetTagType(FITAG*) FI_STUB("FreeImage_GetTagType")
DWORD DLL_CALLCONV FreeImage_GetTagCount(FITAG*) FI_STUB("FreeImage_GetTagCount")
DWORD DLL_CALLCONV FreeImage_GetTagLength(FITAG*) FI_STUB("FreeImage_GetTagLength")
const void* DLL_CALLCONV FreeImage_GetTagValue(FITAG*) FI_STUB("FreeImage_GetTagValue")
BOOL DLL_CALLCONV FreeImage_SetTagKey(FITAG*, const char*) FI_STUB("FreeImage_SetTagKey")
BOOL DLL_CALLCONV FreeImage_SetTagID(FITAG*, WORD) FI_STUB("FreeImage_SetTagID")
BOOL DLL_CALLCONV FreeImage_SetTagType(FITAG*, FREE_IMAGE_MDTYPE) FI_STUB("FreeImage_SetTagType")
BOOL DLL_CALLCONV FreeImage_SetTagCount(FITAG*, DWORD) FI_STUB("FreeImage_SetTagCount")
BOOL DLL_CALLCONV FreeImage_SetTagLength(FITAG*, DWORD) FI_STUB("FreeImage_SetTagLength")
BOOL DLL_CALLCONV FreeImage_SetTagValue(FITAG*, const void*) FI_STUB("FreeImage_SetTagValue")
unsigned FreeImage_TagDataWidth(FREE_IMAGE_MDTYPE) FI_STUB("FreeImage_TagDataWidth")
size_t FreeImage_GetTagMemorySize(FITAG*) FI_STUB("FreeImage_GetTagMemorySize")
void FreeImage_OutputMessageProc(int, const char* fmt, ...) { fprintf(stderr, "FreeImage: %s\n", fmt); }

// BitmapAccess.cpp asks TagLib for tag ids when metadata is set by key; an RGBAF rescale never does.
#include "../Metadata/FreeImageTag.h"
TagLib& TagLib::instance() FI_STUB("TagLib::instance")
int TagLib::getTagID(MDMODEL, const char*) FI_STUB("TagLib::getTagID")
TagLib::TagLib() {}
TagLib::~TagLib() {}

static double srgb_to_linear(double c) { return c <= 0.04045 ? c/12.92 : std::pow((c + 0.055)/1.055, 2.4); }
static double linear_to_srgb(double c) { return c <= 0.0031308 ? c*12.92 : 1.055*std::pow(c, 1.0/2.4) - 0.055; }

static void change_color_space(FIBITMAP* img, bool to_linear)
{
    const unsigned w = FreeImage_GetWidth(img), h = FreeImage_GetHeight(img);
    for (unsigned y = 0; y < h; ++y) {
        float* row = reinterpret_cast<float*>(FreeImage_GetScanLine(img, y));
        for (unsigned x = 0; x < w; ++x)
            for (int c = 0; c < 3; ++c) {
                const double v = row[4*x + c];
                row[4*x + c] = static_cast<float>(to_linear ? srgb_to_linear(v) : linear_to_srgb(v));
            }
    }
}

// src/dst: RGBA32F, rows top-down, tightly packed. filter: cuttlefish::Image::ResizeFilter. Returns 0 on success.
extern "C" int cfref_resize_rgbaf(const float* src, int sw, int sh, float* dst, int dw, int dh, int filter, int srgb)
{
    static const FREE_IMAGE_FILTER map[5] = {FILTER_BOX, FILTER_BILINEAR, FILTER_BICUBIC, FILTER_CATMULLROM, FILTER_BSPLINE};
    if (filter < 0 || filter > 4 || sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0) return -1;
    if (sw == dw && sh == dh) { memcpy(dst, src, static_cast<size_t>(sw)*sh*16); return 0; }
    FIBITMAP* in = FreeImage_AllocateT(FIT_RGBAF, sw, sh, 128, 0, 0, 0);
    if (!in) return -2;
    for (int y = 0; y < sh; ++y)
        memcpy(FreeImage_GetScanLine(in, sh - 1 - y), src + static_cast<size_t>(y)*sw*4, static_cast<size_t>(sw)*16);
    if (srgb) change_color_space(in, true);
    FIBITMAP* out = FreeImage_Rescale(in, dw, dh, map[filter]);
    FreeImage_Unload(in);
    if (!out) return -3;
    if (srgb) change_color_space(out, false);
    for (int y = 0; y < dh; ++y)
        memcpy(dst + static_cast<size_t>(y)*dw*4, FreeImage_GetScanLine(out, dh - 1 - y), static_cast<size_t>(dw)*16);
    FreeImage_Unload(out);
    return 0;
}
