// Image::resize() for RGBAF images and the mip chain of Texture::generateMipmaps(), on the GPU.
//
// What the reference does (lib/src/Image.cpp:1324-1379 -> FreeImage_Rescale, lib/FreeImage/Source/
// FreeImageToolkit/Rescale.cpp:30-95, Resize.cpp:140-218 weights, :232-506 pass order, :1236-1273 and :2070-2112
// the float passes): a separable two-pass filter. Per destination index a window of source indices and
// double-precision weights (filter stretched by the scale when minifying, normalised to sum 1, trailing zero
// weights dropped); each pass accumulates weight*(double)texel in window order in double precision and rounds
// the sum to float; the pass along x runs first when the width does not grow, else the pass along y. The image
// is stored bottom-up, so the y tables are built for bottom-up row indices. sRGB images are taken to linear
// before and back after (Image.cpp:1337-1344, Color.h:224-242).
//
// Here: the window tables are built on the host exactly as above (cheap: one row per destination index), the two
// passes are one kernel each with one thread per destination texel, and the accumulation uses explicit
// __dmul_rn/__dadd_rn so that no fused multiply-add changes a rounding: the result is bit-identical to
// FreeImage's for linear images. Both passes are HBM-bound (16 B/texel read, 16 B/texel written).
#include "common.cuh"
#include "../../include/cfx.h"

#include <cmath>
#include <vector>

namespace cfx {

namespace {

double filter_support(uint32_t f)
{
    return f == CFX_FILTER_BOX ? 0.5 : (f == CFX_FILTER_LINEAR ? 1.0 : 2.0);      // Filters.h: the constructors
}

// The filters' impulse responses, Filters.h:69-257.
double filter_eval(uint32_t f, double v)
{
    switch (f) {
    case CFX_FILTER_BOX:
        return std::fabs(v) <= 0.5 ? 1.0 : 0.0;
    case CFX_FILTER_LINEAR:
        v = std::fabs(v);
        return v < 1.0 ? 1.0 - v : 0.0;
    case CFX_FILTER_CUBIC: {
        // Mitchell-Netravali with b = c = 1/3
        const double b = 1/static_cast<double>(3), c = 1/static_cast<double>(3);
        const double p0 = (6 - 2*b)/6, p2 = (-18 + 12*b + 6*c)/6, p3 = (12 - 9*b - 6*c)/6;
        const double q0 = (8*b + 24*c)/6, q1 = (-12*b - 48*c)/6, q2 = (6*b + 30*c)/6, q3 = (-b - 6*c)/6;
        v = std::fabs(v);
        if (v < 1) return p0 + v*v*(p2 + v*p3);
        if (v < 2) return q0 + v*(q1 + v*(q2 + v*q3));
        return 0;
    }
    case CFX_FILTER_CATMULL_ROM:
        if (v < -2) return 0;
        if (v < -1) return 0.5*(4 + v*(8 + v*(5 + v)));
        if (v < 0) return 0.5*(2 + v*v*(-5 - 3*v));
        if (v < 1) return 0.5*(2 + v*v*(-5 + 3*v));
        if (v < 2) return 0.5*(4 + v*(-8 + v*(5 - v)));
        return 0;
    default: {   // CFX_FILTER_BSPLINE
        v = std::fabs(v);
        if (v < 1) return (4 + v*v*(-6 + 3*v))/6;
        if (v < 2) { const double t = 2 - v; return t*t*t/6; }
        return 0;
    }
    }
}

struct WindowTable {
    int window = 0;                 // doubles per destination index
    std::vector<int2> span;         // (first source index, tap count)
    std::vector<double> weight;     // [dst][window]
};

// Resize.cpp:140-218
void build_windows(uint32_t filter, uint32_t dst_size, uint32_t src_size, WindowTable& t)
{
    const double support = filter_support(filter);
    const double scale = static_cast<double>(dst_size)/static_cast<double>(src_size);
    double width, fscale;
    if (scale < 1.0) { width = support/scale; fscale = scale; } else { width = support; fscale = 1.0; }
    t.window = 2*static_cast<int>(std::ceil(width)) + 1;
    t.span.assign(dst_size, make_int2(0, 0));
    t.weight.assign(static_cast<size_t>(dst_size)*t.window, 0.0);
    const double offset = 0.5/scale;
    for (uint32_t u = 0; u < dst_size; ++u) {
        const double center = static_cast<double>(u)/scale + offset;
        const int left = std::max(0, static_cast<int>(center - width + 0.5));
        int right = std::min(static_cast<int>(center + width + 0.5), static_cast<int>(src_size));
        double* w = t.weight.data() + static_cast<size_t>(u)*t.window;
        double total = 0;
        for (int i = left; i < right; ++i) {
            const double wi = fscale*filter_eval(filter, fscale*(static_cast<double>(i) + 0.5 - center));
            w[i - left] = wi;
            total += wi;
        }
        if (total > 0 && total != 1)
            for (int i = left; i < right; ++i) w[i - left] /= total;
        // trailing zero weights are dropped (the window never becomes empty)
        int trailing = right - left - 1;
        while (trailing >= 0 && w[trailing] == 0) {
            --right; --trailing;
            if (right == left) break;
        }
        t.span[u] = make_int2(left, right - left);
    }
}

__device__ __forceinline__ double srgb_to_linear(double c)      // Color.h:224-229
{
    return c <= 0.04045 ? c/12.92 : pow((c + 0.055)/1.055, 2.4);
}

__device__ __forceinline__ double linear_to_srgb(double c)      // Color.h:236-242
{
    return c <= 0.0031308 ? c*12.92 : 1.055*pow(c, 1.0/2.4) - 0.055;
}

// One pass. ALONG_X: dst(x, y) = sum_k w[x][k] * src(left[x] + k, y). Else along y with the tables indexed by the
// bottom-up row number: dst row y is table entry dh-1-y, and its k-th tap is source row sh-1-(left+k).
// SRGB_IN converts each tap to linear on the fly (first pass), SRGB_OUT the result back (last pass).
// SRC_U8: the source is an RGBA8 surface read as (float)v/255.0F, FreeImage_ConvertToRGBAF's conversion
// (lib/FreeImage/Source/FreeImage/ConversionRGBAF.cpp:116-119) -- what Image::convert(RGBAF) stores for an 8-bit image.
// The 256 possible values are tabulated once per CTA, already widened to double (and taken to linear for sRGB), so a
// tap costs four shared-memory loads instead of four divisions and four conversions.
// One CTA row per image row; rows beyond the 65535 limit of gridDim.y continue in gridDim.z.
constexpr uint32_t kMaxGridY = 65535;
static inline dim3 rows_grid(uint32_t w, uint32_t h) { return dim3((w + 255)/256, h < kMaxGridY ? h : kMaxGridY, (h + kMaxGridY - 1)/kMaxGridY); }

template <bool ALONG_X, bool SRGB_IN, bool SRGB_OUT, bool SRC_U8>
__global__ void __launch_bounds__(256) resize_pass_kernel(const uint8_t* __restrict__ src, size_t src_pitch, uint32_t src_h,
    float4* __restrict__ dst, size_t dst_pitch4, uint32_t dw, uint32_t dh,
    const int2* __restrict__ span, const double* __restrict__ weight, int window)
{
    __shared__ double lut[SRC_U8 ? 512 : 1];           // [0..255] colour channels, [256..511] alpha (never sRGB)
    if (SRC_U8) {
        const float f = __fdiv_rn(static_cast<float>(threadIdx.x), 255.0f);
        lut[threadIdx.x] = SRGB_IN ? static_cast<double>(static_cast<float>(srgb_to_linear(f))) : static_cast<double>(f);
        lut[256 + threadIdx.x] = static_cast<double>(f);
        __syncthreads();
    }
    const uint32_t x = blockIdx.x*blockDim.x + threadIdx.x, y = blockIdx.z*kMaxGridY + blockIdx.y;
    if (x >= dw || y >= dh) return;
    const uint32_t u = ALONG_X ? x : dh - 1u - y;
    const int2 sp = span[u];
    const double* w = weight + static_cast<size_t>(u)*window;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int k = 0; k < sp.y; ++k) {
        const uint32_t i = static_cast<uint32_t>(sp.x + k);
        const uint32_t tx = ALONG_X ? i : x, ty = ALONG_X ? y : src_h - 1u - i;
        const double wk = w[k];
        double c0, c1, c2, c3;
        if (SRC_U8) {
            const uchar4 v = *reinterpret_cast<const uchar4*>(src + ty*src_pitch + static_cast<size_t>(tx)*4u);
            c0 = lut[v.x]; c1 = lut[v.y]; c2 = lut[v.z]; c3 = lut[256 + v.w];
        } else {
            const float4 v = *reinterpret_cast<const float4*>(src + ty*src_pitch + static_cast<size_t>(tx)*16u);
            if (SRGB_IN) {
                c0 = static_cast<float>(srgb_to_linear(v.x)); c1 = static_cast<float>(srgb_to_linear(v.y));
                c2 = static_cast<float>(srgb_to_linear(v.z));
            } else {
                c0 = v.x; c1 = v.y; c2 = v.z;
            }
            c3 = v.w;
        }
        a0 = __dadd_rn(a0, __dmul_rn(wk, c0));
        a1 = __dadd_rn(a1, __dmul_rn(wk, c1));
        a2 = __dadd_rn(a2, __dmul_rn(wk, c2));
        a3 = __dadd_rn(a3, __dmul_rn(wk, c3));
    }
    float4 o = make_float4(static_cast<float>(a0), static_cast<float>(a1), static_cast<float>(a2), static_cast<float>(a3));
    if (SRGB_OUT) {
        o.x = static_cast<float>(linear_to_srgb(o.x)); o.y = static_cast<float>(linear_to_srgb(o.y));
        o.z = static_cast<float>(linear_to_srgb(o.z));
    }
    dst[y*dst_pitch4 + x] = o;
}

struct DeviceTable {
    int2* span = nullptr;
    double* weight = nullptr;
    int window = 0;
};

int upload_table(const WindowTable& t, uint8_t* scratch, size_t& used, size_t cap, DeviceTable& d, cudaStream_t s)
{
    const size_t span_bytes = (t.span.size()*sizeof(int2) + 15) & ~static_cast<size_t>(15);
    const size_t w_bytes = (t.weight.size()*sizeof(double) + 15) & ~static_cast<size_t>(15);
    if (used + span_bytes + w_bytes > cap) return CFX_ERR_INVALID;
    d.span = reinterpret_cast<int2*>(scratch + used);
    d.weight = reinterpret_cast<double*>(scratch + used + span_bytes);
    d.window = t.window;
    if (cudaMemcpyAsync(d.span, t.span.data(), t.span.size()*sizeof(int2), cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(d.weight, t.weight.data(), t.weight.size()*sizeof(double), cudaMemcpyHostToDevice, s) != cudaSuccess)
        return CFX_ERR_CUDA;
    used += span_bytes + w_bytes;
    return CFX_OK;
}

template <bool ALONG_X, bool SRC_U8>
void launch_pass(bool srgb_in, bool srgb_out, const uint8_t* src, size_t sp, uint32_t sh, float4* dst, size_t dp4,
    uint32_t dw, uint32_t dh, const DeviceTable& t, cudaStream_t s)
{
    const dim3 grid = rows_grid(dw, dh), block(256);
    if (srgb_in && srgb_out)
        resize_pass_kernel<ALONG_X, true, true, SRC_U8><<<grid, block, 0, s>>>(src, sp, sh, dst, dp4, dw, dh, t.span, t.weight, t.window);
    else if (srgb_in)
        resize_pass_kernel<ALONG_X, true, false, SRC_U8><<<grid, block, 0, s>>>(src, sp, sh, dst, dp4, dw, dh, t.span, t.weight, t.window);
    else if (srgb_out)
        resize_pass_kernel<ALONG_X, false, true, SRC_U8><<<grid, block, 0, s>>>(src, sp, sh, dst, dp4, dw, dh, t.span, t.weight, t.window);
    else
        resize_pass_kernel<ALONG_X, false, false, SRC_U8><<<grid, block, 0, s>>>(src, sp, sh, dst, dp4, dw, dh, t.span, t.weight, t.window);
}

template <bool ALONG_X>
void launch_pass(bool src_u8, bool srgb_in, bool srgb_out, const uint8_t* src, size_t sp, uint32_t sh, float4* dst, size_t dp4,
    uint32_t dw, uint32_t dh, const DeviceTable& t, cudaStream_t s)
{
    if (src_u8) launch_pass<ALONG_X, true>(srgb_in, srgb_out, src, sp, sh, dst, dp4, dw, dh, t, s);
    else launch_pass<ALONG_X, false>(srgb_in, srgb_out, src, sp, sh, dst, dp4, dw, dh, t, s);
}

// a same-size "resize" of an RGBA8 source: only the conversion
__global__ void __launch_bounds__(256) widen_u8_kernel(const uint8_t* __restrict__ src, size_t src_pitch, float4* __restrict__ dst,
    size_t dst_pitch4, uint32_t w, uint32_t h)
{
    const uint32_t x = blockIdx.x*blockDim.x + threadIdx.x, y = blockIdx.z*kMaxGridY + blockIdx.y;
    if (x < w && y < h) {
        const uchar4 v = *reinterpret_cast<const uchar4*>(src + y*src_pitch + static_cast<size_t>(x)*4u);
        dst[y*dst_pitch4 + x] = make_float4(__fdiv_rn(static_cast<float>(v.x), 255.0f), __fdiv_rn(static_cast<float>(v.y), 255.0f),
            __fdiv_rn(static_cast<float>(v.z), 255.0f), __fdiv_rn(static_cast<float>(v.w), 255.0f));
    }
}

} // namespace

size_t resize_scratch_bytes(uint32_t sw, uint32_t sh, uint32_t dw, uint32_t dh)
{
    // the intermediate image of the two-pass filter + both window tables (window <= 2*ceil(2*ratio)+1 doubles)
    const size_t tmp = static_cast<size_t>(dw <= sw ? dw : sw)*(dw <= sw ? sh : dh)*16u;
    auto table = [](uint32_t d, uint32_t s) {
        const double ratio = d < s ? static_cast<double>(s)/d : 1.0;
        const size_t window = 2*static_cast<size_t>(std::ceil(2.0*ratio)) + 1;
        return static_cast<size_t>(d)*(window*sizeof(double) + sizeof(int2)) + 64;
    };
    return ((tmp + 255) & ~static_cast<size_t>(255)) + table(dw, sw) + table(dh, sh);
}

// Resize one surface resident in device memory (RGBA32F, or RGBA8 when src_u8; pitches in bytes; the RGBA32F result
// has a 16-byte-multiple pitch) on `stream`. `scratch` must hold resize_scratch_bytes(). Returns the number of kernels
// launched, or a negative CFX error.
int resize_device(const uint8_t* src, size_t src_pitch, bool src_u8, uint32_t sw, uint32_t sh, uint8_t* dst, size_t dst_pitch,
    uint32_t dw, uint32_t dh, uint32_t filter, bool srgb, uint8_t* scratch, size_t scratch_cap, cudaStream_t stream)
{
    float4* d4 = reinterpret_cast<float4*>(dst);
    const size_t dp4 = dst_pitch/16;
    if (sw == dw && sh == dh) {                     // Image.cpp:1330-1334: a plain copy
        if (src_u8) {
            widen_u8_kernel<<<rows_grid(dw, dh), 256, 0, stream>>>(src, src_pitch, d4, dp4, dw, dh);
            return cudaGetLastError() == cudaSuccess ? 1 : CFX_ERR_CUDA;
        }
        if (cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, static_cast<size_t>(dw)*16u, dh, cudaMemcpyDeviceToDevice,
                stream) != cudaSuccess)
            return CFX_ERR_CUDA;
        return 0;
    }
    const bool x_first = dw <= sw;                  // Resize.cpp:371
    const bool need_x = sw != dw, need_y = sh != dh;
    const uint32_t tw = x_first ? dw : sw, th = x_first ? sh : dh;      // the intermediate image
    size_t used = (static_cast<size_t>(tw)*th*16u + 255) & ~static_cast<size_t>(255);
    if (used > scratch_cap) return CFX_ERR_INVALID;
    float4* tmp = reinterpret_cast<float4*>(scratch);
    const uint8_t* tmp8 = scratch;
    const size_t tp = static_cast<size_t>(tw)*16u;
    DeviceTable tx, ty;
    WindowTable hx, hy;
    if (need_x) {
        build_windows(filter, dw, sw, hx);
        int rc = upload_table(hx, scratch, used, scratch_cap, tx, stream);
        if (rc != CFX_OK) return rc;
    }
    if (need_y) {
        build_windows(filter, dh, sh, hy);
        int rc = upload_table(hy, scratch, used, scratch_cap, ty, stream);
        if (rc != CFX_OK) return rc;
    }
    int launches = 0;
    if (need_x && need_y) {
        if (x_first) {
            launch_pass<true>(src_u8, srgb, false, src, src_pitch, sh, tmp, tw, dw, sh, tx, stream);
            launch_pass<false>(false, false, srgb, tmp8, tp, sh, d4, dp4, dw, dh, ty, stream);
        } else {
            launch_pass<false>(src_u8, srgb, false, src, src_pitch, sh, tmp, tw, sw, dh, ty, stream);
            launch_pass<true>(false, false, srgb, tmp8, tp, dh, d4, dp4, dw, dh, tx, stream);
        }
        launches = 2;
    } else if (need_x) {
        launch_pass<true>(src_u8, srgb, srgb, src, src_pitch, sh, d4, dp4, dw, dh, tx, stream);
        launches = 1;
    } else {
        launch_pass<false>(src_u8, srgb, srgb, src, src_pitch, sh, d4, dp4, dw, dh, ty, stream);
        launches = 1;
    }
    // the host tables are pageable: the copies above have been staged by the time cudaMemcpyAsync returned
    if (cudaGetLastError() != cudaSuccess) return CFX_ERR_CUDA;
    return launches;
}

} // namespace cfx
