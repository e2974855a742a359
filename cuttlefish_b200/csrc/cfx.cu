// C-ABI of the encoder library (include/cfx.h): descriptor validation, device context
// (streams, device and pinned buffers), the chunked H2D -> kernel -> D2H pipeline of
// cfx_encode(), and dispatch to the per-format kernel launchers.
//
// This file is the device-side stand-in for Converter::convert()'s per-surface job loop
// (lib/src/Converter.cpp:508-593): where the reference enumerates jobsX x jobsY process(x,y)
// calls over a std::thread pool, this enqueues one persistent kernel per chunk of block rows.
#include "../../include/cfx.h"
#include "kernels.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

namespace cfx {

static thread_local char t_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
    return code;
}

#define CFX_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(CFX_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));         \
    } while (0)

static const uint32_t kAstcDims[14][2] = {{4,4},{5,4},{5,5},{6,5},{6,6},{8,5},{8,6},{8,8},{10,5},
    {10,6},{10,8},{10,10},{12,10},{12,12}};

static bool block_info(uint32_t format, uint32_t& bw, uint32_t& bh, uint32_t& bytes)
{
    bw = bh = 4;
    switch (format) {
        case CFX_FORMAT_BC1_RGB: case CFX_FORMAT_BC1_RGBA: case CFX_FORMAT_BC4: case CFX_FORMAT_ETC1:
        case CFX_FORMAT_ETC2_R8G8B8: case CFX_FORMAT_ETC2_R8G8B8A1: case CFX_FORMAT_EAC_R11:
            bytes = 8; return true;
        case CFX_FORMAT_BC2: case CFX_FORMAT_BC3: case CFX_FORMAT_BC5: case CFX_FORMAT_BC6H:
        case CFX_FORMAT_BC7: case CFX_FORMAT_ETC2_R8G8B8A8: case CFX_FORMAT_EAC_R11G11:
            bytes = 16; return true;
        default:
            if (format >= CFX_FORMAT_ASTC_4x4 && format <= CFX_FORMAT_ASTC_12x12) {
                bw = kAstcDims[format - CFX_FORMAT_ASTC_4x4][0];
                bh = kAstcDims[format - CFX_FORMAT_ASTC_4x4][1];
                bytes = 16;
                return true;
            }
            return false;
    }
}

typedef int (*Launcher)(const EncodeParams&, cudaStream_t);

// (format,type) -> launcher. Mirrors the compressed cases of createConverter(),
// lib/src/Converter.cpp:339-502; anything absent here is CFX_ERR_UNSUPPORTED.
static Launcher find_launcher(uint32_t format, uint32_t type)
{
    switch (format) {
        case CFX_FORMAT_BC4: case CFX_FORMAT_BC5:
            return (type == CFX_TYPE_UNORM || type == CFX_TYPE_SNORM) ? launch_bc45 : nullptr;
#ifdef CFX_HAVE_BC7
        case CFX_FORMAT_BC7:
            return type == CFX_TYPE_UNORM ? launch_bc7 : nullptr;
#endif
#ifdef CFX_HAVE_BC1
        case CFX_FORMAT_BC1_RGB: case CFX_FORMAT_BC1_RGBA: case CFX_FORMAT_BC2: case CFX_FORMAT_BC3:
            return type == CFX_TYPE_UNORM ? launch_bc123 : nullptr;
#endif
#ifdef CFX_HAVE_ETC
        case CFX_FORMAT_ETC1: case CFX_FORMAT_ETC2_R8G8B8: case CFX_FORMAT_ETC2_R8G8B8A1: case CFX_FORMAT_ETC2_R8G8B8A8:
            return type == CFX_TYPE_UNORM ? launch_etc : nullptr;
        case CFX_FORMAT_EAC_R11: case CFX_FORMAT_EAC_R11G11:
            return (type == CFX_TYPE_UNORM || type == CFX_TYPE_SNORM) ? launch_etc : nullptr;
#endif
#ifdef CFX_HAVE_BC6H
        case CFX_FORMAT_BC6H:
            // Type::Float (signed): checked against a spec decoder (tests/util.py decode_bc6h) -- the reference's own
            // signed output (Compressonator) is not a valid encoding of its input, see DESIGN.md
            return (type == CFX_TYPE_UFLOAT || type == CFX_TYPE_FLOAT) ? launch_bc6h : nullptr;
#endif
        default:
#ifdef CFX_HAVE_ASTC
            // all 14 footprints: UNorm = LDR profile, UFloat = HDR profile (colour in end point mode 11, opaque alpha)
            if (format >= CFX_FORMAT_ASTC_4x4 && format <= CFX_FORMAT_ASTC_12x12) {
                const uint32_t* d = kAstcDims[format - CFX_FORMAT_ASTC_4x4];
                (void)d;
                return (type == CFX_TYPE_UNORM || type == CFX_TYPE_UFLOAT) ? launch_astc : nullptr;
            }
#endif
            return nullptr;
    }
}

static uint32_t src_texel_bytes(uint32_t src_format)
{
    return src_format == CFX_SRC_RGBA8 ? 4u : src_format == CFX_SRC_RGBA16F ? 8u : 16u;
}

// ---- device context ---------------------------------------------------------------------------

constexpr int kStreams = 3;
constexpr int kMaxMipLevels = 32;

struct Context {
    std::mutex mutex;
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    cudaStream_t streams[kStreams] = {};
    uint8_t* d_src = nullptr; size_t d_src_cap = 0;
    uint8_t* d_dst = nullptr; size_t d_dst_cap = 0;
    uint8_t* d_mip = nullptr; size_t d_mip_cap = 0;       // resized surfaces + the resize scratch
    cudaEvent_t fork[kMaxMipLevels] = {};                 // "level k is filtered" / "stream i has encoded its levels"
    cudaEvent_t join[kStreams] = {};
    cudaEvent_t uploaded[kStreams] = {};                  // "this stream's share of a surface is in HBM"
};
static Context g_ctx;

uint32_t persistent_ctas(const void* kernel, int threads, size_t dyn_smem)
{
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem) != cudaSuccess ||
        per_sm < 1)
        per_sm = 1;
    int sms = g_ctx.sm_count > 0 ? g_ctx.sm_count : 148;
    return static_cast<uint32_t>(sms*per_sm);
}

static int ensure_init(int device)
{
    if (g_ctx.ready && (device < 0 || device == g_ctx.device)) {
        CFX_CUDA(cudaSetDevice(g_ctx.device));
        return CFX_OK;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(CFX_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
            e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= count) return fail(CFX_ERR_INVALID, "device %d out of range (%d devices)", device, count);
    if (g_ctx.ready) {
        // switching device: drop the old context's resources
        cudaSetDevice(g_ctx.device);
        for (auto& s : g_ctx.streams) if (s) { cudaStreamDestroy(s); s = nullptr; }
        for (auto& e : g_ctx.fork) if (e) { cudaEventDestroy(e); e = nullptr; }
        for (auto& e : g_ctx.join) if (e) { cudaEventDestroy(e); e = nullptr; }
        for (auto& e : g_ctx.uploaded) if (e) { cudaEventDestroy(e); e = nullptr; }
        if (g_ctx.d_src) cudaFree(g_ctx.d_src);
        if (g_ctx.d_dst) cudaFree(g_ctx.d_dst);
        if (g_ctx.d_mip) cudaFree(g_ctx.d_mip);
        g_ctx.d_src = g_ctx.d_dst = g_ctx.d_mip = nullptr; g_ctx.d_src_cap = g_ctx.d_dst_cap = g_ctx.d_mip_cap = 0;
        g_ctx.ready = false;
    }
    cudaDeviceProp prop;
    CFX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(CFX_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
            device, prop.major, prop.minor);
    CFX_CUDA(cudaSetDevice(device));
    for (auto& s : g_ctx.streams) CFX_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    for (auto& e : g_ctx.fork) CFX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : g_ctx.join) CFX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : g_ctx.uploaded) CFX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    g_ctx.device = device;
    g_ctx.sm_count = prop.multiProcessorCount;
    g_ctx.ready = true;
    return CFX_OK;
}

static int reserve(uint8_t*& ptr, size_t& cap, size_t bytes)
{
    if (bytes <= cap) return CFX_OK;
    if (ptr) { CFX_CUDA(cudaDeviceSynchronize()); CFX_CUDA(cudaFree(ptr)); ptr = nullptr; cap = 0; }
    size_t want = bytes + bytes/8 + 4096;
    CFX_CUDA(cudaMalloc(&ptr, want));
    cap = want;
    return CFX_OK;
}

static int validate(const cfx_surface_desc* d, EncodeParams& p, Launcher& launcher)
{
    if (!d) return fail(CFX_ERR_INVALID, "null descriptor");
    uint32_t bw, bh, bytes;
    if (!block_info(d->format, bw, bh, bytes))
        return fail(CFX_ERR_UNSUPPORTED, "format %u is not a block-compressed format", d->format);
    launcher = find_launcher(d->format, d->type);
    if (!launcher)
        return fail(CFX_ERR_UNSUPPORTED, "no GPU encoder for format %u type %u", d->format, d->type);
    if (d->width == 0 || d->height == 0) return fail(CFX_ERR_INVALID, "empty surface %ux%u", d->width, d->height);
    if (d->quality > CFX_QUALITY_HIGHEST) return fail(CFX_ERR_INVALID, "quality %u out of range", d->quality);
    if (d->alpha_type > CFX_ALPHA_ENCODED) return fail(CFX_ERR_INVALID, "alpha type %u out of range", d->alpha_type);
    if (d->src_format > CFX_SRC_RGBA32F) return fail(CFX_ERR_INVALID, "source format %u out of range", d->src_format);
    if (d->reserved != 0) return fail(CFX_ERR_INVALID, "reserved field must be 0");
    uint64_t min_pitch = static_cast<uint64_t>(d->width)*src_texel_bytes(d->src_format);
    if (d->src_row_pitch < min_pitch)
        return fail(CFX_ERR_INVALID, "row pitch %llu < %llu", (unsigned long long)d->src_row_pitch,
            (unsigned long long)min_pitch);
    if (d->src_row_pitch % src_texel_bytes(d->src_format))
        return fail(CFX_ERR_INVALID, "row pitch must be a multiple of the texel size");
    uint64_t bx = (d->width + bw - 1)/bw, by = (d->height + bh - 1)/bh;
    if (bx*by > 0x7FFFFFFFull) return fail(CFX_ERR_INVALID, "surface too large");
    memset(&p, 0, sizeof(p));
    p.pitch = d->src_row_pitch;
    p.width = d->width; p.height = d->height; p.src_format = d->src_format;
    p.blocks_x = static_cast<uint32_t>(bx); p.blocks_y = static_cast<uint32_t>(by);
    p.total_blocks = static_cast<uint32_t>(bx*by);
    p.block_w = bw; p.block_h = bh; p.block_bytes = bytes;
    p.format = d->format; p.type = d->type; p.quality = d->quality; p.alpha_type = d->alpha_type;
    p.color_mask = d->color_mask & 15u; p.color_space = d->color_space ? 1u : 0u;
    return CFX_OK;
}

static int launch(Launcher launcher, EncodeParams& p, cudaStream_t stream)
{
    p.aligned16 = ((reinterpret_cast<uintptr_t>(p.src) | p.pitch) & 15) == 0;
    int n = launcher(p, stream);
    if (n < 0) return n;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CFX_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    g_launches += static_cast<uint64_t>(n);
    return CFX_OK;
}

// src_off / dst_off: where this surface lives in the context's device buffers (a batch lays its surfaces out back to
// back and reserves once); sync = false leaves the copies and kernels queued on the context's streams.
// uploaded: if set, receives a bit per stream that carried a piece of the surface; g_ctx.uploaded[i] of those streams
// fires once that stream's last piece is in HBM (before its encode kernel).
static int encode_host(const cfx_surface_desc* desc, const void* src, void* dst, size_t dst_size, size_t src_off = 0,
    size_t dst_off = 0, bool reserve_and_sync = true, int first_stream = 0, uint32_t* uploaded = nullptr)
{
    EncodeParams p; Launcher launcher;
    int rc = validate(desc, p, launcher);
    if (rc != CFX_OK) return rc;
    if (!src || !dst) return fail(CFX_ERR_INVALID, "null buffer");
    size_t out_bytes = static_cast<size_t>(p.total_blocks)*p.block_bytes;
    if (dst_size < out_bytes) return fail(CFX_ERR_INVALID, "dst_size %zu < %zu", dst_size, out_bytes);
    rc = ensure_init(-1);
    if (rc != CFX_OK) return rc;

    const size_t row_bytes = static_cast<size_t>(p.width)*src_texel_bytes(p.src_format);
    const size_t d_pitch = (row_bytes + 255) & ~static_cast<size_t>(255);
    if (reserve_and_sync) {
        rc = reserve(g_ctx.d_src, g_ctx.d_src_cap, src_off + d_pitch*p.height);
        if (rc != CFX_OK) return rc;
        rc = reserve(g_ctx.d_dst, g_ctx.d_dst_cap, dst_off + out_bytes);
        if (rc != CFX_OK) return rc;
    }
    uint8_t* const d_src = g_ctx.d_src + src_off;
    uint8_t* const d_dst = g_ctx.d_dst + dst_off;

    // Chunk by block rows so that copy-in, encode and copy-out of neighbouring chunks overlap
    // on the three streams. ~8 M texels per chunk keeps every kernel a few full waves.
    uint32_t rows_per_chunk = p.blocks_y;
    {
        uint64_t texels_per_row = static_cast<uint64_t>(p.width)*p.block_h;
        uint64_t want = (8ull << 20)/(texels_per_row ? texels_per_row : 1);
        if (want < 1) want = 1;
        if (want < rows_per_chunk) rows_per_chunk = static_cast<uint32_t>(want);
    }
    int k = first_stream;
    for (uint32_t r0 = 0; r0 < p.blocks_y; r0 += rows_per_chunk, ++k) {
        uint32_t r1 = min(p.blocks_y, r0 + rows_per_chunk);
        uint32_t y0 = r0*p.block_h, y1 = min(p.height, r1*p.block_h);
        cudaStream_t s = g_ctx.streams[k % kStreams];
        CFX_CUDA(cudaMemcpy2DAsync(d_src + static_cast<size_t>(y0)*d_pitch, d_pitch,
            static_cast<const uint8_t*>(src) + static_cast<size_t>(y0)*desc->src_row_pitch,
            desc->src_row_pitch, row_bytes, y1 - y0, cudaMemcpyHostToDevice, s));
        if (uploaded) {
            CFX_CUDA(cudaEventRecord(g_ctx.uploaded[k % kStreams], s));
            *uploaded |= 1u << (k % kStreams);
        }
        EncodeParams c = p;
        c.src = d_src + static_cast<size_t>(y0)*d_pitch;
        c.pitch = d_pitch;
        c.height = y1 - y0;   // interior chunks end on a block-row boundary, so the clamp is unchanged
        c.blocks_y = r1 - r0;
        c.total_blocks = c.blocks_x*c.blocks_y;
        size_t off = static_cast<size_t>(r0)*p.blocks_x*p.block_bytes;
        c.dst = d_dst + off;
        rc = launch(launcher, c, s);
        if (rc != CFX_OK) return rc;
        CFX_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(dst) + off, c.dst,
            static_cast<size_t>(c.total_blocks)*p.block_bytes, cudaMemcpyDeviceToHost, s));
    }
    if (reserve_and_sync)
        for (auto& s : g_ctx.streams) CFX_CUDA(cudaStreamSynchronize(s));
    return CFX_OK;
}

// resize.cu
size_t resize_scratch_bytes(uint32_t sw, uint32_t sh, uint32_t dw, uint32_t dh);
int resize_device(const uint8_t* src, size_t src_pitch, bool src_u8, uint32_t sw, uint32_t sh, uint8_t* dst, size_t dst_pitch,
    uint32_t dw, uint32_t dh, uint32_t filter, bool srgb, uint8_t* scratch, size_t scratch_cap, cudaStream_t stream);

static size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

static int resize_host(const void* src, uint32_t sw, uint32_t sh, size_t src_pitch, void* dst, uint32_t dw, uint32_t dh,
    size_t dst_pitch, uint32_t filter, uint32_t color_space)
{
    if (!src || !dst) return fail(CFX_ERR_INVALID, "null buffer");
    if (!sw || !sh || !dw || !dh) return fail(CFX_ERR_INVALID, "empty surface");
    if (filter > CFX_FILTER_BSPLINE) return fail(CFX_ERR_INVALID, "filter %u out of range", filter);
    if (src_pitch < static_cast<size_t>(sw)*16u || dst_pitch < static_cast<size_t>(dw)*16u || (src_pitch & 3) || (dst_pitch & 3))
        return fail(CFX_ERR_INVALID, "row pitch too small or not a multiple of 4");
    int rc = ensure_init(-1);
    if (rc != CFX_OK) return rc;
    const size_t sp = align256(static_cast<size_t>(sw)*16u), dp = align256(static_cast<size_t>(dw)*16u);
    const size_t scratch = resize_scratch_bytes(sw, sh, dw, dh);
    rc = reserve(g_ctx.d_src, g_ctx.d_src_cap, sp*sh);
    if (rc != CFX_OK) return rc;
    rc = reserve(g_ctx.d_mip, g_ctx.d_mip_cap, dp*dh + 256 + scratch);
    if (rc != CFX_OK) return rc;
    cudaStream_t s = g_ctx.streams[0];
    CFX_CUDA(cudaMemcpy2DAsync(g_ctx.d_src, sp, src, src_pitch, static_cast<size_t>(sw)*16u, sh, cudaMemcpyHostToDevice, s));
    uint8_t* d_out = g_ctx.d_mip;
    int n = resize_device(g_ctx.d_src, sp, false, sw, sh, d_out, dp, dw, dh, filter, color_space != 0, d_out + align256(dp*dh),
        scratch, s);
    if (n < 0) return fail(n, "resize failed: %s", cudaGetErrorString(cudaGetLastError()));
    g_launches += static_cast<uint64_t>(n);
    CFX_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, d_out, dp, static_cast<size_t>(dw)*16u, dh, cudaMemcpyDeviceToHost, s));
    CFX_CUDA(cudaStreamSynchronize(s));
    return CFX_OK;
}

static uint32_t mip_levels(uint32_t w, uint32_t h)
{
    uint32_t m = w > h ? w : h, n = 0;
    while (m) { ++n; m >>= 1; }
    return n;
}

// Validates a chain request and fills one descriptor per level (Texture::width/height(mip); levels clamped to
// [1, maxMipmapLevels] like Texture.cpp:1341).
static int mip_chain_descs(const cfx_surface_desc* level0, uint32_t filter, uint32_t levels, const size_t* dst_sizes,
    std::vector<cfx_surface_desc>& descs)
{
    EncodeParams p0; Launcher launcher;
    int rc = validate(level0, p0, launcher);
    if (rc != CFX_OK) return rc;
    if (level0->src_format != CFX_SRC_RGBA32F && level0->src_format != CFX_SRC_RGBA8)
        return fail(CFX_ERR_INVALID, "the mip chain is generated from an RGBA32F level 0 (Image::Format::RGBAF) or an RGBA8 one "
            "(taken as v/255, Image::convert(RGBAF) of an 8-bit image)");
    if (filter > CFX_FILTER_BSPLINE) return fail(CFX_ERR_INVALID, "filter %u out of range", filter);
    if (!dst_sizes) return fail(CFX_ERR_INVALID, "null buffer");
    if (levels < 1) levels = 1;
    const uint32_t max_levels = mip_levels(level0->width, level0->height);
    if (levels > max_levels) levels = max_levels;
    descs.assign(levels, *level0);
    for (uint32_t k = 0; k < levels; ++k) {
        descs[k].width = level0->width >> k ? level0->width >> k : 1u;
        descs[k].height = level0->height >> k ? level0->height >> k : 1u;
        if (k) { descs[k].src_format = CFX_SRC_RGBA32F; descs[k].src_row_pitch = align256(static_cast<size_t>(descs[k].width)*16u); }
        const size_t bytes = cfx_encoded_size(&descs[k]);
        if (dst_sizes[k] < bytes) return fail(CFX_ERR_INVALID, "level %u: dst_size %zu < %zu", k, dst_sizes[k], bytes);
    }
    return CFX_OK;
}

// Levels of a chain whose RGBA32F level 0 is resident at d_level0. The filter chain runs on stream s: each level is
// resized from the level above into the library's mip storage (every level keeps its own region until the next chain
// call). The encoders only depend on their own level, so they fork onto the context's other streams as soon as that
// level is filtered -- the launch-latency-bound tail levels then overlap each other and the big levels -- and join s at
// the end. Level 0 is encoded too when d_outs[0] is set. level_ptr[k] receives where level k's image lives.
static int run_mip_levels(const std::vector<cfx_surface_desc>& descs, const uint8_t* d_level0, size_t pitch0, uint32_t filter,
    uint8_t* const* d_outs, std::vector<const uint8_t*>& level_ptr, cudaStream_t s)
{
    const uint32_t levels = static_cast<uint32_t>(descs.size());
    level_ptr.assign(levels, nullptr);
    level_ptr[0] = d_level0;
    std::vector<size_t> off(levels + 1, 0);
    for (uint32_t k = 1; k < levels; ++k) off[k + 1] = off[k] + align256(descs[k].src_row_pitch*descs[k].height);
    const size_t scratch = levels > 1 ? resize_scratch_bytes(descs[0].width, descs[0].height, descs[1].width, descs[1].height) : 0;
    int rc = reserve(g_ctx.d_mip, g_ctx.d_mip_cap, off[levels] + scratch + 256);
    if (rc != CFX_OK) return rc;
    uint8_t* d_scratch = g_ctx.d_mip + off[levels];
    size_t prev_pitch = pitch0;
    bool used[kStreams] = {};
    for (uint32_t k = 0; k < levels; ++k) {
        const cfx_surface_desc& d = descs[k];
        if (k) {
            uint8_t* cur = g_ctx.d_mip + off[k];
            int n = resize_device(level_ptr[k - 1], prev_pitch, descs[k - 1].src_format == CFX_SRC_RGBA8, descs[k - 1].width,
                descs[k - 1].height, cur, d.src_row_pitch,
                d.width, d.height, filter, descs[0].color_space != 0, d_scratch, scratch, s);
            if (n < 0) return fail(n, "level %u: resize failed", k);
            g_launches += static_cast<uint64_t>(n);
            level_ptr[k] = cur;
            prev_pitch = d.src_row_pitch;
        }
        if (!d_outs[k]) continue;
        EncodeParams p; Launcher l;
        rc = validate(&d, p, l);
        if (rc != CFX_OK) return rc;
        p.src = level_ptr[k];
        p.pitch = k ? d.src_row_pitch : pitch0;
        p.dst = d_outs[k];
        // fork: one of the context's streams that is not s
        int a = static_cast<int>(k % kStreams);
        if (g_ctx.streams[a] == s) a = (a + 1) % kStreams;
        CFX_CUDA(cudaEventRecord(g_ctx.fork[k % kMaxMipLevels], s));
        CFX_CUDA(cudaStreamWaitEvent(g_ctx.streams[a], g_ctx.fork[k % kMaxMipLevels], 0));
        rc = launch(l, p, g_ctx.streams[a]);
        if (rc != CFX_OK) return rc;
        used[a] = true;
    }
    for (int a = 0; a < kStreams; ++a) {
        if (!used[a]) continue;
        CFX_CUDA(cudaEventRecord(g_ctx.join[a], g_ctx.streams[a]));
        CFX_CUDA(cudaStreamWaitEvent(s, g_ctx.join[a], 0));
    }
    return CFX_OK;
}

static int encode_mip_chain(const cfx_surface_desc* level0, const void* src, uint32_t filter, uint32_t levels,
    void* const* dsts, const size_t* dst_sizes, void* const* mip_images)
{
    std::vector<cfx_surface_desc> descs;
    int rc = mip_chain_descs(level0, filter, levels, dst_sizes, descs);
    if (rc != CFX_OK) return rc;
    levels = static_cast<uint32_t>(descs.size());
    if (!src || !dsts) return fail(CFX_ERR_INVALID, "null buffer");
    for (uint32_t k = 0; k < levels; ++k) if (!dsts[k]) return fail(CFX_ERR_INVALID, "level %u: null buffer", k);
    // Level 0 goes through the chunked upload + encode of cfx_encode(), which leaves the whole surface in d_src, but is
    // not awaited: the filter chain only needs the upload, so it starts on the least busy stream as soon as every piece
    // of level 0 is in HBM and runs beside level 0's encoders. One wait at the end.
    std::vector<size_t> bytes(levels, 0), off(levels + 1, 0);
    const size_t out0 = align256(cfx_encoded_size(&descs[0]));
    for (uint32_t k = 1; k < levels; ++k) { bytes[k] = cfx_encoded_size(&descs[k]); off[k + 1] = off[k] + align256(bytes[k]); }
    rc = ensure_init(-1);
    if (rc != CFX_OK) return rc;
    const size_t pitch0 = align256(static_cast<size_t>(level0->width)*src_texel_bytes(level0->src_format));
    rc = reserve(g_ctx.d_src, g_ctx.d_src_cap, pitch0*level0->height);
    if (rc != CFX_OK) return rc;
    rc = reserve(g_ctx.d_dst, g_ctx.d_dst_cap, out0 + off[levels]);
    if (rc != CFX_OK) return rc;
    uint32_t carriers = 0;
    rc = encode_host(level0, src, dsts[0], dst_sizes[0], 0, 0, false, 0, &carriers);
    if (rc == CFX_OK && levels > 1) {
        int pieces = 0;
        for (int i = 0; i < kStreams; ++i) pieces += (carriers >> i) & 1u;
        cudaStream_t s = g_ctx.streams[pieces % kStreams];          // the stream the next piece would have taken
        for (int i = 0; i < kStreams; ++i)
            if ((carriers >> i) & 1u) CFX_CUDA(cudaStreamWaitEvent(s, g_ctx.uploaded[i], 0));
        std::vector<uint8_t*> d_outs(levels, nullptr);
        for (uint32_t k = 1; k < levels; ++k) d_outs[k] = g_ctx.d_dst + out0 + off[k];
        std::vector<const uint8_t*> level_ptr;
        rc = run_mip_levels(descs, g_ctx.d_src, pitch0, filter, d_outs.data(), level_ptr, s);
        for (uint32_t k = 1; rc == CFX_OK && k < levels; ++k) {
            if (cudaMemcpyAsync(dsts[k], d_outs[k], bytes[k], cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = CFX_ERR_CUDA;
            if (rc == CFX_OK && mip_images && mip_images[k] &&
                cudaMemcpy2DAsync(mip_images[k], static_cast<size_t>(descs[k].width)*16u, level_ptr[k], descs[k].src_row_pitch,
                    static_cast<size_t>(descs[k].width)*16u, descs[k].height, cudaMemcpyDeviceToHost, s) != cudaSuccess)
                rc = CFX_ERR_CUDA;
        }
    }
    for (auto& st : g_ctx.streams) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess && rc == CFX_OK) rc = fail(CFX_ERR_CUDA, "%s", cudaGetErrorString(e));
    }
    if (rc == CFX_ERR_CUDA && !t_error[0]) fail(CFX_ERR_CUDA, "%s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

static int encode_mip_chain_device(const cfx_surface_desc* level0, const void* d_src, uint32_t filter, uint32_t levels,
    void* const* d_dsts, const size_t* dst_sizes, cudaStream_t s)
{
    std::vector<cfx_surface_desc> descs;
    int rc = mip_chain_descs(level0, filter, levels, dst_sizes, descs);
    if (rc != CFX_OK) return rc;
    levels = static_cast<uint32_t>(descs.size());
    if (!d_src || !d_dsts) return fail(CFX_ERR_INVALID, "null buffer");
    for (uint32_t k = 0; k < levels; ++k) if (!d_dsts[k]) return fail(CFX_ERR_INVALID, "level %u: null buffer", k);
    if ((reinterpret_cast<uintptr_t>(d_src) | level0->src_row_pitch) & (level0->src_format == CFX_SRC_RGBA8 ? 3 : 15))
        return fail(CFX_ERR_INVALID, "a device-resident level 0 must be texel aligned (16 bytes for RGBA32F, 4 for RGBA8)");
    rc = ensure_init(-1);
    if (rc != CFX_OK) return rc;
    std::vector<uint8_t*> d_outs(levels, nullptr);
    for (uint32_t k = 0; k < levels; ++k) d_outs[k] = static_cast<uint8_t*>(d_dsts[k]);
    std::vector<const uint8_t*> level_ptr;
    return run_mip_levels(descs, static_cast<const uint8_t*>(d_src), level0->src_row_pitch, filter, d_outs.data(), level_ptr, s);
}

} // namespace cfx

using namespace cfx;

extern "C" {

int cfx_init(int device)
{
    std::lock_guard<std::mutex> lock(g_ctx.mutex);
    t_error[0] = 0;
    return ensure_init(device);
}

void cfx_shutdown(void)
{
    std::lock_guard<std::mutex> lock(g_ctx.mutex);
    if (!g_ctx.ready) return;
    cudaSetDevice(g_ctx.device);
    cudaDeviceSynchronize();
    for (auto& s : g_ctx.streams) if (s) { cudaStreamDestroy(s); s = nullptr; }
    for (auto& e : g_ctx.fork) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (auto& e : g_ctx.join) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (auto& e : g_ctx.uploaded) if (e) { cudaEventDestroy(e); e = nullptr; }
    if (g_ctx.d_src) cudaFree(g_ctx.d_src);
    if (g_ctx.d_dst) cudaFree(g_ctx.d_dst);
    if (g_ctx.d_mip) cudaFree(g_ctx.d_mip);
    g_ctx.d_src = g_ctx.d_dst = g_ctx.d_mip = nullptr; g_ctx.d_src_cap = g_ctx.d_dst_cap = g_ctx.d_mip_cap = 0;
    g_ctx.ready = false;
}

int cfx_format_supported(uint32_t format, uint32_t type) { return find_launcher(format, type) != nullptr; }

int cfx_format_is_exact(uint32_t format, uint32_t type, uint32_t quality)
{
    if (!find_launcher(format, type) || quality > CFX_QUALITY_HIGHEST) return 0;
    switch (format) {
        case CFX_FORMAT_BC4: case CFX_FORMAT_BC5: return 1;
#ifdef CFX_HAVE_BC1
        // BC1_RGBA: exact for blocks without transparent texels (the others go through libsquish in the reference)
        case CFX_FORMAT_BC1_RGB: case CFX_FORMAT_BC2: case CFX_FORMAT_BC3: return bc1_color_is_exact(quality) ? 1 : 0;
#endif
#ifdef CFX_HAVE_ETC
        case CFX_FORMAT_ETC1: return etc1_is_exact(quality) ? 1 : 0;       // linear colour space
#endif
        default: return 0;
    }
}

int cfx_block_info(uint32_t format, uint32_t* bw, uint32_t* bh, uint32_t* bytes)
{
    uint32_t w, h, b;
    if (!block_info(format, w, h, b)) return CFX_ERR_UNSUPPORTED;
    if (bw) *bw = w;
    if (bh) *bh = h;
    if (bytes) *bytes = b;
    return CFX_OK;
}

size_t cfx_encoded_size(const cfx_surface_desc* d)
{
    uint32_t bw, bh, bytes;
    if (!d || !block_info(d->format, bw, bh, bytes)) return 0;
    return static_cast<size_t>((d->width + bw - 1)/bw)*((d->height + bh - 1)/bh)*bytes;
}

int cfx_encode(const cfx_surface_desc* desc, const void* src, void* dst, size_t dst_size)
{
    std::lock_guard<std::mutex> lock(g_ctx.mutex);
    t_error[0] = 0;
    return encode_host(desc, src, dst, dst_size);
}

int cfx_encode_batch(int n, const cfx_surface_desc* descs, const void* const* srcs, void* const* dsts,
    const size_t* dst_sizes)
{
    std::lock_guard<std::mutex> lock(g_ctx.mutex);
    t_error[0] = 0;
    if (n < 0 || (n > 0 && (!descs || !srcs || !dsts || !dst_sizes))) return fail(CFX_ERR_INVALID, "bad batch arguments");
    // Validate everything first so a bad surface fails the batch before any work is queued
    // (Converter::convert only tolerates a missing converter on the first surface).
    for (int i = 0; i < n; ++i) {
        EncodeParams p; Launcher l;
        int rc = validate(&descs[i], p, l);
        if (rc != CFX_OK) return rc;
    }
    // The surfaces of a batch (a mip chain, array layers) live back to back in the device buffers, their copies and
    // kernels are queued round-robin on the context's streams and awaited once: the small levels of a chain, which are
    // launch- and latency-bound, overlap each other and the tail of the big ones.
    std::vector<size_t> src_off(n + 1, 0), dst_off(n + 1, 0);
    for (int i = 0; i < n; ++i) {
        const size_t pitch = align256(static_cast<size_t>(descs[i].width)*src_texel_bytes(descs[i].src_format));
        src_off[i + 1] = src_off[i] + align256(pitch*descs[i].height);
        dst_off[i + 1] = dst_off[i] + align256(cfx_encoded_size(&descs[i]));
    }
    int rc = ensure_init(-1);
    if (rc != CFX_OK) return rc;
    rc = reserve(g_ctx.d_src, g_ctx.d_src_cap, src_off[n]);
    if (rc != CFX_OK) return rc;
    rc = reserve(g_ctx.d_dst, g_ctx.d_dst_cap, dst_off[n]);
    if (rc != CFX_OK) return rc;
    for (int i = 0; i < n; ++i) {
        rc = encode_host(&descs[i], srcs[i], dsts[i], dst_sizes[i], src_off[i], dst_off[i], false, i);
        if (rc != CFX_OK) break;
    }
    for (auto& s : g_ctx.streams) {
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess && rc == CFX_OK) rc = fail(CFX_ERR_CUDA, "%s", cudaGetErrorString(e));
    }
    return rc;
}

int cfx_encode_device(const cfx_surface_desc* desc, const void* d_src, void* d_dst, size_t dst_size,
    void* cuda_stream)
{
    std::lock_guard<std::mutex> lock(g_ctx.mutex);
    t_error[0] = 0;
    EncodeParams p; Launcher launcher;
    int rc = validate(desc, p, launcher);
    if (rc != CFX_OK) return rc;
    if (!d_src || !d_dst) return fail(CFX_ERR_INVALID, "null buffer");
    size_t out_bytes = static_cast<size_t>(p.total_blocks)*p.block_bytes;
    if (dst_size < out_bytes) return fail(CFX_ERR_INVALID, "dst_size %zu < %zu", dst_size, out_bytes);
    rc = ensure_init(-1);
    if (rc != CFX_OK) return rc;
    p.src = static_cast<const uint8_t*>(d_src);
    p.dst = static_cast<uint8_t*>(d_dst);
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);   // NULL = the CUDA default stream
    return launch(launcher, p, s);
}

int cfx_resize(const void* src, uint32_t src_width, uint32_t src_height, size_t src_row_pitch, void* dst, uint32_t dst_width,
    uint32_t dst_height, size_t dst_row_pitch, uint32_t filter, uint32_t color_space)
{
    std::lock_guard<std::mutex> lock(g_ctx.mutex);
    t_error[0] = 0;
    return resize_host(src, src_width, src_height, src_row_pitch, dst, dst_width, dst_height, dst_row_pitch, filter, color_space);
}

uint32_t cfx_mip_levels(uint32_t width, uint32_t height) { return mip_levels(width, height); }

int cfx_encode_mip_chain(const cfx_surface_desc* level0, const void* src, uint32_t filter, uint32_t levels, void* const* dsts,
    const size_t* dst_sizes, void* const* mip_images)
{
    std::lock_guard<std::mutex> lock(g_ctx.mutex);
    t_error[0] = 0;
    return encode_mip_chain(level0, src, filter, levels, dsts, dst_sizes, mip_images);
}

int cfx_encode_mip_chain_device(const cfx_surface_desc* level0, const void* d_src, uint32_t filter, uint32_t levels,
    void* const* d_dsts, const size_t* dst_sizes, void* cuda_stream)
{
    std::lock_guard<std::mutex> lock(g_ctx.mutex);
    t_error[0] = 0;
    return encode_mip_chain_device(level0, d_src, filter, levels, d_dsts, dst_sizes, static_cast<cudaStream_t>(cuda_stream));
}

void* cfx_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { fail(CFX_ERR_CUDA, "cudaMallocHost(%zu) failed", bytes); return nullptr; }
    return p;
}

void cfx_host_free(void* p) { if (p) cudaFreeHost(p); }

uint64_t cfx_kernel_launches(void) { return g_launches.load(); }
const char* cfx_last_error(void) { return t_error; }
const char* cfx_version(void) { return "cuttlefish-b200 0.1 (sm_100a)"; }

} // extern "C"
