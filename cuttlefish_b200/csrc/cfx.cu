// C-ABI of the encoder library (include/cfx.h): descriptor validation, per-device contexts (streams, device buffers,
// pinned staging), the device pool a host-buffer call shards a surface across, the chunked H2D -> kernel -> D2H pipeline
// of cfx_encode() / cfx_encode_batch(), and dispatch to the per-format kernel launchers.
//
// This file is the device-side stand-in for Converter::convert()'s per-surface job loop
// (lib/src/Converter.cpp:508-593): where the reference enumerates jobsX x jobsY process(x,y)
// calls over a std::thread pool, this splits the surface's block rows over the pool's GPUs, and on each
// GPU enqueues one persistent kernel per chunk of block rows, with the chunk's upload and the download of its packed
// blocks (straight into the caller's buffer at the chunk's byte offset) overlapping its neighbours' kernels.
#include "../../include/cfx.h"
#include "host_stage.h"
#include "kernels.h"

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

namespace cfx {

static thread_local char t_error[512] = "";
static thread_local int t_sm_count = 0;          // SM count of the device the current launch goes to
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
            if (format >= CFX_FORMAT_ASTC_4x4 && format <= CFX_FORMAT_ASTC_12x12)
                return (type == CFX_TYPE_UNORM || type == CFX_TYPE_UFLOAT) ? launch_astc : nullptr;
#endif
            return nullptr;
    }
}

static uint32_t src_texel_bytes(uint32_t src_format)
{
    return src_format == CFX_SRC_RGBA8 ? 4u : src_format == CFX_SRC_RGBA16F ? 8u : 16u;
}

static size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

// ---- device contexts and the pool ------------------------------------------------------------

constexpr int kStreams = 3;
constexpr int kMaxMipLevels = 32;
constexpr int kSlots = 4;                         // pinned staging slots per context (pageable sources)
constexpr size_t kSlotBytes = 8u << 20;

struct Context {
    int device = -1;
    int sm_count = 0;
    cudaStream_t streams[kStreams] = {};
    uint8_t* d_src = nullptr; size_t d_src_cap = 0;
    uint8_t* d_dst = nullptr; size_t d_dst_cap = 0;
    uint8_t* d_mip = nullptr; size_t d_mip_cap = 0;       // resized surfaces + the resize scratch
    cudaEvent_t fork[kMaxMipLevels] = {};                 // "level k is filtered" / "stream i has encoded its levels"
    cudaEvent_t join[kStreams] = {};
    cudaEvent_t uploaded[kStreams] = {};                  // "this stream's share of a surface is in HBM"
    // pinned staging: kSlots upload slots (a pageable source passes through them) and one download area
    uint8_t* h_in = nullptr; size_t h_slot_bytes = 0;
    cudaEvent_t slot_free[kSlots] = {};
    bool slot_busy[kSlots] = {};
    int next_slot = 0;
    uint8_t* h_out = nullptr; size_t h_out_cap = 0;
    std::vector<cudaEvent_t> chunk_done;                  // grows on demand
    int next_stream = 0;
};

static std::mutex g_mutex;                                  // one library call at a time
static std::vector<std::unique_ptr<Context>> g_contexts;    // every context alive
static std::vector<Context*> g_pool;                        // the devices host-buffer calls shard across

uint32_t persistent_ctas(const void* kernel, int threads, size_t dyn_smem)
{
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem) != cudaSuccess ||
        per_sm < 1)
        per_sm = 1;
    int sms = t_sm_count > 0 ? t_sm_count : 148;
    return static_cast<uint32_t>(sms*per_sm);
}

// Every entry point leaves the caller's current CUDA device as it found it.
struct DeviceGuard {
    int saved = -1;
    DeviceGuard() { if (cudaGetDevice(&saved) != cudaSuccess) { saved = -1; cudaGetLastError(); } }
    ~DeviceGuard() { if (saved >= 0) cudaSetDevice(saved); }
};

static void destroy_context(Context& c)
{
    if (c.device < 0) return;
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    for (auto& s : c.streams) if (s) { cudaStreamDestroy(s); s = nullptr; }
    for (auto& e : c.fork) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (auto& e : c.join) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (auto& e : c.uploaded) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (auto& e : c.slot_free) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (auto& e : c.chunk_done) cudaEventDestroy(e);
    c.chunk_done.clear();
    if (c.d_src) cudaFree(c.d_src);
    if (c.d_dst) cudaFree(c.d_dst);
    if (c.d_mip) cudaFree(c.d_mip);
    if (c.h_in) cudaFreeHost(c.h_in);
    if (c.h_out) cudaFreeHost(c.h_out);
    c.d_src = c.d_dst = c.d_mip = c.h_in = c.h_out = nullptr;
    c.d_src_cap = c.d_dst_cap = c.d_mip_cap = c.h_slot_bytes = c.h_out_cap = 0;
    c.device = -1;
}

static int device_count(int& count)
{
    count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(CFX_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
            e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return CFX_OK;
}

static int create_context(int device, Context*& out)
{
    cudaDeviceProp prop;
    CFX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(CFX_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
            device, prop.major, prop.minor);
    CFX_CUDA(cudaSetDevice(device));
    std::unique_ptr<Context> c(new Context());
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    int rc = CFX_OK;
    auto ok = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == CFX_OK) rc = fail(CFX_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
    };
    for (auto& s : c->streams) ok(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate");
    for (auto& e : c->fork) ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    for (auto& e : c->join) ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    for (auto& e : c->uploaded) ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    for (auto& e : c->slot_free) ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    if (rc != CFX_OK) { destroy_context(*c); return rc; }
    out = c.get();
    g_contexts.push_back(std::move(c));
    return CFX_OK;
}

// The first context living on `device`, created on demand (device-pointer entry points, pool members).
static int context_for_device(int device, Context*& out)
{
    for (auto& c : g_contexts)
        if (c->device == device) { out = c.get(); return CFX_OK; }
    int count;
    int rc = device_count(count);
    if (rc != CFX_OK) return rc;
    if (device < 0 || device >= count) return fail(CFX_ERR_INVALID, "device %d out of range (%d devices)", device, count);
    return create_context(device, out);
}

// Pool := the given devices, in order. A device may be listed more than once: each mention gets its own context
// (streams, buffers), which is how the sharding logic is tested on a one-GPU box.
static int set_pool(int n, const int* devices)
{
    int count;
    int rc = device_count(count);
    if (rc != CFX_OK) return rc;
    if (n < 1 || n > 64) return fail(CFX_ERR_INVALID, "device pool of %d entries", n);
    std::vector<Context*> pool;
    std::vector<bool> taken(g_contexts.size(), false);
    for (int i = 0; i < n; ++i) {
        if (devices[i] < 0 || devices[i] >= count)
            return fail(CFX_ERR_INVALID, "device %d out of range (%d devices)", devices[i], count);
        Context* c = nullptr;
        for (size_t k = 0; k < taken.size() && !c; ++k)
            if (!taken[k] && g_contexts[k]->device == devices[i]) { taken[k] = true; c = g_contexts[k].get(); }
        if (!c) {
            rc = create_context(devices[i], c);
            if (rc != CFX_OK) return rc;
            taken.push_back(true);
        }
        pool.push_back(c);
    }
    g_pool.swap(pool);
    return CFX_OK;
}

static int ensure_pool()
{
    if (!g_pool.empty()) return CFX_OK;
    int count;
    int rc = device_count(count);
    if (rc != CFX_OK) return rc;
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) { cudaGetLastError(); device = 0; }
    return set_pool(1, &device);
}

static int sync_streams(Context& c, int rc)
{
    cudaSetDevice(c.device);
    for (auto& s : c.streams) {
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess && rc == CFX_OK) rc = fail(CFX_ERR_CUDA, "device %d: %s", c.device, cudaGetErrorString(e));
    }
    return rc;
}

// Grows a device buffer of `c` (current device). Growth only happens while none of the context's own work is in
// flight (host-buffer calls return synchronised), so waiting for its streams costs nothing in the steady state.
static int reserve(Context& c, uint8_t*& ptr, size_t& cap, size_t bytes)
{
    if (bytes <= cap) return CFX_OK;
    if (ptr) {
        int rc = sync_streams(c, CFX_OK);
        if (rc != CFX_OK) return rc;
        CFX_CUDA(cudaFree(ptr));
        ptr = nullptr; cap = 0;
    }
    size_t want = bytes + bytes/4 + 4096;
    CFX_CUDA(cudaMalloc(&ptr, want));
    cap = want;
    return CFX_OK;
}

static int reserve_pinned(uint8_t*& ptr, size_t& cap, size_t bytes)
{
    if (bytes <= cap) return CFX_OK;
    if (ptr) { CFX_CUDA(cudaFreeHost(ptr)); ptr = nullptr; cap = 0; }
    size_t want = bytes + bytes/4 + 4096;
    CFX_CUDA(cudaHostAlloc(&ptr, want, cudaHostAllocPortable));
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
    if (d->flags & ~static_cast<uint32_t>(CFX_FLAG_BOTTOM_UP)) return fail(CFX_ERR_INVALID, "unknown flags 0x%x", d->flags);
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

// Launches on the current device (= c.device).
static int launch(Context& c, Launcher launcher, EncodeParams& p, cudaStream_t stream)
{
    p.aligned16 = ((reinterpret_cast<uintptr_t>(p.src) | p.pitch) & 15) == 0;
    t_sm_count = c.sm_count;
    int n = launcher(p, stream);
    if (n < 0) {
        cudaError_t e = cudaGetLastError();
        return fail(n, "format %u: the kernel launcher failed (%s)", p.format,
            e != cudaSuccess ? cudaGetErrorString(e) : n == CFX_ERR_UNSUPPORTED ? "unsupported configuration" : "no CUDA error recorded");
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CFX_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    g_launches += static_cast<uint64_t>(n);
    return CFX_OK;
}

// ---- the host-buffer pipeline ------------------------------------------------------------------

static bool is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// What a PAGEABLE source of this surface is narrowed to on its way through the staging slots (host_stage.h): only
// where the kernel's load stage would compute exactly that view anyway, so the blocks do not change.
static StageOp stage_op_for(const EncodeParams& p)
{
    if (p.src_format != CFX_SRC_RGBA32F) return STAGE_COPY;
    switch (p.format) {
        case CFX_FORMAT_BC1_RGB: case CFX_FORMAT_BC1_RGBA: case CFX_FORMAT_BC2: case CFX_FORMAT_BC3: case CFX_FORMAT_BC7:
            return STAGE_F32_TO_U8;                               // stage_tile_u8 / load_texel_u8 (common.cuh)
        case CFX_FORMAT_BC4: case CFX_FORMAT_BC5:
            return p.type == CFX_TYPE_UNORM ? STAGE_F32_TO_U8 : STAGE_COPY;
        case CFX_FORMAT_BC6H:
            return STAGE_F32_TO_F16;                              // bc6h.cu converts with __float2half_rn
        default:
            return STAGE_COPY;                                    // ETC / ASTC search on the unquantised floats
    }
}

struct Surface {                 // one validated surface of a host-buffer call
    EncodeParams p;              // src/dst unset
    Launcher launcher;
    const uint8_t* src; uint8_t* dst;
    uint64_t host_pitch;
    bool bottom_up, src_pinned, dst_pinned;
    StageOp op;                  // pageable sources only
    uint32_t dev_src_format;     // what the kernel sees
    size_t d_pitch;
};

struct Chunk {                   // block rows [r0, r1) of surface s on context c
    int s; Context* c;
    uint32_t r0, r1;
    size_t d_src_off;            // device offset of the piece this chunk belongs to
    uint32_t piece_s0;           // first stored row of that piece
    size_t d_dst_off;            // device offset of this chunk's blocks
    size_t h_out_off;            // offset in the context's pinned download area (pageable dst)
    int done_event = -1;
};

struct Plan {                    // the chunks of one host-buffer call
    std::vector<std::vector<Chunk>> per_dev;
    std::vector<Chunk*> order;
    std::vector<bool> dev_used;
};

// Lays the surfaces out over the pool's devices and enqueues every chunk's upload, kernel and download.
// dst_extra: bytes to keep free in d_dst behind the chunks' blocks (the mip chain's tail levels); *dst_end receives
// where they start on the pool's first device. uploaded_mask (mip chain): receives a bit per stream of the first device
// that carried a piece of the upload; Context::uploaded[i] of those fires once that stream's last piece is in HBM.
// Returns CFX_OK or the first error; finish_surfaces() must run either way.
static int enqueue_surfaces(std::vector<Surface>& surfs, Plan& plan, size_t dst_extra = 0, size_t* dst_end = nullptr,
    uint32_t* uploaded_mask = nullptr)
{
    const int ndev = static_cast<int>(g_pool.size());
    std::vector<std::vector<Chunk>>& per_dev = plan.per_dev;
    per_dev.assign(ndev, std::vector<Chunk>());
    plan.dev_used.assign(ndev, false);
    std::vector<size_t> src_used(ndev, 0), dst_used(ndev, 0), hout_used(ndev, 0), load(ndev, 0);
    std::vector<bool> need_slots(ndev, false);

    auto add_piece = [&](int s, int dev, uint32_t r0, uint32_t r1) {
        Surface& sf = surfs[s];
        const EncodeParams& p = sf.p;
        const uint32_t y0 = r0*p.block_h, y1 = std::min(p.height, r1*p.block_h);
        const uint32_t s0 = sf.bottom_up ? p.height - y1 : y0;
        const size_t d_src_off = src_used[dev];
        src_used[dev] += align256(sf.d_pitch*(y1 - y0));
        const uint64_t texels = static_cast<uint64_t>(p.width)*(y1 - y0);
        load[dev] += texels;
        // >= 4 chunks per piece once it is big enough to keep every kernel a few full waves; ~8 M texels per chunk above
        uint64_t n = (texels + (8ull << 20) - 1)/(8ull << 20);
        const uint64_t lo = std::min<uint64_t>(4, texels >> 20);
        if (n < lo) n = lo;
        if (n < 1) n = 1;
        if (n > r1 - r0) n = r1 - r0;
        for (uint64_t k = 0; k < n; ++k) {
            Chunk c;
            c.s = s; c.c = g_pool[dev];
            c.r0 = r0 + static_cast<uint32_t>((r1 - r0)*k/n);
            c.r1 = r0 + static_cast<uint32_t>((r1 - r0)*(k + 1)/n);
            c.d_src_off = d_src_off; c.piece_s0 = s0;
            const size_t bytes = static_cast<size_t>(c.r1 - c.r0)*p.blocks_x*p.block_bytes;
            c.d_dst_off = dst_used[dev]; dst_used[dev] += align256(bytes);
            c.h_out_off = hout_used[dev];
            if (!sf.dst_pinned) hout_used[dev] += align256(bytes);
            per_dev[dev].push_back(c);
        }
        if (!sf.src_pinned) need_slots[dev] = true;
    };

    for (int s = 0; s < static_cast<int>(surfs.size()); ++s) {
        const EncodeParams& p = surfs[s].p;
        const uint64_t texels = static_cast<uint64_t>(p.width)*p.height;
        if (ndev > 1 && p.blocks_y >= static_cast<uint32_t>(ndev) && texels >= static_cast<uint64_t>(ndev) << 18) {
            // contiguous block-row ranges, rank k gets [k*R/P, (k+1)*R/P)
            for (int k = 0; k < ndev; ++k) {
                const uint32_t r0 = static_cast<uint32_t>(static_cast<uint64_t>(p.blocks_y)*k/ndev);
                const uint32_t r1 = static_cast<uint32_t>(static_cast<uint64_t>(p.blocks_y)*(k + 1)/ndev);
                if (r1 > r0) add_piece(s, k, r0, r1);
            }
        } else {
            int dev = 0;                               // a small surface goes whole to the least loaded device
            for (int k = 1; k < ndev; ++k) if (load[k] < load[dev]) dev = k;
            add_piece(s, dev, 0, p.blocks_y);
        }
    }

    int rc = CFX_OK;
    for (int k = 0; k < ndev && rc == CFX_OK; ++k) {
        if (per_dev[k].empty()) continue;
        Context& c = *g_pool[k];
        if (cudaSetDevice(c.device) != cudaSuccess) { rc = fail(CFX_ERR_CUDA, "cudaSetDevice(%d) failed", c.device); break; }
        rc = reserve(c, c.d_src, c.d_src_cap, src_used[k]);
        if (rc == CFX_OK) rc = reserve(c, c.d_dst, c.d_dst_cap, dst_used[k] + dst_extra);
        if (rc == CFX_OK && hout_used[k]) rc = reserve_pinned(c.h_out, c.h_out_cap, hout_used[k]);
        if (rc == CFX_OK && need_slots[k]) {
            size_t slot = kSlotBytes;
            for (const Chunk& ch : per_dev[k]) slot = std::max(slot, surfs[ch.s].d_pitch);
            if (slot > c.h_slot_bytes) {
                size_t cap = c.h_slot_bytes*kSlots;
                rc = reserve_pinned(c.h_in, cap, slot*kSlots);
                c.h_slot_bytes = rc == CFX_OK ? cap/kSlots : 0;
                for (auto& b : c.slot_busy) b = false;
            }
        }
        while (rc == CFX_OK && c.chunk_done.size() < per_dev[k].size()) {
            cudaEvent_t e;
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) rc = fail(CFX_ERR_CUDA, "cudaEventCreate failed");
            else c.chunk_done.push_back(e);
        }
        for (size_t i = 0; i < per_dev[k].size(); ++i) per_dev[k][i].done_event = static_cast<int>(i);
    }

    if (dst_end) *dst_end = dst_used[0];

    // round-robin over the devices so that every GPU gets its first chunk early
    std::vector<Chunk*>& order = plan.order;
    order.clear();
    for (size_t i = 0;; ++i) {
        bool any = false;
        for (int k = 0; k < ndev; ++k)
            if (i < per_dev[k].size()) { order.push_back(&per_dev[k][i]); any = true; }
        if (!any) break;
    }

    std::vector<bool>& dev_used = plan.dev_used;
    for (size_t oi = 0; oi < order.size() && rc == CFX_OK; ++oi) {
        Chunk& ch = *order[oi];
        Surface& sf = surfs[ch.s];
        Context& c = *ch.c;
        const EncodeParams& p = sf.p;
        if (cudaSetDevice(c.device) != cudaSuccess) { rc = fail(CFX_ERR_CUDA, "cudaSetDevice(%d) failed", c.device); break; }
        for (int k = 0; k < ndev; ++k) if (g_pool[k] == &c) dev_used[k] = true;
        const int si = c.next_stream++ % kStreams;
        cudaStream_t st = c.streams[si];
        const uint32_t y0 = ch.r0*p.block_h, y1 = std::min(p.height, ch.r1*p.block_h);
        // stored rows [s0, s1) of the host image hold top-down rows [y0, y1)
        const uint32_t s0 = sf.bottom_up ? p.height - y1 : y0, s1 = s0 + (y1 - y0);
        uint8_t* d_rows = c.d_src + ch.d_src_off + static_cast<size_t>(s0 - ch.piece_s0)*sf.d_pitch;
        const size_t dev_row_bytes = static_cast<size_t>(p.width)*src_texel_bytes(sf.dev_src_format);
        cudaError_t e = cudaSuccess;
        if (sf.src_pinned) {
            e = cudaMemcpy2DAsync(d_rows, sf.d_pitch, sf.src + static_cast<size_t>(s0)*sf.host_pitch, sf.host_pitch, dev_row_bytes,
                s1 - s0, cudaMemcpyHostToDevice, st);
        } else {
            const uint32_t rows_per_slot = static_cast<uint32_t>(std::max<size_t>(1, c.h_slot_bytes/sf.d_pitch));
            const size_t in_texel = src_texel_bytes(p.src_format);
            for (uint32_t a = s0; a < s1 && e == cudaSuccess; a += rows_per_slot) {
                const uint32_t b = std::min(s1, a + rows_per_slot);
                const int slot = c.next_slot;
                c.next_slot = (c.next_slot + 1) % kSlots;
                if (c.slot_busy[slot]) { e = cudaEventSynchronize(c.slot_free[slot]); if (e != cudaSuccess) break; }
                uint8_t* h = c.h_in + static_cast<size_t>(slot)*c.h_slot_bytes;
                const uint8_t* from = sf.src + static_cast<size_t>(a)*sf.host_pitch;
                const size_t d_pitch = sf.d_pitch, host_pitch = sf.host_pitch, width = p.width;
                const StageOp op = sf.op;
                auto row = [=](size_t r) { stage_row(op, h + r*d_pitch, from + r*host_pitch, width, in_texel); };
                if (static_cast<uint64_t>(b - a)*width < (64u << 10)) for (uint32_t r = 0; r < b - a; ++r) row(r);
                else parallel_for(b - a, row);
                e = cudaMemcpyAsync(d_rows + static_cast<size_t>(a - s0)*d_pitch, h, static_cast<size_t>(b - a)*d_pitch,
                    cudaMemcpyHostToDevice, st);
                if (e == cudaSuccess) e = cudaEventRecord(c.slot_free[slot], st);
                c.slot_busy[slot] = true;
            }
        }
        if (e == cudaSuccess && uploaded_mask) {
            e = cudaEventRecord(c.uploaded[si], st);
            *uploaded_mask |= 1u << si;
        }
        if (e != cudaSuccess) { rc = fail(CFX_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e)); break; }
        EncodeParams k = p;
        k.src_format = sf.dev_src_format;
        if (sf.bottom_up) {
            // top-down row y lives at stored row (height-1-y): start at the chunk's last stored row and walk backwards
            k.src = c.d_src + ch.d_src_off + static_cast<size_t>(p.height - 1 - y0 - ch.piece_s0)*sf.d_pitch;
            k.pitch = static_cast<uint64_t>(0) - static_cast<uint64_t>(sf.d_pitch);
        } else {
            k.src = d_rows;
            k.pitch = sf.d_pitch;
        }
        k.height = y1 - y0;   // interior chunks end on a block-row boundary, so the clamp is unchanged
        k.blocks_y = ch.r1 - ch.r0;
        k.total_blocks = k.blocks_x*k.blocks_y;
        k.dst = c.d_dst + ch.d_dst_off;
        rc = launch(c, sf.launcher, k, st);
        if (rc != CFX_OK) break;
        const size_t bytes = static_cast<size_t>(k.total_blocks)*p.block_bytes;
        const size_t host_off = static_cast<size_t>(ch.r0)*p.blocks_x*p.block_bytes;
        e = cudaMemcpyAsync(sf.dst_pinned ? sf.dst + host_off : c.h_out + ch.h_out_off, k.dst, bytes, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && !sf.dst_pinned) e = cudaEventRecord(c.chunk_done[ch.done_event], st);
        if (e != cudaSuccess) { rc = fail(CFX_ERR_CUDA, "download failed: %s", cudaGetErrorString(e)); break; }
    }
    return rc;
}

// Awaits what enqueue_surfaces() queued (always, also after an error: no copy may be left in flight on the caller's
// buffers). Pageable destinations: every chunk is copied out of the pinned download area as soon as it has landed.
static int finish_surfaces(std::vector<Surface>& surfs, Plan& plan, int rc)
{
    const int ndev = static_cast<int>(g_pool.size());
    std::vector<Chunk*>& order = plan.order;
    for (size_t oi = 0; oi < order.size() && rc == CFX_OK; ++oi) {
        Chunk& ch = *order[oi];
        Surface& sf = surfs[ch.s];
        if (sf.dst_pinned) continue;
        cudaSetDevice(ch.c->device);
        cudaError_t e = cudaEventSynchronize(ch.c->chunk_done[ch.done_event]);
        if (e != cudaSuccess) { rc = fail(CFX_ERR_CUDA, "device %d: %s", ch.c->device, cudaGetErrorString(e)); break; }
        const size_t bytes = static_cast<size_t>(ch.r1 - ch.r0)*sf.p.blocks_x*sf.p.block_bytes;
        memcpy(sf.dst + static_cast<size_t>(ch.r0)*sf.p.blocks_x*sf.p.block_bytes, ch.c->h_out + ch.h_out_off, bytes);
    }
    for (int k = 0; k < ndev; ++k)
        if (plan.dev_used[k] || rc != CFX_OK) rc = sync_streams(*g_pool[k], rc);
    return rc;
}

static int make_surface(const cfx_surface_desc* desc, const void* src, void* dst, size_t dst_size, Surface& sf)
{
    int rc = validate(desc, sf.p, sf.launcher);
    if (rc != CFX_OK) return rc;
    if (!src || !dst) return fail(CFX_ERR_INVALID, "null buffer");
    const size_t out_bytes = static_cast<size_t>(sf.p.total_blocks)*sf.p.block_bytes;
    if (dst_size < out_bytes) return fail(CFX_ERR_INVALID, "dst_size %zu < %zu", dst_size, out_bytes);
    sf.src = static_cast<const uint8_t*>(src);
    sf.dst = static_cast<uint8_t*>(dst);
    sf.host_pitch = desc->src_row_pitch;
    sf.bottom_up = (desc->flags & CFX_FLAG_BOTTOM_UP) != 0;
    return CFX_OK;
}

static void classify_buffers(Surface& sf)
{
    sf.src_pinned = is_pinned(sf.src);
    sf.dst_pinned = is_pinned(sf.dst);
    sf.op = sf.src_pinned ? STAGE_COPY : stage_op_for(sf.p);
    sf.dev_src_format = sf.op == STAGE_F32_TO_U8 ? CFX_SRC_RGBA8 : sf.op == STAGE_F32_TO_F16 ? CFX_SRC_RGBA16F : sf.p.src_format;
    sf.d_pitch = align256(static_cast<size_t>(sf.p.width)*src_texel_bytes(sf.dev_src_format));
}

// resize.cu
size_t resize_scratch_bytes(uint32_t sw, uint32_t sh, uint32_t dw, uint32_t dh);
int resize_device(const uint8_t* src, size_t src_pitch, bool src_u8, uint32_t sw, uint32_t sh, uint8_t* dst, size_t dst_pitch,
    uint32_t dw, uint32_t dh, uint32_t filter, bool srgb, uint8_t* scratch, size_t scratch_cap, cudaStream_t stream);

constexpr uint32_t kMaxResizeRows = 65535;        // the filter passes launch with grid.y = rows

static int resize_host(const void* src, uint32_t sw, uint32_t sh, size_t src_pitch, void* dst, uint32_t dw, uint32_t dh,
    size_t dst_pitch, uint32_t filter, uint32_t color_space)
{
    if (!src || !dst) return fail(CFX_ERR_INVALID, "null buffer");
    if (!sw || !sh || !dw || !dh) return fail(CFX_ERR_INVALID, "empty surface");
    if (sh > kMaxResizeRows || dh > kMaxResizeRows)
        return fail(CFX_ERR_INVALID, "resize handles at most %u rows (got %u -> %u)", kMaxResizeRows, sh, dh);
    if (filter > CFX_FILTER_BSPLINE) return fail(CFX_ERR_INVALID, "filter %u out of range", filter);
    if (src_pitch < static_cast<size_t>(sw)*16u || dst_pitch < static_cast<size_t>(dw)*16u || (src_pitch & 3) || (dst_pitch & 3))
        return fail(CFX_ERR_INVALID, "row pitch too small or not a multiple of 4");
    int rc = ensure_pool();
    if (rc != CFX_OK) return rc;
    Context& c = *g_pool[0];
    CFX_CUDA(cudaSetDevice(c.device));
    const size_t sp = align256(static_cast<size_t>(sw)*16u), dp = align256(static_cast<size_t>(dw)*16u);
    const size_t scratch = resize_scratch_bytes(sw, sh, dw, dh);
    rc = reserve(c, c.d_src, c.d_src_cap, sp*sh);
    if (rc != CFX_OK) return rc;
    rc = reserve(c, c.d_mip, c.d_mip_cap, dp*dh + 256 + scratch);
    if (rc != CFX_OK) return rc;
    cudaStream_t s = c.streams[0];
    CFX_CUDA(cudaMemcpy2DAsync(c.d_src, sp, src, src_pitch, static_cast<size_t>(sw)*16u, sh, cudaMemcpyHostToDevice, s));
    uint8_t* d_out = c.d_mip;
    int n = resize_device(c.d_src, sp, false, sw, sh, d_out, dp, dw, dh, filter, color_space != 0, d_out + align256(dp*dh),
        scratch, s);
    if (n < 0) {
        cudaError_t e = cudaGetLastError();
        rc = fail(n, "resize failed: %s", e != cudaSuccess ? cudaGetErrorString(e) : "scratch too small");
        sync_streams(c, rc);
        return rc;
    }
    g_launches += static_cast<uint64_t>(n);
    cudaError_t e = cudaMemcpy2DAsync(dst, dst_pitch, d_out, dp, static_cast<size_t>(dw)*16u, dh, cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) rc = fail(CFX_ERR_CUDA, "download failed: %s", cudaGetErrorString(e));
    return sync_streams(c, rc);
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
    if (level0->flags) return fail(CFX_ERR_INVALID, "the mip chain takes a top-down level 0 (flags must be 0)");
    if (level0->src_format != CFX_SRC_RGBA32F && level0->src_format != CFX_SRC_RGBA8)
        return fail(CFX_ERR_INVALID, "the mip chain is generated from an RGBA32F level 0 (Image::Format::RGBAF) or an RGBA8 one "
            "(taken as v/255, Image::convert(RGBAF) of an 8-bit image)");
    if (level0->height > kMaxResizeRows)
        return fail(CFX_ERR_INVALID, "the mip chain handles at most %u rows (got %u)", kMaxResizeRows, level0->height);
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

// Levels of a chain whose level 0 is resident at d_level0 on c's device (current). The filter chain runs on stream s:
// each level is resized from the level above into the library's mip storage (every level keeps its own region until the
// next chain call). The encoders only depend on their own level, so they fork onto the context's other streams as soon
// as that level is filtered -- the launch-latency-bound tail levels then overlap each other and the big levels -- and
// join s at the end. Level 0 is encoded too when d_outs[0] is set. level_ptr[k] receives where level k's image lives.
static int run_mip_levels(Context& c, const std::vector<cfx_surface_desc>& descs, const uint8_t* d_level0, size_t pitch0,
    uint32_t filter, uint8_t* const* d_outs, std::vector<const uint8_t*>& level_ptr, cudaStream_t s)
{
    const uint32_t levels = static_cast<uint32_t>(descs.size());
    level_ptr.assign(levels, nullptr);
    level_ptr[0] = d_level0;
    std::vector<size_t> off(levels + 1, 0);
    for (uint32_t k = 1; k < levels; ++k) off[k + 1] = off[k] + align256(descs[k].src_row_pitch*descs[k].height);
    const size_t scratch = levels > 1 ? resize_scratch_bytes(descs[0].width, descs[0].height, descs[1].width, descs[1].height) : 0;
    int rc = reserve(c, c.d_mip, c.d_mip_cap, off[levels] + scratch + 256);
    if (rc != CFX_OK) return rc;
    uint8_t* d_scratch = c.d_mip + off[levels];
    size_t prev_pitch = pitch0;
    bool used[kStreams] = {};
    for (uint32_t k = 0; k < levels; ++k) {
        const cfx_surface_desc& d = descs[k];
        if (k) {
            uint8_t* cur = c.d_mip + off[k];
            int n = resize_device(level_ptr[k - 1], prev_pitch, descs[k - 1].src_format == CFX_SRC_RGBA8, descs[k - 1].width,
                descs[k - 1].height, cur, d.src_row_pitch,
                d.width, d.height, filter, descs[0].color_space != 0, d_scratch, scratch, s);
            if (n < 0) {
                cudaError_t e = cudaGetLastError();
                return fail(n, "level %u: resize failed (%s)", k, e != cudaSuccess ? cudaGetErrorString(e) : "scratch too small");
            }
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
        if (c.streams[a] == s) a = (a + 1) % kStreams;
        CFX_CUDA(cudaEventRecord(c.fork[k % kMaxMipLevels], s));
        CFX_CUDA(cudaStreamWaitEvent(c.streams[a], c.fork[k % kMaxMipLevels], 0));
        rc = launch(c, l, p, c.streams[a]);
        if (rc != CFX_OK) return rc;
        used[a] = true;
    }
    for (int a = 0; a < kStreams; ++a) {
        if (!used[a]) continue;
        CFX_CUDA(cudaEventRecord(c.join[a], c.streams[a]));
        CFX_CUDA(cudaStreamWaitEvent(s, c.join[a], 0));
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
    std::vector<Surface> surfs(1);
    rc = make_surface(level0, src, dsts[0], dst_sizes[0], surfs[0]);
    if (rc != CFX_OK) return rc;
    rc = ensure_pool();
    if (rc != CFX_OK) return rc;
    // The chain lives on ONE device (every level depends on the one above): the pool's first. Level 0 goes through the
    // chunked upload + encode of cfx_encode(), which leaves the whole surface in d_src, but is not awaited: the filter
    // chain only needs the upload, so it starts on the least busy stream as soon as every piece of level 0 is in HBM
    // and runs beside level 0's encoders. One wait at the end.
    std::vector<Context*> one(1, g_pool[0]);
    one.swap(g_pool);
    Context& c = *g_pool[0];
    std::vector<size_t> bytes(levels, 0), off(levels + 1, 0);
    for (uint32_t k = 1; k < levels; ++k) { bytes[k] = cfx_encoded_size(&descs[k]); off[k + 1] = off[k] + align256(bytes[k]); }
    classify_buffers(surfs[0]);
    if (!surfs[0].src_pinned) {
        // the filter reads level 0 as it was given (RGBA8 or RGBA32F): no narrowing of a pageable float source here
        surfs[0].op = STAGE_COPY; surfs[0].dev_src_format = level0->src_format;
        surfs[0].d_pitch = align256(static_cast<size_t>(level0->width)*src_texel_bytes(level0->src_format));
    }
    uint32_t carriers = 0;
    size_t tail_base = 0;
    Plan plan;
    rc = enqueue_surfaces(surfs, plan, off[levels], &tail_base, &carriers);
    if (rc == CFX_OK && levels > 1) {
        cudaSetDevice(c.device);
        cudaStream_t s = c.streams[c.next_stream % kStreams];          // the stream the next piece would have taken
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < kStreams && e == cudaSuccess; ++i)
            if ((carriers >> i) & 1u) e = cudaStreamWaitEvent(s, c.uploaded[i], 0);
        if (e != cudaSuccess) rc = fail(CFX_ERR_CUDA, "cudaStreamWaitEvent failed: %s", cudaGetErrorString(e));
        std::vector<uint8_t*> d_outs(levels, nullptr);
        for (uint32_t k = 1; k < levels; ++k) d_outs[k] = c.d_dst + tail_base + off[k];
        std::vector<const uint8_t*> level_ptr;
        if (rc == CFX_OK) rc = run_mip_levels(c, descs, c.d_src, surfs[0].d_pitch, filter, d_outs.data(), level_ptr, s);
        for (uint32_t k = 1; rc == CFX_OK && k < levels; ++k) {
            e = cudaMemcpyAsync(dsts[k], d_outs[k], bytes[k], cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess && mip_images && mip_images[k])
                e = cudaMemcpy2DAsync(mip_images[k], static_cast<size_t>(descs[k].width)*16u, level_ptr[k], descs[k].src_row_pitch,
                    static_cast<size_t>(descs[k].width)*16u, descs[k].height, cudaMemcpyDeviceToHost, s);
            if (e != cudaSuccess) rc = fail(CFX_ERR_CUDA, "level %u: download failed: %s", k, cudaGetErrorString(e));
        }
    }
    rc = finish_surfaces(surfs, plan, rc);
    rc = sync_streams(c, rc);
    one.swap(g_pool);
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
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, d_src) != cudaSuccess || (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged)) {
        cudaGetLastError();
        return fail(CFX_ERR_INVALID, "d_src is not a device pointer");
    }
    Context* c = nullptr;
    rc = context_for_device(attr.device, c);
    if (rc != CFX_OK) return rc;
    CFX_CUDA(cudaSetDevice(c->device));
    std::vector<uint8_t*> d_outs(levels, nullptr);
    for (uint32_t k = 0; k < levels; ++k) d_outs[k] = static_cast<uint8_t*>(d_dsts[k]);
    std::vector<const uint8_t*> level_ptr;
    return run_mip_levels(*c, descs, static_cast<const uint8_t*>(d_src), level0->src_row_pitch, filter, d_outs.data(), level_ptr, s);
}

void astc_release_tables();      // astc.cu: frees the per-device table blobs

} // namespace cfx

using namespace cfx;

extern "C" {

int cfx_init(int device)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceGuard guard;
    t_error[0] = 0;
    if (device < 0) return ensure_pool();
    return set_pool(1, &device);
}

int cfx_init_devices(int device_count_wanted)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceGuard guard;
    t_error[0] = 0;
    int count;
    int rc = device_count(count);
    if (rc != CFX_OK) return rc;
    if (device_count_wanted < 0) return fail(CFX_ERR_INVALID, "device count %d", device_count_wanted);
    // 0 = every visible sm_100 device; n = the first n of them
    std::vector<int> ids;
    for (int d = 0; d < count; ++d) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10) ids.push_back(d);
        if (device_count_wanted && static_cast<int>(ids.size()) == device_count_wanted) break;
    }
    if (ids.empty()) return fail(CFX_ERR_NO_DEVICE, "no sm_100 device among the %d visible", count);
    if (device_count_wanted && static_cast<int>(ids.size()) < device_count_wanted)
        return fail(CFX_ERR_INVALID, "%d devices wanted, %zu sm_100 devices visible", device_count_wanted, ids.size());
    return set_pool(static_cast<int>(ids.size()), ids.data());
}

int cfx_set_devices(int n, const int* devices)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceGuard guard;
    t_error[0] = 0;
    if (!devices) return fail(CFX_ERR_INVALID, "null device list");
    return set_pool(n, devices);
}

int cfx_device_count(void)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    return static_cast<int>(g_pool.size());
}

void cfx_shutdown(void)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceGuard guard;
    for (auto& c : g_contexts) destroy_context(*c);
    g_contexts.clear();
    g_pool.clear();
#ifdef CFX_HAVE_ASTC
    astc_release_tables();
#endif
}

int cfx_format_supported(uint32_t format, uint32_t type) { return find_launcher(format, type) != nullptr; }

int cfx_format_is_exact(uint32_t format, uint32_t type, uint32_t quality)
{
    if (!find_launcher(format, type) || quality > CFX_QUALITY_HIGHEST) return 0;
    switch (format) {
        case CFX_FORMAT_BC4: case CFX_FORMAT_BC5: return type == CFX_TYPE_UNORM ? 1 : 0;
#ifdef CFX_HAVE_BC1
        // BC1_RGBA: exact for blocks without transparent texels (the others go through libsquish in the reference)
        case CFX_FORMAT_BC1_RGB: case CFX_FORMAT_BC2: case CFX_FORMAT_BC3: return bc1_color_is_exact(quality) ? 1 : 0;
#endif
#ifdef CFX_HAVE_ETC
        case CFX_FORMAT_ETC1: return etc1_is_exact(quality) ? 1 : 0;       // linear and sRGB (RGBX / REC709 metric)
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

int cfx_encode_batch(int n, const cfx_surface_desc* descs, const void* const* srcs, void* const* dsts,
    const size_t* dst_sizes)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceGuard guard;
    t_error[0] = 0;
    if (n < 0 || (n > 0 && (!descs || !srcs || !dsts || !dst_sizes))) return fail(CFX_ERR_INVALID, "bad batch arguments");
    // Validate everything first so a bad surface fails the batch before any work is queued
    // (Converter::convert only tolerates a missing converter on the first surface).
    std::vector<Surface> surfs(static_cast<size_t>(n));
    for (int i = 0; i < n; ++i) {
        int rc = make_surface(&descs[i], srcs[i], dsts[i], dst_sizes[i], surfs[i]);
        if (rc != CFX_OK) return rc;
    }
    if (n == 0) return CFX_OK;
    int rc = ensure_pool();
    if (rc != CFX_OK) return rc;
    for (auto& sf : surfs) classify_buffers(sf);
    // The surfaces of a batch (a mip chain, array layers) live back to back in the device buffers, their copies and
    // kernels are queued round-robin on each device's streams and awaited once: the small levels of a chain, which are
    // launch- and latency-bound, overlap each other and the tail of the big ones.
    Plan plan;
    rc = enqueue_surfaces(surfs, plan);
    return finish_surfaces(surfs, plan, rc);
}

int cfx_encode(const cfx_surface_desc* desc, const void* src, void* dst, size_t dst_size)
{
    return cfx_encode_batch(1, desc, &src, &dst, &dst_size);
}

int cfx_encode_device(const cfx_surface_desc* desc, const void* d_src, void* d_dst, size_t dst_size,
    void* cuda_stream)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceGuard guard;
    t_error[0] = 0;
    EncodeParams p; Launcher launcher;
    int rc = validate(desc, p, launcher);
    if (rc != CFX_OK) return rc;
    if (!d_src || !d_dst) return fail(CFX_ERR_INVALID, "null buffer");
    size_t out_bytes = static_cast<size_t>(p.total_blocks)*p.block_bytes;
    if (dst_size < out_bytes) return fail(CFX_ERR_INVALID, "dst_size %zu < %zu", dst_size, out_bytes);
    const uint32_t texel = src_texel_bytes(desc->src_format);
    if ((reinterpret_cast<uintptr_t>(d_src) | desc->src_row_pitch) & (texel - 1))
        return fail(CFX_ERR_INVALID, "a device-resident surface must be texel aligned (%u bytes: pointer and pitch)", texel);
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, d_src) != cudaSuccess || (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged)) {
        cudaGetLastError();
        return fail(CFX_ERR_INVALID, "d_src is not a device pointer");
    }
    Context* c = nullptr;
    rc = context_for_device(attr.device, c);
    if (rc != CFX_OK) return rc;
    CFX_CUDA(cudaSetDevice(c->device));
    p.src = static_cast<const uint8_t*>(d_src);
    p.dst = static_cast<uint8_t*>(d_dst);
    if (desc->flags & CFX_FLAG_BOTTOM_UP) {
        p.src += static_cast<size_t>(p.height - 1)*desc->src_row_pitch;
        p.pitch = static_cast<uint64_t>(0) - desc->src_row_pitch;
    }
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);   // NULL = the CUDA default stream
    return launch(*c, launcher, p, s);
}

int cfx_resize(const void* src, uint32_t src_width, uint32_t src_height, size_t src_row_pitch, void* dst, uint32_t dst_width,
    uint32_t dst_height, size_t dst_row_pitch, uint32_t filter, uint32_t color_space)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceGuard guard;
    t_error[0] = 0;
    return resize_host(src, src_width, src_height, src_row_pitch, dst, dst_width, dst_height, dst_row_pitch, filter, color_space);
}

uint32_t cfx_mip_levels(uint32_t width, uint32_t height) { return mip_levels(width, height); }

int cfx_encode_mip_chain(const cfx_surface_desc* level0, const void* src, uint32_t filter, uint32_t levels, void* const* dsts,
    const size_t* dst_sizes, void* const* mip_images)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceGuard guard;
    t_error[0] = 0;
    return encode_mip_chain(level0, src, filter, levels, dsts, dst_sizes, mip_images);
}

int cfx_encode_mip_chain_device(const cfx_surface_desc* level0, const void* d_src, uint32_t filter, uint32_t levels,
    void* const* d_dsts, const size_t* dst_sizes, void* cuda_stream)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceGuard guard;
    t_error[0] = 0;
    return encode_mip_chain_device(level0, d_src, filter, levels, d_dsts, dst_sizes, static_cast<cudaStream_t>(cuda_stream));
}

// ---- cross-process peer memory ------------------------------------------------------------------

struct IpcBlob {                       // what cfx_ipc_export() writes into the caller's CFX_IPC_HANDLE_BYTES
    cudaIpcMemHandle_t handle;         // names the whole allocation d_ptr lives in
    uint64_t offset;                   // d_ptr - allocation base
    uint64_t magic;
};
static_assert(sizeof(IpcBlob) <= CFX_IPC_HANDLE_BYTES, "CFX_IPC_HANDLE_BYTES too small");
constexpr uint64_t kIpcMagic = 0x4346584950433031ull;    // "CFXIPC01"

typedef int (*MemGetAddressRangeFn)(unsigned long long*, size_t*, unsigned long long);

int cfx_ipc_export(const void* d_ptr, void* handle)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceGuard guard;
    t_error[0] = 0;
    if (!d_ptr || !handle) return fail(CFX_ERR_INVALID, "null argument");
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, d_ptr) != cudaSuccess || attr.type != cudaMemoryTypeDevice) {
        cudaGetLastError();
        return fail(CFX_ERR_INVALID, "d_ptr is not a device pointer");
    }
    CFX_CUDA(cudaSetDevice(attr.device));
    // the handle names the whole allocation (a torch tensor is a slice of a caching-allocator segment): find its base
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult res;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &res) != cudaSuccess ||
        res != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        return fail(CFX_ERR_CUDA, "cuMemGetAddressRange is not available");
    }
    unsigned long long base = 0;
    size_t size = 0;
    if (reinterpret_cast<MemGetAddressRangeFn>(fn)(&base, &size, reinterpret_cast<unsigned long long>(d_ptr)) != 0)
        return fail(CFX_ERR_CUDA, "cuMemGetAddressRange failed");
    IpcBlob blob;
    memset(&blob, 0, sizeof(blob));
    CFX_CUDA(cudaIpcGetMemHandle(&blob.handle, reinterpret_cast<void*>(base)));
    blob.offset = reinterpret_cast<unsigned long long>(d_ptr) - base;
    blob.magic = kIpcMagic;
    memset(handle, 0, CFX_IPC_HANDLE_BYTES);
    memcpy(handle, &blob, sizeof(blob));
    return CFX_OK;
}

int cfx_ipc_open(const void* handle, int device, void** d_ptr)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceGuard guard;
    t_error[0] = 0;
    if (!handle || !d_ptr) return fail(CFX_ERR_INVALID, "null argument");
    IpcBlob blob;
    memcpy(&blob, handle, sizeof(blob));
    if (blob.magic != kIpcMagic) return fail(CFX_ERR_INVALID, "not a handle made by cfx_ipc_export");
    Context* c = nullptr;
    int rc = context_for_device(device, c);
    if (rc != CFX_OK) return rc;
    CFX_CUDA(cudaSetDevice(device));
    void* base = nullptr;
    // peer access between `device` and the exporting device is switched on by the driver as part of the mapping
    CFX_CUDA(cudaIpcOpenMemHandle(&base, blob.handle, cudaIpcMemLazyEnablePeerAccess));
    *d_ptr = static_cast<uint8_t*>(base) + blob.offset;
    return CFX_OK;
}

int cfx_ipc_close(void* d_ptr, const void* handle)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    t_error[0] = 0;
    if (!d_ptr || !handle) return fail(CFX_ERR_INVALID, "null argument");
    IpcBlob blob;
    memcpy(&blob, handle, sizeof(blob));
    if (blob.magic != kIpcMagic) return fail(CFX_ERR_INVALID, "not a handle made by cfx_ipc_export");
    CFX_CUDA(cudaIpcCloseMemHandle(static_cast<uint8_t*>(d_ptr) - blob.offset));
    return CFX_OK;
}

void* cfx_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        fail(CFX_ERR_CUDA, "cudaHostAlloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}

void cfx_host_free(void* p) { if (p) cudaFreeHost(p); }

uint64_t cfx_kernel_launches(void) { return g_launches.load(); }
const char* cfx_last_error(void) { return t_error; }
const char* cfx_version(void) { return "cuttlefish-b200 0.2 (sm_100a)"; }

} // extern "C"
