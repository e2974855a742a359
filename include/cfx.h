/*
 * cfx.h -- C-ABI of the B200 block-texture encoder that drops in behind Cuttlefish's
 * Texture::convert() / Converter interface.
 *
 * Every entry point is plain C: pointers, sizes and ints.  No C++/torch types cross it.
 * The library owns device memory, streams and pinned staging; the caller owns host buffers.
 * There is NO CPU fallback: if no sm_100 device is usable every encode call fails with
 * CFX_ERR_NO_DEVICE / CFX_ERR_CUDA and cfx_last_error() says why.
 *
 * Which reference interface each entry point replaces (paths relative to the Cuttlefish tree):
 *   cfx_format_supported  <- createConverter() returning nullptr for an unsupported (format,type)
 *                            lib/src/Converter.cpp:32-506, failure contract :530-536
 *   cfx_block_info /
 *   cfx_encoded_size      <- Texture::blockWidth/blockHeight/blockSize, lib/src/Texture.cpp:529-773;
 *                            S3tcConverter ctor data().resize(), lib/src/S3tcConverter.cpp:230-240
 *   cfx_encode            <- the whole per-surface job loop of Converter::convert(),
 *                            lib/src/Converter.cpp:538-587, i.e. every Converter::process(x,y)
 *                            call for one surface (lib/src/S3tcConverter.cpp:242-255,
 *                            lib/src/EtcConverter.cpp:120-152, lib/src/AstcConverter.cpp:208-230)
 *   cfx_encode_batch      <- the mip/depth/face loop around it, lib/src/Converter.cpp:521-527
 *   cfx_encode_device     <- same as cfx_encode for callers that already hold the surface in HBM
 *   cfx_resize            <- Image::resize() on an RGBAF image, lib/src/Image.cpp:1324-1379 (FreeImage_Rescale,
 *                            lib/FreeImage/Source/FreeImageToolkit/Resize.cpp:140-218, :232-506, :1236-1273,
 *                            :2070-2112): bit-identical for linear images
 *   cfx_mip_levels        <- Texture::maxMipmapLevels() for a 2D texture, lib/src/Texture.cpp:514-527
 *   cfx_encode_mip_chain  <- Texture::generateMipmaps() (2D: each level resized from the one above,
 *                            lib/src/Texture.cpp:1457-1511) followed by Texture::convert(): level 0 is
 *                            uploaded once, the chain never leaves the GPU
 *   cfx_encode_mip_chain_device <- the same for callers that already hold level 0 in HBM
 *   cfx_init/cfx_shutdown <- the one-time encoder table inits (rgbcx::init, bc7enc_compress_block_init,
 *                            astcenc context alloc), lib/src/S3tcConverter.cpp:54-64,158-168
 *   cfx_init_devices /
 *   cfx_set_devices       <- the `threadCount` argument of Converter::convert(), lib/src/Converter.cpp:508-509,
 *                            :548-583: how wide one convert() call fans out -- here over the GPUs of one box
 *                            (contiguous block-row ranges per GPU) instead of over std::threads
 */
#ifndef CFX_H
#define CFX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cuttlefish::Texture::Format, lib/include/cuttlefish/Texture.h:59-130 (same numeric values). */
enum {
    CFX_FORMAT_BC1_RGB = 29, CFX_FORMAT_BC1_RGBA = 30, CFX_FORMAT_BC2 = 31, CFX_FORMAT_BC3 = 32,
    CFX_FORMAT_BC4 = 33, CFX_FORMAT_BC5 = 34, CFX_FORMAT_BC6H = 35, CFX_FORMAT_BC7 = 36,
    CFX_FORMAT_ETC1 = 37, CFX_FORMAT_ETC2_R8G8B8 = 38, CFX_FORMAT_ETC2_R8G8B8A1 = 39,
    CFX_FORMAT_ETC2_R8G8B8A8 = 40, CFX_FORMAT_EAC_R11 = 41, CFX_FORMAT_EAC_R11G11 = 42,
    CFX_FORMAT_ASTC_4x4 = 43, CFX_FORMAT_ASTC_5x4 = 44, CFX_FORMAT_ASTC_5x5 = 45,
    CFX_FORMAT_ASTC_6x5 = 46, CFX_FORMAT_ASTC_6x6 = 47, CFX_FORMAT_ASTC_8x5 = 48,
    CFX_FORMAT_ASTC_8x6 = 49, CFX_FORMAT_ASTC_8x8 = 50, CFX_FORMAT_ASTC_10x5 = 51,
    CFX_FORMAT_ASTC_10x6 = 52, CFX_FORMAT_ASTC_10x8 = 53, CFX_FORMAT_ASTC_10x10 = 54,
    CFX_FORMAT_ASTC_12x10 = 55, CFX_FORMAT_ASTC_12x12 = 56
};
/* Texture::Type :135-143, Texture::Alpha :161-167, Texture::Quality :181-188. */
enum { CFX_TYPE_UNORM = 0, CFX_TYPE_SNORM = 1, CFX_TYPE_UINT = 2, CFX_TYPE_INT = 3,
       CFX_TYPE_UFLOAT = 4, CFX_TYPE_FLOAT = 5 };
enum { CFX_ALPHA_NONE = 0, CFX_ALPHA_STANDARD = 1, CFX_ALPHA_PREMULTIPLIED = 2, CFX_ALPHA_ENCODED = 3 };
enum { CFX_QUALITY_LOWEST = 0, CFX_QUALITY_LOW = 1, CFX_QUALITY_NORMAL = 2, CFX_QUALITY_HIGH = 3,
       CFX_QUALITY_HIGHEST = 4 };
/* Texel layout of the source surface handed to the encoder. RGBA32F is what Cuttlefish's
 * Converter holds (Image::Format::RGBAF, lib/src/Converter.h:52-56); the kernels do the
 * float->u8 / float->half step of lib/src/S3tcConverter.cpp:97-129 themselves. */
enum { CFX_SRC_RGBA8 = 0, CFX_SRC_RGBA16F = 1, CFX_SRC_RGBA32F = 2 };

/* cuttlefish::Image::ResizeFilter, lib/include/cuttlefish/Image.h:79-86 (same numeric values). */
enum { CFX_FILTER_BOX = 0, CFX_FILTER_LINEAR = 1, CFX_FILTER_CUBIC = 2, CFX_FILTER_CATMULL_ROM = 3,
       CFX_FILTER_BSPLINE = 4 };

enum {
    CFX_OK = 0,
    CFX_ERR_INVALID = -1,      /* bad descriptor / null pointer / dst too small          */
    CFX_ERR_UNSUPPORTED = -2,  /* (format,type) has no GPU encoder: host should treat it
                                  like createConverter() == nullptr                       */
    CFX_ERR_NO_DEVICE = -3,    /* no CUDA device, or not compute capability 10.x         */
    CFX_ERR_CUDA = -4          /* a CUDA runtime call failed; see cfx_last_error()       */
};

typedef struct cfx_surface_desc {
    uint32_t format;        /* CFX_FORMAT_*                                           */
    uint32_t type;          /* CFX_TYPE_*                                             */
    uint32_t quality;       /* CFX_QUALITY_*                                          */
    uint32_t alpha_type;    /* CFX_ALPHA_*                                            */
    uint32_t color_mask;    /* bit0..3 = r,g,b,a enabled (Texture::ColorMask)         */
    uint32_t color_space;   /* 0 linear, 1 sRGB (image().colorSpace())                */
    uint32_t width, height; /* texels                                                 */
    uint32_t src_format;    /* CFX_SRC_*                                              */
    uint32_t flags;         /* CFX_FLAG_* bits, 0 by default                          */
    uint64_t src_row_pitch; /* bytes between stored rows (> 0). Stored row 0 is the top
                               row of the image, unless CFX_FLAG_BOTTOM_UP is set.     */
} cfx_surface_desc;

/* The surface is stored BOTTOM-UP: src points at the first stored row, which is the image's bottom row, and image row y
 * (0 = top) lives at src + (height-1-y)*src_row_pitch. This is how cuttlefish::Image keeps its pixels (FreeImage,
 * lib/src/Image.cpp:340-343, scanline(y) at :1092-1098), so the adapter hands the image over as it lies, without a
 * flipped host copy. Accepted by cfx_encode / cfx_encode_batch / cfx_encode_device (not by the mip-chain calls). */
#define CFX_FLAG_BOTTOM_UP 1u

/* The device pool: the GPUs that ONE host-buffer call (cfx_encode, cfx_encode_batch) spreads its work over. A surface
 * is cut into contiguous block-row ranges, range k of P going to pool device k (rows [k*R/P, (k+1)*R/P)); every device
 * uploads only its rows, encodes them and writes its packed blocks straight into the caller's dst at the range's byte
 * offset, so the result is the concatenation and no collective is needed. Surfaces too small to be worth cutting go
 * whole to the least loaded device. The pool is process-wide; one library call runs at a time.
 *
 *   cfx_init(device)        pool := {device}; device < 0 keeps the pool, or makes it {current CUDA device} when unset.
 *   cfx_init_devices(n)     pool := the first n sm_100 devices visible to the process; n = 0 means all of them.
 *   cfx_set_devices(n, ids) pool := ids[0..n), in this order. A device listed twice gets two independent contexts.
 *   cfx_device_count()      number of pool entries (0 before the first init).
 * All are idempotent, create streams/buffers on first use, return CFX_OK or an error, and leave the caller's current
 * CUDA device unchanged. Host-buffer calls without any init behave like cfx_init(-1). Device-pointer calls
 * (cfx_encode_device, cfx_encode_mip_chain_device) ignore the pool: they run on the device that owns d_src. */
int cfx_init(int device);
int cfx_init_devices(int device_count);
int cfx_set_devices(int n, const int* devices);
int cfx_device_count(void);
void cfx_shutdown(void);

/* 1 if the (format,type) pair has a GPU encoder, else 0. */
int cfx_format_supported(uint32_t format, uint32_t type);
/* 1 if the GPU encoder's bytes are IDENTICAL to the reference CPU encoder's for this (format, type,
 * quality) -- BC4/BC5 UNorm always; BC1_RGB/BC2/BC3 at every quality level when the library was built with
 * the reference's rgbcx tables (tools/gen_rgbcx_tables.py); ETC1 at every quality level, linear and sRGB
 * colour space -- else 0: the format is held to PSNR parity.
 * No reference analogue; lets an integrator (and the tests) know which guarantee applies. */
int cfx_format_is_exact(uint32_t format, uint32_t type, uint32_t quality);
/* Block footprint and bytes per block; returns CFX_OK or CFX_ERR_UNSUPPORTED. */
int cfx_block_info(uint32_t format, uint32_t* block_w, uint32_t* block_h, uint32_t* block_bytes);
/* ceil(w/bw)*ceil(h/bh)*block_bytes, 0 if the format is unknown. */
size_t cfx_encoded_size(const cfx_surface_desc* desc);

/* Encode one surface held in HOST memory into HOST memory (blocks row-major, y*blocksX+x), on every device of the
 * pool. Does H2D, kernels, D2H in overlapping chunks of block rows; returns when dst is complete. Pinned buffers
 * (cfx_host_alloc, cudaHostAlloc, cudaHostRegister) are DMA'd in place; a pageable src passes through the library's
 * pinned staging slots on a few host threads, and on that pass an RGBA32F source is narrowed to what the encoder's load
 * stage would make of it anyway (RGBA8 for BC1-5/BC7 UNorm, RGBA16F for BC6H: 4 or 8 instead of 16 bytes per texel over
 * PCIe, identical blocks). */
int cfx_encode(const cfx_surface_desc* desc, const void* src, void* dst, size_t dst_size);
/* Encode n surfaces (a mip chain / array layers); same semantics per surface. */
int cfx_encode_batch(int n, const cfx_surface_desc* descs, const void* const* srcs,
                     void* const* dsts, const size_t* dst_sizes);
/* Encode one surface already resident in DEVICE memory into DEVICE memory, asynchronously on
 * cuda_stream (a cudaStream_t cast to void*; NULL = the CUDA default stream), on the device that owns d_src.
 * d_src and the pitch must be multiples of the texel size (4 / 8 / 16 bytes; CFX_ERR_INVALID otherwise); a 16-byte
 * aligned pointer and pitch take the vectorised load path. */
int cfx_encode_device(const cfx_surface_desc* desc, const void* d_src, void* d_dst, size_t dst_size,
                      void* cuda_stream);

/* Resize one RGBA32F surface (host memory in, host memory out; rows top-down, pitches in bytes) the way
 * Image::resize() does: FreeImage's separable filter with double-precision weights and accumulation, the pass
 * along x first unless the width grows; color_space 1 (sRGB) filters in linear space. Linear results are
 * bit-identical to the reference's; sRGB ones differ only where pow() rounds differently. */
int cfx_resize(const void* src, uint32_t src_width, uint32_t src_height, size_t src_row_pitch,
               void* dst, uint32_t dst_width, uint32_t dst_height, size_t dst_row_pitch,
               uint32_t filter, uint32_t color_space);
/* floor(log2(max(width, height))) + 1: the length of a full 2D mip chain. */
uint32_t cfx_mip_levels(uint32_t width, uint32_t height);
/* generateMipmaps(filter, levels) + convert() for one 2D surface. level0 describes the level-0 image (src): RGBA32F,
 * or RGBA8 taken as (float)v/255 -- what Image::convert(RGBAF) makes of an 8-bit image (FreeImage_ConvertToRGBAF), so
 * an 8-bit source need not be widened on the host; level k has size max(1, w >> k) x max(1, h >> k), is resized on the GPU from level k-1 and encoded
 * with level0's format/type/quality/...; dsts[k] / dst_sizes[k] receive its blocks (k = 0 .. levels-1).
 * mip_images, if not NULL, is an array of `levels` host pointers (entries may be NULL; entry 0 is ignored)
 * that receive the generated RGBA32F levels, tightly packed rows, top-down. */
int cfx_encode_mip_chain(const cfx_surface_desc* level0, const void* src, uint32_t filter, uint32_t levels,
                         void* const* dsts, const size_t* dst_sizes, void* const* mip_images);

/* The same for a level 0 already resident in DEVICE memory (texel aligned: 16 bytes for RGBA32F, 4 for RGBA8), blocks written
 * to DEVICE buffers d_dsts[k], everything queued on cuda_stream. The generated levels live in library-owned memory
 * that the next chain call reuses: one chain in flight per context. */
int cfx_encode_mip_chain_device(const cfx_surface_desc* level0, const void* d_src, uint32_t filter, uint32_t levels,
                                void* const* d_dsts, const size_t* dst_sizes, void* cuda_stream);

/* Containers around the packed blocks, 2D textures and 2D arrays.
 * cfx_dds_header / cfx_ktx_header write the file header the reference's writers produce for a texture of level0's format /
 * type / colour space / alpha type and size with `mip_levels` levels (array_size 0 = not an array) -- saveDds(),
 * lib/src/SaveDds.cpp:565-683 (magic + DDS_HEADER + DX10 header = 148 bytes), saveKtx(), lib/src/SaveKtx.cpp:1189-1214
 * (64 bytes; every level is then a 32-bit imageSize followed by its blocks) -- and return its size, or 0 when the
 * reference has no such file: DDS knows no ETC / EAC / ASTC format (isValidForDds), and Texture::convert() refuses an sRGB
 * image for formats without an sRGB variant (BC4, BC5, BC6H, ETC1, EAC, ASTC HDR). `out` must hold 148 bytes.
 * cfx_encode_mip_chain_to_file = generateMipmaps(filter, levels) + convert() + save(): the header is written into a
 * mapping of the output file and every level's blocks are delivered straight into that mapping at their final offset
 * (levels = 0: the full chain). Returns CFX_OK, CFX_ERR_UNSUPPORTED (no such container format / no GPU encoder) or the
 * encode's error; the file is removed on failure. */
enum { CFX_CONTAINER_DDS = 0, CFX_CONTAINER_KTX = 1 };
size_t cfx_dds_header(const cfx_surface_desc* level0, uint32_t mip_levels, uint32_t array_size, void* out);
size_t cfx_ktx_header(const cfx_surface_desc* level0, uint32_t mip_levels, uint32_t array_size, void* out);
int cfx_encode_mip_chain_to_file(const cfx_surface_desc* level0, const void* src, uint32_t filter, uint32_t levels,
                                 uint32_t container, const char* path);

/* Cross-process peer memory, for hosts that run one PROCESS per GPU (bench.py under torchrun): the process
 * that is to own the assembled output exports the device buffer, the others open it and pass the mapped pointer plus
 * their slab's byte offset as d_dst of cfx_encode_device(). The encode kernel then stores its packed blocks straight
 * into the owner's HBM over NVLink: the gather of Converter::convert()'s `textureData[mip][d][f] = data()`
 * (lib/src/Converter.cpp:587) is fused into the kernel -- no collective, no extra launch, no staging copy.
 *   cfx_ipc_export(d_ptr, handle)        handle: CFX_IPC_HANDLE_BYTES bytes, to be sent to the other processes.
 *   cfx_ipc_open(handle, device, &ptr)   maps the buffer for kernels of `device` (peer access is enabled on the way).
 *   cfx_ipc_close(ptr, handle)           unmaps it.
 * The owner learns that a slab is complete the way it would for any peer write: a stream-ordered signal after the
 * kernel (an interprocess event, or the tiny collective bench.py uses). Same-process multi-GPU hosts need none of this:
 * cfx_encode() already writes every device's blocks into the one caller-owned buffer. */
#define CFX_IPC_HANDLE_BYTES 96
int cfx_ipc_export(const void* d_ptr, void* handle);
int cfx_ipc_open(const void* handle, int device, void** d_ptr);
int cfx_ipc_close(void* d_ptr, const void* handle);

/* Pinned host memory helpers for callers that want zero-copy staging. */
void* cfx_host_alloc(size_t bytes);
void cfx_host_free(void* p);

/* Number of encoder kernel launches issued by this process so far. */
uint64_t cfx_kernel_launches(void);
/* Thread-local description of the last failure ("" if none). */
const char* cfx_last_error(void);
const char* cfx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CFX_H */
