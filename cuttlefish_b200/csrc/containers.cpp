// DDS and KTX containers around the packed blocks: the headers of the reference's writers for
// 2D textures and 2D arrays -- saveDds(), lib/src/SaveDds.cpp:565-683 (DX10 header, format table :436-551) and
// saveKtx(), lib/src/SaveKtx.cpp:1189-1290 (format table :689-1180) -- and one call that generates the mip chain, encodes
// it and lands every level's blocks directly in a mapping of the output file: nothing is assembled in host memory first.
// Host code only.  Byte-identical to the reference's files wherever the encoder is (tests/golden/containers/, written by
// the reference's real Texture::save()).
#include "../../include/cfx.h"

#include <cstdio>
#include <cstring>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

namespace {

bool has_alpha(uint32_t f)             // Texture::hasAlpha(), lib/src/Texture.cpp:470-512 (block formats)
{
    return f == CFX_FORMAT_BC1_RGBA || f == CFX_FORMAT_BC2 || f == CFX_FORMAT_BC3 || f == CFX_FORMAT_BC7 ||
        f == CFX_FORMAT_ETC2_R8G8B8A1 || f == CFX_FORMAT_ETC2_R8G8B8A8 || (f >= CFX_FORMAT_ASTC_4x4 && f <= CFX_FORMAT_ASTC_12x12);
}

// Formats whose sRGB variant exists; Texture::convert() refuses an sRGB image for the others.
bool srgb_ok(uint32_t f, uint32_t type)
{
    if (f >= CFX_FORMAT_ASTC_4x4 && f <= CFX_FORMAT_ASTC_12x12) return type == CFX_TYPE_UNORM;
    return f == CFX_FORMAT_BC1_RGB || f == CFX_FORMAT_BC1_RGBA || f == CFX_FORMAT_BC2 || f == CFX_FORMAT_BC3 || f == CFX_FORMAT_BC7 ||
        f == CFX_FORMAT_ETC2_R8G8B8 || f == CFX_FORMAT_ETC2_R8G8B8A1 || f == CFX_FORMAT_ETC2_R8G8B8A8;
}

uint32_t dxgi_format(uint32_t f, uint32_t type, bool srgb)      // getDdsFormat(), SaveDds.cpp:436-551
{
    switch (f) {
        case CFX_FORMAT_BC1_RGB: case CFX_FORMAT_BC1_RGBA: return type == CFX_TYPE_UNORM ? (srgb ? 72u : 71u) : 0u;
        case CFX_FORMAT_BC2: return type == CFX_TYPE_UNORM ? (srgb ? 75u : 74u) : 0u;
        case CFX_FORMAT_BC3: return type == CFX_TYPE_UNORM ? (srgb ? 78u : 77u) : 0u;
        case CFX_FORMAT_BC4: return type == CFX_TYPE_UNORM ? 80u : (type == CFX_TYPE_SNORM ? 81u : 0u);
        case CFX_FORMAT_BC5: return type == CFX_TYPE_UNORM ? 83u : (type == CFX_TYPE_SNORM ? 84u : 0u);
        case CFX_FORMAT_BC6H: return type == CFX_TYPE_UFLOAT ? 95u : (type == CFX_TYPE_FLOAT ? 96u : 0u);
        case CFX_FORMAT_BC7: return type == CFX_TYPE_UNORM ? (srgb ? 99u : 98u) : 0u;
        default: return 0u;                                         // ETC, EAC, ASTC: no DXGI format
    }
}

// glInternalFormat / glBaseInternalFormat, getFormatInfo(), SaveKtx.cpp:689-1180
bool gl_format(uint32_t f, uint32_t type, bool srgb, uint32_t& internal, uint32_t& base)
{
    const uint32_t GL_RED = 0x1903, GL_RG = 0x8227, GL_RGB = 0x1907, GL_RGBA = 0x1908;
    const bool un = type == CFX_TYPE_UNORM, sn = type == CFX_TYPE_SNORM;
    if (f >= CFX_FORMAT_ASTC_4x4 && f <= CFX_FORMAT_ASTC_12x12) {
        if (!(un || type == CFX_TYPE_UFLOAT)) return false;
        base = GL_RGBA; internal = (srgb ? 0x93D0u : 0x93B0u) + (f - CFX_FORMAT_ASTC_4x4);
        return true;
    }
    switch (f) {
        case CFX_FORMAT_BC1_RGB: base = GL_RGB; internal = srgb ? 0x8C4Cu : 0x83F0u; return un;
        case CFX_FORMAT_BC1_RGBA: base = GL_RGBA; internal = srgb ? 0x8C4Du : 0x83F1u; return un;
        case CFX_FORMAT_BC2: base = GL_RGBA; internal = srgb ? 0x8C4Eu : 0x83F2u; return un;
        case CFX_FORMAT_BC3: base = GL_RGBA; internal = srgb ? 0x8C4Fu : 0x83F3u; return un;
        case CFX_FORMAT_BC4: base = GL_RED; internal = un ? 0x8DBBu : 0x8DBCu; return un || sn;
        case CFX_FORMAT_BC5: base = GL_RG; internal = un ? 0x8DBDu : 0x8DBEu; return un || sn;
        case CFX_FORMAT_BC6H: base = GL_RGB; internal = type == CFX_TYPE_UFLOAT ? 0x8E8Fu : 0x8E8Eu; return type == CFX_TYPE_UFLOAT || type == CFX_TYPE_FLOAT;
        case CFX_FORMAT_BC7: base = GL_RGBA; internal = srgb ? 0x8E8Du : 0x8E8Cu; return un;
        case CFX_FORMAT_ETC1: base = GL_RGB; internal = 0x8D64u; return un;
        case CFX_FORMAT_ETC2_R8G8B8: base = GL_RGB; internal = srgb ? 0x9275u : 0x9274u; return un;
        case CFX_FORMAT_ETC2_R8G8B8A1: base = GL_RGBA; internal = srgb ? 0x9277u : 0x9276u; return un;
        case CFX_FORMAT_ETC2_R8G8B8A8: base = GL_RGBA; internal = srgb ? 0x9279u : 0x9278u; return un;
        case CFX_FORMAT_EAC_R11: base = GL_RED; internal = un ? 0x9270u : 0x9271u; return un || sn;
        case CFX_FORMAT_EAC_R11G11: base = GL_RG; internal = un ? 0x9272u : 0x9273u; return un || sn;
        default: return false;
    }
}

void put32(uint8_t* p, uint32_t v) { std::memcpy(p, &v, 4); }

} // namespace

extern "C" {

size_t cfx_dds_header(const cfx_surface_desc* d, uint32_t mip_levels, uint32_t array_size, void* out)
{
    if (!d || !out || mip_levels == 0) return 0;
    const bool srgb = d->color_space != 0;
    if (srgb && !srgb_ok(d->format, d->type)) return 0;
    const uint32_t dxgi = dxgi_format(d->format, d->type, srgb);
    uint32_t bw, bh, bytes;
    if (!dxgi || cfx_block_info(d->format, &bw, &bh, &bytes) != CFX_OK) return 0;
    uint8_t* h = static_cast<uint8_t*>(out);
    std::memset(h, 0, 148);
    put32(h, 0x20534444u);                                   // "DDS "
    put32(h + 4, 124);                                       // DdsHeader::size
    put32(h + 8, 0x1007u | 0x20000u | 0x8u);                 // caps | height | width | pixel format | mip count | pitch
    put32(h + 12, d->height); put32(h + 16, d->width);
    put32(h + 20, (d->width + bw - 1)/bw*bytes);             // computePitch(): one block row
    put32(h + 28, mip_levels);
    put32(h + 76, 32); put32(h + 80, 0x4u); std::memcpy(h + 84, "DX10", 4);
    const bool is_array = array_size > 0;
    put32(h + 108, 0x1000u | (mip_levels > 1 ? 0x400000u : 0u) | (mip_levels > 1 || is_array ? 0x8u : 0u));
    put32(h + 128, dxgi);
    put32(h + 132, 3);                                       // DdsTextureDim_TEXTURE2D
    put32(h + 140, is_array ? array_size : 1u);              // Texture::depth() of a plain 2D texture is 1
    uint32_t alpha_mode = 3;                                 // opaque
    if (has_alpha(d->format))
        alpha_mode = d->alpha_type == CFX_ALPHA_NONE ? 3u : (d->alpha_type == CFX_ALPHA_STANDARD ? 1u : (d->alpha_type == CFX_ALPHA_PREMULTIPLIED ? 2u : 4u));
    put32(h + 144, alpha_mode);
    return 148;
}

size_t cfx_ktx_header(const cfx_surface_desc* d, uint32_t mip_levels, uint32_t array_size, void* out)
{
    if (!d || !out || mip_levels == 0) return 0;
    const bool srgb = d->color_space != 0;
    if (srgb && !srgb_ok(d->format, d->type)) return 0;
    uint32_t internal = 0, base = 0;
    if (!gl_format(d->format, d->type, srgb, internal, base)) return 0;
    static const uint8_t id[12] = {0xAB, 0x4B, 0x54, 0x58, 0x20, 0x31, 0x31, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A};
    uint8_t* h = static_cast<uint8_t*>(out);
    std::memcpy(h, id, 12);
    const uint32_t v[13] = {0x04030201u, 0u /* glType */, 1u /* glTypeSize */, 0u /* glFormat */, internal, base,
        d->width, d->height, 0u /* depth */, array_size, 1u /* faces */, mip_levels, 0u /* key/value bytes */};
    std::memcpy(h + 12, v, sizeof(v));
    return 64;
}

int cfx_encode_mip_chain_to_file(const cfx_surface_desc* level0, const void* src, uint32_t filter, uint32_t levels,
    uint32_t container, const char* path)
{
    if (!level0 || !src || !path || container > CFX_CONTAINER_KTX) return CFX_ERR_INVALID;
    const uint32_t max_levels = cfx_mip_levels(level0->width, level0->height);
    if (levels == 0 || levels > max_levels) levels = max_levels;
    uint8_t header[148];
    const size_t hbytes = container == CFX_CONTAINER_DDS ? cfx_dds_header(level0, levels, 0, header) : cfx_ktx_header(level0, levels, 0, header);
    if (!hbytes) return CFX_ERR_UNSUPPORTED;
    if (!cfx_format_supported(level0->format, level0->type)) return CFX_ERR_UNSUPPORTED;
    // file layout: DDS = header, then the levels back to back; KTX = header, then per level a 32-bit imageSize and the
    // blocks (block formats are always a multiple of 4 bytes: no padding)
    std::vector<size_t> sizes(levels), offsets(levels);
    size_t total = hbytes;
    for (uint32_t k = 0; k < levels; ++k) {
        cfx_surface_desc d = *level0;
        d.width = level0->width >> k ? level0->width >> k : 1u;
        d.height = level0->height >> k ? level0->height >> k : 1u;
        sizes[k] = cfx_encoded_size(&d);
        if (container == CFX_CONTAINER_KTX) total += 4;
        offsets[k] = total;
        total += sizes[k];
    }
    const int fd = ::open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return CFX_ERR_INVALID;
    int rc = CFX_ERR_INVALID;
    if (::ftruncate(fd, static_cast<off_t>(total)) == 0) {
        void* map = ::mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        if (map != MAP_FAILED) {
            uint8_t* base = static_cast<uint8_t*>(map);
            std::memcpy(base, header, hbytes);
            std::vector<void*> dsts(levels);
            for (uint32_t k = 0; k < levels; ++k) {
                dsts[k] = base + offsets[k];
                if (container == CFX_CONTAINER_KTX) put32(base + offsets[k] - 4, static_cast<uint32_t>(sizes[k]));
            }
            rc = cfx_encode_mip_chain(level0, src, filter, levels, dsts.data(), sizes.data(), nullptr);
            ::munmap(map, total);
        }
    }
    ::close(fd);
    if (rc != CFX_OK) ::unlink(path);
    return rc;
}

} // extern "C"
