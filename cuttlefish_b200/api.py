"""Host API over libcfx.so: enums (same numeric values as cuttlefish::Texture), array-level
encode calls, block-row sharding for multi-GPU runs, and a Texture class mirroring the part of
cuttlefish::Texture that the convert path touches."""
import ctypes
import os

import numpy as np

from ._lib import SurfaceDesc, load

# cuttlefish::Texture::Format (lib/include/cuttlefish/Texture.h:59-130)
FORMATS = {
    "BC1_RGB": 29, "BC1_RGBA": 30, "BC2": 31, "BC3": 32, "BC4": 33, "BC5": 34, "BC6H": 35, "BC7": 36,
    "ETC1": 37, "ETC2_R8G8B8": 38, "ETC2_R8G8B8A1": 39, "ETC2_R8G8B8A8": 40, "EAC_R11": 41,
    "EAC_R11G11": 42, "ASTC_4x4": 43, "ASTC_5x4": 44, "ASTC_5x5": 45, "ASTC_6x5": 46, "ASTC_6x6": 47,
    "ASTC_8x5": 48, "ASTC_8x6": 49, "ASTC_8x8": 50, "ASTC_10x5": 51, "ASTC_10x6": 52, "ASTC_10x8": 53,
    "ASTC_10x10": 54, "ASTC_12x10": 55, "ASTC_12x12": 56,
}
TYPES = {"UNorm": 0, "SNorm": 1, "UInt": 2, "Int": 3, "UFloat": 4, "Float": 5}
QUALITY = {"Lowest": 0, "Low": 1, "Normal": 2, "High": 3, "Highest": 4}
ALPHA = {"None": 0, "Standard": 1, "PreMultiplied": 2, "Encoded": 3}
SRC_FORMATS = {"RGBA8": 0, "RGBA16F": 1, "RGBA32F": 2}
# cuttlefish::Image::ResizeFilter (lib/include/cuttlefish/Image.h:79-86)
FILTERS = {"Box": 0, "Linear": 1, "Cubic": 2, "CatmullRom": 3, "BSpline": 4}

CFX_ERR_UNSUPPORTED = -2


class CfxError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("cfx error %d: %s" % (code, message))
        self.code = code


class ColorMask:
    """cuttlefish::Texture::ColorMask (Texture.h:216-237)."""

    def __init__(self, r=True, g=True, b=True, a=True):
        self.r, self.g, self.b, self.a = bool(r), bool(g), bool(b), bool(a)

    def bits(self):
        return (1 if self.r else 0) | (2 if self.g else 0) | (4 if self.b else 0) | (8 if self.a else 0)


def _enum(table, v):
    return table[v] if isinstance(v, str) else int(v)


def _mask_bits(m):
    return m.bits() if isinstance(m, ColorMask) else int(m)


def _check(rc):
    if rc != 0:
        raise CfxError(rc, load().cfx_last_error().decode())


def init(device=-1):
    """cfx_init: the device pool of host-buffer calls becomes {device} (device < 0: keep it / the current CUDA device)."""
    _check(load().cfx_init(int(device)))


def init_devices(count=0):
    """cfx_init_devices: host-buffer calls shard every surface by block row over the first `count` sm_100 devices
    (0 = all visible) -- the multi-GPU form of Converter::convert's threadCount."""
    _check(load().cfx_init_devices(int(count)))


def set_devices(devices):
    """cfx_set_devices: explicit pool, in order (a device listed twice gets two independent contexts)."""
    ids = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
    _check(load().cfx_set_devices(len(devices), ids))


def device_count():
    return int(load().cfx_device_count())


def shutdown():
    load().cfx_shutdown()


def version():
    return load().cfx_version().decode()


def last_error():
    return load().cfx_last_error().decode()


def kernel_launches():
    return int(load().cfx_kernel_launches())


def format_supported(fmt, type="UNorm"):
    return bool(load().cfx_format_supported(_enum(FORMATS, fmt), _enum(TYPES, type)))


def format_is_exact(fmt, type="UNorm", quality="Normal"):
    """True if the GPU bytes are identical to the reference CPU encoder's (else: PSNR parity)."""
    return bool(load().cfx_format_is_exact(_enum(FORMATS, fmt), _enum(TYPES, type), _enum(QUALITY, quality)))


def block_info(fmt):
    bw, bh, nb = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
    rc = load().cfx_block_info(_enum(FORMATS, fmt), ctypes.byref(bw), ctypes.byref(bh), ctypes.byref(nb))
    if rc != 0:
        raise CfxError(rc, "format %r is not block compressed" % (fmt,))
    return bw.value, bh.value, nb.value


FLAG_BOTTOM_UP = 1


def make_desc(fmt, width, height, src_format, row_pitch, type="UNorm", quality="Normal", alpha="Standard",
              color_mask=15, srgb=False, bottom_up=False):
    return SurfaceDesc(_enum(FORMATS, fmt), _enum(TYPES, type), _enum(QUALITY, quality), _enum(ALPHA, alpha),
                       _mask_bits(color_mask), 1 if srgb else 0, int(width), int(height),
                       _enum(SRC_FORMATS, src_format), FLAG_BOTTOM_UP if bottom_up else 0, int(row_pitch))


def encoded_size(fmt, width, height):
    d = make_desc(fmt, width, height, "RGBA8", width * 4)
    return int(load().cfx_encoded_size(ctypes.byref(d)))


def _src_format_of(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.uint8:
        return "RGBA8", 4
    if dtype == np.float16:
        return "RGBA16F", 8
    if dtype == np.float32:
        return "RGBA32F", 16
    raise TypeError("source texels must be uint8, float16 or float32 RGBA, got %s" % dtype)


def encode(img, fmt, out=None, **kw):
    """Encode a HOST image [H,W,4] (uint8 / float16 / float32, row 0 = top; bottom_up=True: row 0 = bottom, the way
    cuttlefish::Image stores it) -> uint8 block bytes.

    Goes through cfx_encode: host->device copy, kernels, device->host copy, on every device of the pool."""
    img = np.asarray(img)
    if img.ndim != 3 or img.shape[2] != 4:
        raise ValueError("expected [H,W,4] RGBA texels")
    if not img.flags["C_CONTIGUOUS"]:
        img = np.ascontiguousarray(img)
    h, w, _ = img.shape
    src_format, texel = _src_format_of(img.dtype)
    d = make_desc(fmt, w, h, src_format, w * texel, **kw)
    n = int(load().cfx_encoded_size(ctypes.byref(d)))
    if n == 0:
        raise CfxError(CFX_ERR_UNSUPPORTED, "format %r is not block compressed" % (fmt,))
    if out is None:
        out = np.empty(n, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size >= n and out.flags["C_CONTIGUOUS"]
    _check(load().cfx_encode(ctypes.byref(d), img.ctypes.data, out.ctypes.data, out.size))
    return out[:n]


def encode_batch(images, fmt, outs=None, **kw):
    """Encode several HOST surfaces (a mip chain / array layers) with one cfx_encode_batch call, on every device of
    the pool. Returns a list of uint8 arrays, one per surface (`outs`: caller-owned, e.g. pinned, output arrays)."""
    lib = load()
    given, imgs, descs, outs = outs, [], [], []
    for i, img in enumerate(images):
        img = np.ascontiguousarray(np.asarray(img))
        if img.ndim != 3 or img.shape[2] != 4:
            raise ValueError("expected [H,W,4] RGBA texels")
        h, w, _ = img.shape
        src_format, texel = _src_format_of(img.dtype)
        d = make_desc(fmt, w, h, src_format, w * texel, **kw)
        n = int(lib.cfx_encoded_size(ctypes.byref(d)))
        if n == 0:
            raise CfxError(CFX_ERR_UNSUPPORTED, "format %r is not block compressed" % (fmt,))
        if given is None:
            out = np.empty(n, dtype=np.uint8)
        else:
            out = given[i]
            assert out.dtype == np.uint8 and out.size >= n and out.flags["C_CONTIGUOUS"]
            out = out[:n]
        imgs.append(img); descs.append(d); outs.append(out)
    n = len(imgs)
    c_descs = (SurfaceDesc * n)(*descs)
    c_srcs = (ctypes.c_void_p * n)(*[im.ctypes.data for im in imgs])
    c_dsts = (ctypes.c_void_p * n)(*[o.ctypes.data for o in outs])
    c_sizes = (ctypes.c_size_t * n)(*[o.size for o in outs])
    _check(lib.cfx_encode_batch(n, c_descs, c_srcs, c_dsts, c_sizes))
    return outs


IPC_HANDLE_BYTES = 96


def ipc_export(tensor):
    """cfx_ipc_export: a handle (bytes) other processes can open to write into this CUDA tensor's memory."""
    buf = ctypes.create_string_buffer(IPC_HANDLE_BYTES)
    _check(load().cfx_ipc_export(ctypes.c_void_p(tensor.data_ptr()), buf))
    return buf.raw


def ipc_open(handle, device):
    """cfx_ipc_open: maps an exported buffer for kernels running on `device`; returns the device pointer (int)."""
    ptr = ctypes.c_void_p()
    _check(load().cfx_ipc_open(handle, int(device), ctypes.byref(ptr)))
    return int(ptr.value)


def ipc_close(ptr, handle):
    _check(load().cfx_ipc_close(ctypes.c_void_p(ptr), handle))


def encode_device(src, fmt, out=None, stream=None, out_ptr=None, out_bytes=0, **kw):
    """Encode a torch CUDA tensor [H,W,4] (uint8/float16/float32) into a CUDA uint8 tensor,
    asynchronously on `stream` (default: torch's current stream). Goes through cfx_encode_device.
    out_ptr / out_bytes: a raw device pointer instead of `out`, e.g. peer memory opened with ipc_open()."""
    import torch
    if not src.is_cuda:
        raise ValueError("encode_device needs a CUDA tensor; use encode() for host arrays")
    if src.dim() != 3 or src.shape[2] != 4 or src.stride(2) != 1 or src.stride(1) != 4:
        raise ValueError("expected [H,W,4] texels with contiguous rows")
    h, w, _ = src.shape
    table = {torch.uint8: ("RGBA8", 4), torch.float16: ("RGBA16F", 8), torch.float32: ("RGBA32F", 16)}
    if src.dtype not in table:
        raise TypeError("unsupported texel dtype %s" % src.dtype)
    src_format, texel = table[src.dtype]
    pitch = src.stride(0) * src.element_size()
    d = make_desc(fmt, w, h, src_format, pitch, **kw)
    n = int(load().cfx_encoded_size(ctypes.byref(d)))
    if n == 0:
        raise CfxError(CFX_ERR_UNSUPPORTED, "format %r is not block compressed" % (fmt,))
    if out_ptr is not None:
        assert out_bytes >= n
        dst, cap = int(out_ptr), int(out_bytes)
    else:
        if out is None:
            out = torch.empty(n, dtype=torch.uint8, device=src.device)
        assert out.is_cuda and out.dtype == torch.uint8 and out.numel() >= n and out.is_contiguous()
        dst, cap = out.data_ptr(), out.numel()
    with torch.cuda.device(src.device):
        if stream is None:
            stream = torch.cuda.current_stream()
        _check(load().cfx_encode_device(ctypes.byref(d), src.data_ptr(), ctypes.c_void_p(dst), cap,
                                        ctypes.c_void_p(stream.cuda_stream)))
    return None if out_ptr is not None else out[:n]


def encode_mip_chain_device(src, fmt, filter="CatmullRom", levels=None, outs=None, stream=None, **kw):
    """generateMipmaps + convert for a float32 torch CUDA tensor [H,W,4]: the chain is made and encoded on the GPU,
    asynchronously on `stream`. Returns one CUDA uint8 tensor per level. Goes through cfx_encode_mip_chain_device."""
    import torch
    if not src.is_cuda or src.dtype not in (torch.float32, torch.uint8):
        raise ValueError("encode_mip_chain_device needs a float32 or uint8 CUDA tensor")
    if src.dim() != 3 or src.shape[2] != 4 or src.stride(2) != 1 or src.stride(1) != 4:
        raise ValueError("expected [H,W,4] texels with contiguous rows")
    h, w, _ = src.shape
    n = mip_levels(w, h)
    n = n if levels is None else max(1, min(int(levels), n))
    d = make_desc(fmt, w, h, "RGBA32F" if src.dtype == torch.float32 else "RGBA8", src.stride(0) * src.element_size(), **kw)
    sizes = [encoded_size(fmt, max(1, w >> k), max(1, h >> k)) for k in range(n)]
    if sizes[0] == 0:
        raise CfxError(CFX_ERR_UNSUPPORTED, "format %r is not block compressed" % (fmt,))
    if outs is None:
        outs = [torch.empty(sz, dtype=torch.uint8, device=src.device) for sz in sizes]
    assert len(outs) >= n and all(o.is_cuda and o.dtype == torch.uint8 and o.numel() >= sz for o, sz in zip(outs, sizes))
    dst = (ctypes.c_void_p * n)(*[o.data_ptr() for o in outs[:n]])
    csizes = (ctypes.c_size_t * n)(*[o.numel() for o in outs[:n]])
    with torch.cuda.device(src.device):
        if stream is None:
            stream = torch.cuda.current_stream()
        _check(load().cfx_encode_mip_chain_device(ctypes.byref(d), src.data_ptr(), _enum(FILTERS, filter), n, dst, csizes,
                                                  ctypes.c_void_p(stream.cuda_stream)))
    return [o[:sz] for o, sz in zip(outs, sizes)]


def shard_block_rows(height, block_h, rank, world):
    """Contiguous block-row range [r0, r1) owned by `rank` out of `world`, and the source rows
    [y0, y1) it needs. Slabs end on block-row boundaries, so each rank encodes its slab as an
    independent surface and the packed outputs concatenate."""
    rows = (int(height) + block_h - 1) // block_h
    r0 = rows * rank // world
    r1 = rows * (rank + 1) // world
    return r0, r1, r0 * block_h, min(int(height), r1 * block_h)


def resize(img, width, height, filter="CatmullRom", srgb=False):
    """Image::resize() (lib/src/Image.cpp:1324-1379) for a HOST float32 [H,W,4] image on the GPU: FreeImage's
    separable filter, bit-identical for linear images. Returns a float32 [height,width,4] array."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    if img.ndim != 3 or img.shape[2] != 4:
        raise ValueError("expected [H,W,4] RGBA texels")
    out = np.empty((int(height), int(width), 4), np.float32)
    _check(load().cfx_resize(img.ctypes.data, img.shape[1], img.shape[0], img.shape[1] * 16, out.ctypes.data,
                             int(width), int(height), int(width) * 16, _enum(FILTERS, filter), 1 if srgb else 0))
    return out


def mip_levels(width, height):
    """Texture::maxMipmapLevels() for a 2D texture (lib/src/Texture.cpp:514-527)."""
    return int(load().cfx_mip_levels(int(width), int(height)))


def encode_mip_chain(img, fmt, filter="CatmullRom", levels=None, return_images=False, outs=None, **kw):
    """Texture::generateMipmaps(filter, levels) + Texture::convert() for one HOST [H,W,4] surface (float32, or uint8 taken
    as v/255 like Image::convert(RGBAF)) with one cfx_encode_mip_chain call: level 0 is uploaded once, the mips are made
    and encoded on the GPU. Returns a list of uint8 block arrays (and, with return_images, the list of mip images)."""
    img = np.asarray(img)
    if img.dtype != np.uint8:
        img = img.astype(np.float32, copy=False)
    img = np.ascontiguousarray(img)
    if img.ndim != 3 or img.shape[2] != 4:
        raise ValueError("expected [H,W,4] RGBA texels")
    h, w, _ = img.shape
    src_format, texel = _src_format_of(img.dtype)
    n = mip_levels(w, h)
    n = n if levels is None else max(1, min(int(levels), n))
    d = make_desc(fmt, w, h, src_format, w * texel, **kw)
    given, outs, images = outs, [], [img]
    for k in range(n):
        dk = make_desc(fmt, max(1, w >> k), max(1, h >> k), "RGBA32F", 16, **kw)
        size = int(load().cfx_encoded_size(ctypes.byref(dk)))
        if size == 0:
            raise CfxError(CFX_ERR_UNSUPPORTED, "format %r is not block compressed" % (fmt,))
        if given is None:
            outs.append(np.empty(size, np.uint8))
        else:                                   # caller-owned (e.g. pinned) output buffers, one per level
            assert given[k].dtype == np.uint8 and given[k].size >= size and given[k].flags["C_CONTIGUOUS"]
            outs.append(given[k][:size])
        if return_images and k:
            images.append(np.empty((max(1, h >> k), max(1, w >> k), 4), np.float32))
    dst = (ctypes.c_void_p * n)(*[o.ctypes.data for o in outs])
    sizes = (ctypes.c_size_t * n)(*[o.size for o in outs])
    mips = (ctypes.c_void_p * n)(*([None] + [m.ctypes.data for m in images[1:]])) if return_images else None
    _check(load().cfx_encode_mip_chain(ctypes.byref(d), img.ctypes.data, _enum(FILTERS, filter), n, dst, sizes, mips))
    return (outs, images) if return_images else outs


CONTAINERS = {"DDS": 0, "KTX": 1}


def container_header(container, fmt, width, height, mip_levels=1, array_size=0, **kw):
    """The DDS (148 bytes) / KTX (64 bytes) file header the reference's Texture::save() writes for such a texture
    (lib/src/SaveDds.cpp:565-683, lib/src/SaveKtx.cpp:1189-1214); None when the reference has no such file."""
    d = make_desc(fmt, width, height, "RGBA32F", 16 * width, **kw)
    buf = np.zeros(148, np.uint8)
    fn = load().cfx_dds_header if _enum(CONTAINERS, container) == 0 else load().cfx_ktx_header
    n = int(fn(ctypes.byref(d), int(mip_levels), int(array_size), buf.ctypes.data))
    return buf[:n].copy() if n else None


def encode_mip_chain_to_file(img, fmt, path, container=None, filter="CatmullRom", levels=None, **kw):
    """generateMipmaps(filter, levels) + convert() + save(path): one call, the blocks land in a mapping of the file."""
    img = np.asarray(img)
    if img.dtype != np.uint8:
        img = img.astype(np.float32, copy=False)
    img = np.ascontiguousarray(img)
    h, w, _ = img.shape
    src_format, texel = _src_format_of(img.dtype)
    if container is None:
        container = "KTX" if str(path).lower().endswith(".ktx") else "DDS"
    d = make_desc(fmt, w, h, src_format, w * texel, **kw)
    _check(load().cfx_encode_mip_chain_to_file(ctypes.byref(d), img.ctypes.data, _enum(FILTERS, filter), int(levels or 0),
                                              _enum(CONTAINERS, container), os.fsencode(path)))


class Texture:
    """The slice of cuttlefish::Texture the convert path uses (lib/include/cuttlefish/Texture.h).

    Images are float32/uint8 [H,W,4] arrays per (mip, depth); convert() replaces them by packed
    block bytes exactly like Texture::convert -> Converter::convert (lib/src/Texture.cpp:1536-1561,
    lib/src/Converter.cpp:508-593): returns False and leaves the texture unconverted when the
    (format, type) pair has no encoder."""

    def __init__(self, width, height, depth=1, mip_levels=1, srgb=False):
        self.width, self.height, self.depth, self.mip_levels = int(width), int(height), int(depth), int(mip_levels)
        self.srgb = bool(srgb)
        self._images = [[None] * self.depth for _ in range(self.mip_levels)]
        self._data = None
        self.format = None
        self.type = None

    def mip_size(self, mip):
        return max(1, self.width >> mip), max(1, self.height >> mip)

    def setImage(self, image, mip=0, depth=0):
        image = np.asarray(image)
        w, h = self.mip_size(mip)
        if image.shape != (h, w, 4):
            return False
        self._images[mip][depth] = image
        return True

    def generateMipmaps(self, filter="CatmullRom", mipLevels=None):
        """Texture::generateMipmaps (lib/src/Texture.cpp:1320-1514) for a 2D texture without custom mips: every level is
        resized on the GPU from the level above. Level 0 must be set; returns False otherwise."""
        if self._images is None or any(im is None for im in self._images[0]):
            return False
        n = mip_levels(self.width, self.height)
        n = n if mipLevels is None else max(1, min(int(mipLevels), n))
        images = [list(self._images[0])] + [[None] * self.depth for _ in range(n - 1)]
        for d in range(self.depth):
            for mip in range(1, n):
                w, h = self.mip_size(mip)
                above = images[mip - 1][d]
                if above.dtype == np.uint8:         # Image::convert(RGBAF) of an 8-bit image
                    above = above.astype(np.float32) / np.float32(255.0)
                images[mip][d] = resize(above, w, h, filter, self.srgb)
        self._images, self.mip_levels = images, n
        return True

    def imagesComplete(self):
        return all(im is not None for level in self._images for im in level)

    def converted(self):
        return self._data is not None

    def convert(self, format, type="UNorm", quality="Normal", alphaType="Standard", colorMask=None, threads=None):
        if not self.imagesComplete() or self.converted():
            return False
        if not format_supported(format, type):
            return False
        # an sRGB texture converts only to formats with an sRGB variant (Texture::hasNativeSRGB, lib/src/Texture.cpp:421-468,
        # checked at :1542): a container header exists exactly for those
        if self.srgb and container_header("KTX", format, 4, 4, type=type, srgb=True) is None:
            return False
        mask = colorMask if colorMask is not None else ColorMask()
        # one batch for every surface of the texture (the mip/depth loop of Converter::convert)
        flat = [image for level in self._images for image in level]
        outs = encode_batch(flat, format, type=type, quality=quality, alpha=alphaType, color_mask=mask, srgb=self.srgb)
        it = iter(outs)
        self._data = [[next(it) for _ in level] for level in self._images]
        self._images = None
        self.format, self.type = format, type
        return True

    def dataSize(self, mip=0, depth=0):
        return 0 if self._data is None else int(self._data[mip][depth].size)

    def data(self, mip=0, depth=0):
        return None if self._data is None else self._data[mip][depth]
