"""TEST INFRASTRUCTURE ONLY: ctypes wrapper over oracle/_ref/libcfref.so + input generator G.

encode()/decode() run the reference's CPU encoders/decoders (see oracle/cfref.cpp for the
file:line map).  gen_image() is generator G of SURVEY.md section 8(d): the synthetic inputs
every BASELINE.md number was measured on.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libcfref.so")

# cuttlefish::Texture::Format values (lib/include/cuttlefish/Texture.h:59-130)
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


class Desc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in
                ("format", "type", "quality", "alpha_type", "color_mask", "color_space", "width", "height")]


_lib = None


def build(quiet=True):
    """Build oracle/_ref/libcfref.so from /root/reference (only possible where it is mounted)."""
    ref = os.environ.get("CFX_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "lib", "bc7enc_rdo")):
        return os.path.exists(_LIB_PATH)
    r = subprocess.run(["make", "-C", _HERE, "-j8", "REF=" + ref], capture_output=quiet, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return True


def available():
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.cfref_encoded_size.restype = ctypes.c_size_t
        L.cfref_encoded_size.argtypes = [ctypes.POINTER(Desc)]
        L.cfref_encode.restype = ctypes.c_int
        L.cfref_encode.argtypes = [ctypes.POINTER(Desc), ctypes.c_void_p, ctypes.c_size_t,
                                   ctypes.c_void_p, ctypes.c_uint]
        L.cfref_decode.restype = ctypes.c_int
        L.cfref_decode.argtypes = [ctypes.POINTER(Desc), ctypes.c_void_p, ctypes.c_void_p]
        L.cfref_hardware_threads.restype = ctypes.c_uint
        _lib = L
    return _lib


def make_desc(fmt, width, height, type="UNorm", quality="Normal", alpha="Standard", color_mask=15,
              srgb=False):
    g = lambda table, v: table[v] if isinstance(v, str) else int(v)
    return Desc(g(FORMATS, fmt), g(TYPES, type), g(QUALITY, quality), g(ALPHA, alpha),
                int(color_mask), 1 if srgb else 0, int(width), int(height))


def encode(img, fmt, threads=0, **kw):
    """img: float32 [H,W,4] RGBAF (row 0 = top).  Returns uint8 packed blocks (row-major blocks)."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    h, w, c = img.shape
    assert c == 4
    d = make_desc(fmt, w, h, **kw)
    n = lib().cfref_encoded_size(ctypes.byref(d))
    if n == 0:
        raise ValueError("unsupported format %r" % (fmt,))
    out = np.empty(n, dtype=np.uint8)
    rc = lib().cfref_encode(ctypes.byref(d), img.ctypes.data, w * 4, out.ctypes.data, threads)
    if rc != 0:
        raise RuntimeError("cfref_encode failed: %d" % rc)
    return out


_GLUE_PATH = os.path.join(_HERE, "_ref", "libcfglue.so")
_glue = None


def glue_available():
    return os.path.exists(_GLUE_PATH)


def encode_glue(img, fmt, threads=0, out_bytes=None, **kw):
    """Same as encode(), but through the reference's REAL converter glue (lib/src/Converter.cpp and the
    *Converter.cpp files compiled from /root/reference over oracle/glue_stub.cpp) instead of our
    restatement in cfref.cpp.  Used only to pin the restatement."""
    global _glue
    if _glue is None:
        _glue = ctypes.CDLL(_GLUE_PATH)
        _glue.cfglue_encode.restype = ctypes.c_int
        _glue.cfglue_encode.argtypes = [ctypes.POINTER(Desc), ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                        ctypes.c_size_t, ctypes.c_uint]
    img = np.ascontiguousarray(img, dtype=np.float32)
    h, w, _ = img.shape
    d = make_desc(fmt, w, h, **kw)
    n = out_bytes if out_bytes is not None else lib().cfref_encoded_size(ctypes.byref(d))   # out_bytes: non-block formats
    out = np.empty(n, dtype=np.uint8)
    rc = _glue.cfglue_encode(ctypes.byref(d), img.ctypes.data, w * 4, out.ctypes.data, n, threads)
    if rc != n:
        raise RuntimeError("cfglue_encode failed: %d" % rc)
    return out


_GLUE_CUDA_PATH = os.path.join(_HERE, "_ref", "libcfglue_cuda.so")
_glue_cuda = None


def glue_cuda_available():
    return os.path.exists(_GLUE_CUDA_PATH)


def encode_glue_cuda(img, fmt, threads=0, bottom_up=False, out_bytes=None, **kw):
    """The reference's REAL Converter::convert() with adapter/CudaConverter.cpp hooked into its createConverter()
    (oracle/_ref/libcfglue_cuda.so): supported pairs are encoded by libcfx.so through the drop-in boundary, the rest by
    the reference's CPU converters. bottom_up makes the stand-in Image store its rows the way cuttlefish::Image does.
    Returns (blocks, seconds spent in Converter::convert, surfaces the adapter sent to the GPU during this call)."""
    global _glue_cuda
    if _glue_cuda is None:
        _glue_cuda = ctypes.CDLL(_GLUE_CUDA_PATH)
        _glue_cuda.cfglue_encode_timed.restype = ctypes.c_int
        _glue_cuda.cfglue_encode_timed.argtypes = [ctypes.POINTER(Desc), ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                                   ctypes.c_size_t, ctypes.c_uint, ctypes.POINTER(ctypes.c_double)]
        _glue_cuda.cfglue_set_bottom_up.argtypes = [ctypes.c_int]
        _glue_cuda.cfx_adapter_surfaces_encoded.restype = ctypes.c_uint
    img = np.ascontiguousarray(img, dtype=np.float32)
    h, w, _ = img.shape
    d = make_desc(fmt, w, h, **kw)
    n = out_bytes if out_bytes is not None else lib().cfref_encoded_size(ctypes.byref(d))
    out = np.empty(n, dtype=np.uint8)
    secs = ctypes.c_double(0.0)
    before = _glue_cuda.cfx_adapter_surfaces_encoded()
    _glue_cuda.cfglue_set_bottom_up(1 if bottom_up else 0)
    try:
        rc = _glue_cuda.cfglue_encode_timed(ctypes.byref(d), img.ctypes.data, w * 4, out.ctypes.data, n, threads,
                                            ctypes.byref(secs))
    finally:
        _glue_cuda.cfglue_set_bottom_up(0)
    if rc != n:
        raise RuntimeError("cfglue_encode failed: %d" % rc)
    return out, secs.value, _glue_cuda.cfx_adapter_surfaces_encoded() - before


def decode(blocks, fmt, width, height, **kw):
    """Decode packed blocks with the reference's decoders -> float32 [H,W,4]."""
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
    d = make_desc(fmt, width, height, **kw)
    assert blocks.size == lib().cfref_encoded_size(ctypes.byref(d))
    out = np.zeros((height, width, 4), dtype=np.float32)
    rc = lib().cfref_decode(ctypes.byref(d), blocks.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("cfref_decode failed: %d" % rc)
    return out


def hardware_threads():
    return int(lib().cfref_hardware_threads())


# ---- generator G lives with the product's synthetic-input helpers; it is not
# part of the oracle algorithm, the tests just reach it through this module.
import sys as _sys
_sys.path.insert(0, os.path.dirname(_HERE))
from cuttlefish_b200.synth import gen_image, psnr_rgb, to_rgba8  # noqa: E402,F401


def fnv1a64(data):
    h = 0xcbf29ce484222325
    for b in bytes(data):
        h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h
