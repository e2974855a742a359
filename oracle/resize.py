"""TEST INFRASTRUCTURE ONLY: the oracle for Image::resize() / Texture::generateMipmaps().

Two things live here:
  * resize_ref(): the reference itself -- FreeImage_Rescale compiled from /root/reference into
    oracle/_ref/libfiresize.so by oracle/Makefile (see oracle/fi_resize.cpp);
  * resize_np(): a numpy restatement of the same algorithm (float64 weights and accumulation in window order,
    each pass rounded to float32), which travels without the compiled library. tests/test_resize_cpu.py pins it
    bit-for-bit to resize_ref() and to the committed vectors in tests/golden/resize/cases.npz.
File:line references are to /root/reference/lib.
"""
import ctypes
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_FI_PATH = os.path.join(_HERE, "_ref", "libfiresize.so")

# cuttlefish::Image::ResizeFilter, include/cuttlefish/Image.h:79-86
FILTERS = {"Box": 0, "Linear": 1, "Cubic": 2, "CatmullRom": 3, "BSpline": 4}

_fi = None


def ref_available():
    return os.path.exists(_FI_PATH)


def resize_ref(img, dw, dh, filter="CatmullRom", srgb=False):
    """Image::resize() through the real FreeImage_Rescale. img: (h, w, 4) float32, row 0 = top."""
    global _fi
    if _fi is None:
        _fi = ctypes.CDLL(_FI_PATH)
        _fi.cfref_resize_rgbaf.restype = ctypes.c_int
        _fi.cfref_resize_rgbaf.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int]
    img = np.ascontiguousarray(img, np.float32)
    out = np.empty((dh, dw, 4), np.float32)
    rc = _fi.cfref_resize_rgbaf(img.ctypes.data, img.shape[1], img.shape[0], out.ctypes.data, dw, dh,
                                FILTERS[filter] if isinstance(filter, str) else int(filter), int(bool(srgb)))
    if rc != 0:
        raise RuntimeError("cfref_resize_rgbaf failed: %d" % rc)
    return out


def _filter_eval(f, v):
    """FreeImage/Source/FreeImageToolkit/Filters.h:69-257."""
    if f == 0:
        return 1.0 if abs(v) <= 0.5 else 0.0
    if f == 1:
        v = abs(v)
        return 1.0 - v if v < 1.0 else 0.0
    if f == 2:
        b = c = 1 / 3.0
        p0, p2, p3 = (6 - 2 * b) / 6, (-18 + 12 * b + 6 * c) / 6, (12 - 9 * b - 6 * c) / 6
        q0, q1, q2, q3 = (8 * b + 24 * c) / 6, (-12 * b - 48 * c) / 6, (6 * b + 30 * c) / 6, (-b - 6 * c) / 6
        v = abs(v)
        if v < 1:
            return p0 + v * v * (p2 + v * p3)
        if v < 2:
            return q0 + v * (q1 + v * (q2 + v * q3))
        return 0.0
    if f == 3:
        if v < -2:
            return 0.0
        if v < -1:
            return 0.5 * (4 + v * (8 + v * (5 + v)))
        if v < 0:
            return 0.5 * (2 + v * v * (-5 - 3 * v))
        if v < 1:
            return 0.5 * (2 + v * v * (-5 + 3 * v))
        if v < 2:
            return 0.5 * (4 + v * (-8 + v * (5 - v)))
        return 0.0
    v = abs(v)
    if v < 1:
        return (4 + v * v * (-6 + 3 * v)) / 6
    if v < 2:
        t = 2 - v
        return t * t * t / 6
    return 0.0


def windows(f, dst, src):
    """CWeightsTable, Resize.cpp:140-218 -> (left[dst], count[dst], weight[dst][window])."""
    support = 0.5 if f == 0 else (1.0 if f == 1 else 2.0)
    scale = float(dst) / float(src)
    if scale < 1.0:
        width, fscale = support / scale, scale
    else:
        width, fscale = support, 1.0
    window = 2 * int(math.ceil(width)) + 1
    left = np.zeros(dst, np.int64)
    count = np.zeros(dst, np.int64)
    weight = np.zeros((dst, window), np.float64)
    offset = 0.5 / scale
    for u in range(dst):
        center = u / scale + offset
        lo = max(0, int(center - width + 0.5))
        hi = min(int(center + width + 0.5), src)
        w = [fscale * _filter_eval(f, fscale * (i + 0.5 - center)) for i in range(lo, hi)]
        total = 0.0
        for x in w:
            total += x
        if total > 0 and total != 1:
            w = [x / total for x in w]
        while hi - lo >= 1 and w[hi - lo - 1] == 0:
            hi -= 1
            if hi == lo:
                break
        left[u], count[u] = lo, hi - lo
        weight[u, :hi - lo] = w[:hi - lo]
    return left, count, weight


def _pass(img, f, dst, axis):
    """One filter pass along `axis` (0 = rows of a bottom-up image, 1 = x): Resize.cpp:1236-1273 / :2070-2112."""
    src = img.shape[axis]
    left, count, weight = windows(f, dst, src)
    a = np.moveaxis(img, axis, 0).astype(np.float64)          # [src, other, 4]
    acc = np.zeros((dst,) + a.shape[1:], np.float64)
    for k in range(weight.shape[1]):
        live = count > k                                         # taps past a window's end do not exist
        idx = np.minimum(left + k, src - 1)
        term = weight[:, k].reshape(-1, 1, 1) * a[idx]
        acc[live] = acc[live] + term[live]
    return np.moveaxis(acc.astype(np.float32), 0, axis)


def _srgb_to_linear(c):
    c = c.astype(np.float64)
    return np.where(c <= 0.04045, c / 12.92, np.power((np.maximum(c, 0.04045) + 0.055) / 1.055, 2.4))


def _linear_to_srgb(c):
    c = c.astype(np.float64)
    return np.where(c <= 0.0031308, c * 12.92, 1.055 * np.power(np.maximum(c, 0.0031308), 1.0 / 2.4) - 0.055)


def resize_np(img, dw, dh, filter="CatmullRom", srgb=False):
    """Restatement of Image::resize() (lib/src/Image.cpp:1324-1379) for an RGBAF image, row 0 = top."""
    f = FILTERS[filter] if isinstance(filter, str) else int(filter)
    img = np.ascontiguousarray(img, np.float32)
    sh, sw = img.shape[:2]
    if (sw, sh) == (dw, dh):
        return img.copy()
    if srgb:                                                     # Image.cpp:1337-1344, Color.h:224-242
        img = img.copy()
        img[..., :3] = _srgb_to_linear(img[..., :3]).astype(np.float32)
    a = img[::-1]                                                # FreeImage bitmaps are bottom-up
    if dw <= sw:                                                 # Resize.cpp:371: x first unless the width grows
        if sw != dw:
            a = _pass(a, f, dw, 1)
        if sh != dh:
            a = _pass(a, f, dh, 0)
    else:
        if sh != dh:
            a = _pass(a, f, dh, 0)
        if sw != dw:
            a = _pass(a, f, dw, 1)
    out = np.ascontiguousarray(a[::-1])
    if srgb:
        out[..., :3] = _linear_to_srgb(out[..., :3]).astype(np.float32)
    return out


def mip_sizes(w, h, levels=None):
    """Texture::maxMipmapLevels / Texture::width(mip), lib/src/Texture.cpp:514-527."""
    n = max(w, h).bit_length()
    n = n if levels is None else max(1, min(levels, n))
    return [(max(1, w >> k), max(1, h >> k)) for k in range(n)]


def mip_chain(img, filter="CatmullRom", levels=None, srgb=False, fn=None):
    """Texture::generateMipmaps() for one 2D surface: every level resized from the one above
    (lib/src/Texture.cpp:1457-1511)."""
    fn = fn or resize_np
    out = [np.ascontiguousarray(img, np.float32)]
    for (w, h) in mip_sizes(img.shape[1], img.shape[0], levels)[1:]:
        out.append(fn(out[-1], w, h, filter, srgb))
    return out
