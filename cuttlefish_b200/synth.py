"""Synthetic inputs (generator G of SURVEY.md section 8d) used by the tests and bench.py.

    gradient:   r = x/(w-1), g = y/(h-1), b = (w-1-x)/(w-1)           (the ramp of the reference's
                test helper getTestColor, lib/test/TextureTest.cpp:53-61, minus alpha)
    noise+grad: 0.75*gradient + 0.25*n, n = (s>>24)/255 from the LCG s = s*1664525 + 1013904223
                (uint32, three draws per texel in r,g,b order, row-major)
    both LDR kinds snapped to 8 bit: v = round(v*255)/255
    hdr:        t = (x + y*w)/(w*h); r = 64t, g = 8x/(w-1), b = 0.5y/(h-1)   (not snapped)

gen_image(..., rows=(y0, y1)) produces only a slab of rows of the same image (LCG jump-ahead),
which is what a rank of a block-row-sharded run needs.
"""
import numpy as np

_A, _C, _MASK = 1664525, 1013904223, 0xFFFFFFFF


def _lcg_skip(seed, n):
    """State after n steps of the LCG from `seed` (affine map exponentiation)."""
    a, c = _A, _C
    ra, rc = 1, 0
    while n:
        if n & 1:
            ra, rc = (ra * a) & _MASK, (rc * a + c) & _MASK
        a, c = (a * a) & _MASK, (c * a + c) & _MASK
        n >>= 1
    return (ra * seed + rc) & _MASK


def _lcg_noise(n_values, seed=12345, skip=0):
    """n_values draws (after skipping `skip` draws) as float32 (s>>24)/255."""
    a, c, mask = np.uint64(_A), np.uint64(_C), np.uint64(_MASK)
    blk = 1 << 16
    # affine powers: s_k = A_k s_0 + C_k for k = 1..blk
    A = np.empty(blk, dtype=np.uint64)
    C = np.empty(blk, dtype=np.uint64)
    ak, ck = 1, 0
    for i in range(blk):
        ak = (ak * _A) & _MASK
        ck = (ck * _A + _C) & _MASK
        A[i], C[i] = ak, ck
    del a, c
    out = np.empty(n_values, dtype=np.uint32)
    s = np.uint64(_lcg_skip(seed, skip))
    pos = 0
    while pos < n_values:
        m = min(blk, n_values - pos)
        vals = (A[:m] * s + C[:m]) & mask
        out[pos:pos + m] = vals.astype(np.uint32)
        s = vals[m - 1]
        pos += m
    return (out >> np.uint32(24)).astype(np.float32) / np.float32(255.0)


def ui_image(n, seed=5):
    """Not part of generator G: a screenshot-like quality probe (gray gradient background, dark text-like strokes,
    soft-edged coloured discs, one smooth colourful quadrant) on which encoders that lack luminance end points or
    good partition choices fall behind the reference.  float32 [n,n,4], 8-bit snapped, alpha 1."""
    rng = np.random.default_rng(seed)
    img = np.zeros((n, n, 4), np.float32)
    img[..., 3] = 1
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float32)
    img[..., :3] = (0.85 - 0.25 * yy / n)[..., None]
    for _ in range(max(1, n * n // 520)):
        x0, y0 = rng.integers(0, n - 20), rng.integers(0, n - 4)
        w_, h_ = rng.integers(3, 18), rng.integers(1, 3)
        img[y0:y0 + h_, x0:x0 + w_, :3] = 0.15 + 0.1 * rng.random()
    for _ in range(max(1, n * n // 3500)):
        cx, cy, r = rng.integers(10, n - 10), rng.integers(10, n - 10), rng.integers(5, 14)
        a = np.clip((r - np.sqrt((xx - cx) ** 2 + (yy - cy) ** 2)) / 1.5, 0, 1)[..., None]
        img[..., :3] = img[..., :3] * (1 - a) + rng.random(3).astype(np.float32) * a
    q = n // 2
    f = np.stack([0.5 + 0.5 * np.sin(xx[:q, :q] / 9.0), 0.5 + 0.5 * np.sin(yy[:q, :q] / 13.0 + 1),
                  0.5 + 0.5 * np.sin((xx[:q, :q] + yy[:q, :q]) / 17.0)], -1)
    img[q:, q:, :3] = f + 0.03 * rng.standard_normal((q, q, 3)).astype(np.float32)
    img = np.clip(img, 0, 1)
    return (np.floor(img * 255 + 0.5) / 255).astype(np.float32)


def gen_image(kind, width, height, seed=12345, rows=None):
    """Generator G. Returns float32 [H,W,4] (or rows y0..y1 of it), row 0 = top, alpha 1."""
    if kind == "ui":
        assert width == height and rows is None
        return ui_image(int(width))
    w, h = int(width), int(height)
    y0, y1 = (0, h) if rows is None else (int(rows[0]), int(rows[1]))
    n = y1 - y0
    x = np.arange(w, dtype=np.float32)[None, :]
    y = np.arange(y0, y1, dtype=np.float32)[:, None]
    dx = np.float32(max(w - 1, 1))
    dy = np.float32(max(h - 1, 1))
    img = np.empty((n, w, 4), dtype=np.float32)
    img[..., 3] = 1.0
    if kind == "hdr":
        t = (x + y * np.float32(w)) / np.float32(w * h)
        img[..., 0] = np.float32(64.0) * t
        img[..., 1] = np.broadcast_to(np.float32(8.0) * x / dx, (n, w))
        img[..., 2] = np.broadcast_to(np.float32(0.5) * y / dy, (n, w))
        return img
    gr = np.broadcast_to(x / dx, (n, w))
    gg = np.broadcast_to(y / dy, (n, w))
    gb = np.broadcast_to((np.float32(w - 1) - x) / dx, (n, w))
    if kind == "gradient":
        img[..., 0], img[..., 1], img[..., 2] = gr, gg, gb
    elif kind in ("noise+grad", "noise"):
        nz = _lcg_noise(w * n * 3, seed, skip=y0 * w * 3).reshape(n, w, 3)
        img[..., 0] = np.float32(0.75) * gr + np.float32(0.25) * nz[..., 0]
        img[..., 1] = np.float32(0.75) * gg + np.float32(0.25) * nz[..., 1]
        img[..., 2] = np.float32(0.75) * gb + np.float32(0.25) * nz[..., 2]
    else:
        raise ValueError(kind)
    img[..., :3] = np.floor(img[..., :3] * np.float32(255.0) + np.float32(0.5)) / np.float32(255.0)
    return img


def to_rgba8(img):
    """round(clamp01(v)*255) -> uint8 (lib/src/S3tcConverter.cpp:97-111; half away from zero)."""
    v = np.clip(img, 0.0, 1.0).astype(np.float32) * np.float32(255.0)
    return np.floor(v + np.float32(0.5)).astype(np.uint8)


def psnr_rgb(a, b, peak=1.0):
    d = (a[..., :3].astype(np.float64) - b[..., :3].astype(np.float64))
    mse = float(np.mean(d * d))
    return float("inf") if mse == 0 else 10.0 * np.log10(peak * peak / mse)
