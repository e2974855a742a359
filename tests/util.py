import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases(prefixes=None):
    out = []
    for path in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        name = os.path.basename(path)[:-4]
        if prefixes is None or any(name.startswith(p + "_") for p in prefixes):
            out.append(name)
    return out


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    kw = eval(str(z["kw"]))  # written by make_goldens.py: a dict literal
    src = z["src"]
    if src.dtype == np.uint16:
        src = src.view(np.float16)
    return src, z["blocks"], str(z["format"]), kw


def src_as_float(src):
    if src.dtype == np.uint8:
        return src.astype(np.float32) / np.float32(255.0)
    return src.astype(np.float32)


def block_mismatches(a, b, block_bytes):
    a = np.asarray(a).reshape(-1, block_bytes)
    b = np.asarray(b).reshape(-1, block_bytes)
    return np.nonzero((a != b).any(axis=1))[0]


def decode_bc4_snorm(blocks, width, height):
    """BC4_SNORM per the D3D spec: int8 end points (-128 reads as -127), 8 interpolated values when
    red0 > red1 else 6 + {-1, +1}; returns float32 [H, W] in [-1, 1]."""
    bx, by = (width + 3) // 4, (height + 3) // 4
    b = np.asarray(blocks, np.uint8).reshape(by, bx, 8)
    e0 = np.maximum(b[..., 0].astype(np.int8).astype(np.float32), -127.0) / 127.0
    e1 = np.maximum(b[..., 1].astype(np.int8).astype(np.float32), -127.0) / 127.0
    pal = np.zeros((by, bx, 8), np.float32)
    pal[..., 0], pal[..., 1] = e0, e1
    gt = b[..., 0].astype(np.int8) > b[..., 1].astype(np.int8)
    for i in range(1, 7):
        pal[..., i + 1] = np.where(gt, (e0 * (7 - i) + e1 * i) / 7.0, 0)
    for i in range(1, 5):
        pal[..., i + 1] = np.where(gt, pal[..., i + 1], (e0 * (5 - i) + e1 * i) / 5.0)
    pal[..., 6] = np.where(gt, pal[..., 6], -1.0)
    pal[..., 7] = np.where(gt, pal[..., 7], 1.0)
    bits = np.zeros((by, bx), np.uint64)
    for k in range(6):
        bits |= b[..., 2 + k].astype(np.uint64) << np.uint64(8 * k)
    out = np.zeros((by * 4, bx * 4), np.float32)
    for t in range(16):
        sel = ((bits >> np.uint64(3 * t)) & np.uint64(7)).astype(np.int64)
        out[(t // 4)::4, (t % 4)::4] = np.take_along_axis(pal, sel[..., None], axis=2)[..., 0]
    return out[:height, :width]


_EAC_TABLE = np.array([
    [-3, -6, -9, -15, 2, 5, 8, 14], [-3, -7, -10, -13, 2, 6, 9, 12], [-2, -5, -8, -13, 1, 4, 7, 12],
    [-2, -4, -6, -13, 1, 3, 5, 12], [-3, -6, -8, -12, 2, 5, 7, 11], [-3, -7, -9, -11, 2, 6, 8, 10],
    [-4, -7, -8, -11, 3, 6, 7, 10], [-3, -5, -8, -11, 2, 4, 7, 10], [-2, -6, -8, -10, 1, 5, 7, 9],
    [-2, -5, -8, -10, 1, 4, 7, 9], [-2, -4, -8, -10, 1, 3, 7, 9], [-2, -5, -7, -10, 1, 4, 6, 9],
    [-3, -4, -7, -10, 2, 3, 6, 9], [-1, -2, -3, -10, 0, 1, 2, 9], [-4, -6, -8, -9, 3, 5, 7, 8],
    [-3, -5, -7, -9, 2, 4, 6, 8]], np.int64)     # ETC2 / EAC modifier table (Khronos data format spec, table "EAC modifier")


def decode_eac_r11(blocks, width, height, signed):
    """EAC R11 per the Khronos spec: unsigned clamp(base*8 + 4 + mod*mul*8, 0, 2047)/2047 (mul 0 -> mod*1),
    signed clamp(base*8 + mod*mul*8, -1023, 1023)/1023 with an int8 base (-128 reads as -127).
    blocks: [n, 8] uint8; returns float32 [H, W]."""
    bx, by = (width + 3) // 4, (height + 3) // 4
    b = np.asarray(blocks, np.uint8).reshape(by, bx, 8).astype(np.int64)
    base = b[..., 0]
    if signed:
        base = np.where(base > 127, base - 256, base)
        base = np.maximum(base, -127)
    mul, tab = b[..., 1] >> 4, b[..., 1] & 15
    bits = np.zeros((by, bx), np.int64)
    for k in range(6):
        bits = (bits << 8) | b[..., 2 + k]
    out = np.zeros((by * 4, bx * 4), np.float32)
    for p in range(16):                      # pixel p = x*4 + y, most significant selector first
        sel = (bits >> (45 - 3 * p)) & 7
        mod = _EAC_TABLE[tab, sel]
        step = np.where(mul > 0, mod * mul * 8, mod)
        if signed:
            v = np.clip(base * 8 + step, -1023, 1023) / 1023.0
        else:
            v = np.clip(base * 8 + 4 + step, 0, 2047) / 2047.0
        out[(p % 4)::4, (p // 4)::4] = v
    return out[:height, :width]


def decode_any(oracle, blocks, fmt, width, height, kw):
    """Decoded float32 [H, W, C] for any golden case: the reference's decoders where the oracle has one, the spec
    decoders above for signed BC4/BC5 and EAC R11/RG11 (the reference's own decode path asserts on those)."""
    typ = kw.get("type", "UNorm")
    blocks = np.asarray(blocks, np.uint8)
    if fmt in ("EAC_R11", "EAC_R11G11") or (fmt in ("BC4", "BC5") and typ == "SNorm"):
        nch = 1 if fmt in ("EAC_R11", "BC4") else 2
        b = blocks.reshape(-1, 8*nch)
        out = np.zeros((height, width, 4), np.float32)
        out[..., 3] = 1.0
        for c in range(nch):
            part = np.ascontiguousarray(b[:, 8*c:8*c + 8])
            out[..., c] = decode_eac_r11(part, width, height, typ == "SNorm") if fmt.startswith("EAC") else decode_bc4_snorm(part, width, height)
        return out
    return oracle.decode(blocks, fmt, width, height, **kw)
