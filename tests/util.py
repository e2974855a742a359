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


# ---- BC6H decoder for the block modes OUR encoder emits (1, 2, 10, 11-14), per the D3D11 BC6H specification.
# Used to check the signed format (Texture::Type::Float): the reference's own signed pipeline (Compressonator) does
# not survive a round trip through its own decoder, so there is no oracle decode to lean on.  The decoder is pinned on
# UNSIGNED blocks against the reference decoder (tests/test_oracle_cpu.py), signed differs only in the sign extension
# of the end points, the signed unquantiser and the signed half conversion the spec mandates.
_BC6_PART2 = [0xcccc, 0x8888, 0xeeee, 0xecc8, 0xc880, 0xfeec, 0xfec8, 0xec80, 0xc800, 0xffec, 0xfe80, 0xe800, 0xffe8,
              0xff00, 0xfff0, 0xf000, 0xf710, 0x008e, 0x7100, 0x08ce, 0x008c, 0x7310, 0x3100, 0x8cce, 0x088c, 0x3110,
              0x6666, 0x366c, 0x17e8, 0x0ff0, 0x718e, 0x399c]
_BC6_ANCHOR2 = [15]*16 + [15, 2, 8, 2, 2, 8, 8, 15, 2, 8, 2, 2, 8, 8, 2, 2]
_BC6_W3 = [0, 9, 18, 27, 37, 46, 55, 64]
_BC6_W4 = [0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64]


def _bc6_unq(x, bits, signed):
    if not signed:
        if bits >= 15:
            return x
        if x == 0:
            return 0
        if x == (1 << bits) - 1:
            return 0xFFFF
        return ((x << 15) + 0x4000) >> (bits - 1)
    if bits >= 16:
        return x
    neg, ax = x < 0, abs(x)
    if ax == 0:
        u = 0
    elif ax >= (1 << (bits - 1)) - 1:
        u = 0x7FFF
    else:
        u = ((ax << 15) + 0x4000) >> (bits - 1)
    return -u if neg else u


def _sext(v, bits):
    v &= (1 << bits) - 1
    return v - (1 << bits) if v & (1 << (bits - 1)) else v


def _bc6_finish(p, signed):
    if not signed:
        return (p * 31) >> 6
    return (0x8000 | (((-p) * 31) >> 5)) if p < 0 else ((p * 31) >> 5)


def decode_bc6h_block(block, signed):
    """16 bytes -> [16, 3] uint16 half bit patterns (texel t = y*4 + x); None for a mode this decoder does not cover."""
    v = int.from_bytes(bytes(block), "little")
    bits = lambda pos, n: (v >> pos) & ((1 << n) - 1)
    bit = lambda pos: (v >> pos) & 1
    m2 = v & 3
    m5 = v & 31
    out = np.zeros((16, 3), np.uint16)
    if m2 >= 2 and m5 in (0x03, 0x07, 0x0B, 0x0F):              # one region: modes 11..14
        k = {0x03: 0, 0x07: 1, 0x0B: 2, 0x0F: 3}[m5]
        wb, tb = [10, 11, 12, 16][k], [10, 9, 8, 4][k]
        e0, e1 = [], []
        for c in range(3):
            w = bits(5 + 10 * c, 10)
            for b in range(10, wb):
                w |= bit(35 + 10 * c + (9 - (b - 10))) << b
            x = bits(35 + 10 * c, tb)
            if k == 0:
                a, b_ = w, x
            else:
                a, b_ = w, (w + _sext(x, tb)) & ((1 << wb) - 1)
            if signed:
                a, b_ = _sext(a, wb), _sext(b_, wb)
            e0.append(_bc6_unq(a, wb, signed)); e1.append(_bc6_unq(b_, wb, signed))
        for t in range(16):
            idx = bits(65, 3) if t == 0 else bits(65 + 3 + 4 * (t - 1), 4)
            wgt = _BC6_W4[idx]
            for c in range(3):
                out[t, c] = _bc6_finish((e0[c] * (64 - wgt) + e1[c] * wgt + 32) >> 6, signed)
        return out
    if m2 == 0 or m2 == 1 or m5 in (0x1E, 0x0E):                # two regions: modes 1 (10.5.5.5), 2 (7.6.6.6), 6 (9.5.5.5), 10 (6.6.6.6)
        if m5 == 0x1E:
            wb, tb, direct = 6, 6, True
            w = [bits(5, 6), bits(15, 6), bits(25, 6)]; x = [bits(35, 6), bits(45, 6), bits(55, 6)]
            y = [bits(65, 6), bits(41, 4) | bit(24) << 4 | bit(21) << 5, bits(61, 4) | bit(14) << 4 | bit(22) << 5]
            z = [bits(71, 6), bits(51, 4) | bit(11) << 4 | bit(31) << 5,
                 bit(12) | bit(13) << 1 | bit(23) << 2 | bit(32) << 3 | bit(34) << 4 | bit(33) << 5]
        elif m5 == 0x0E:
            wb, tb, direct = 9, 5, False
            w = [bits(5, 9), bits(15, 9), bits(25, 9)]; x = [bits(35, 5), bits(45, 5), bits(55, 5)]
            y = [bits(65, 5), bits(41, 4) | bit(24) << 4, bits(61, 4) | bit(14) << 4]
            z = [bits(71, 5), bits(51, 4) | bit(40) << 4, bit(50) | bit(60) << 1 | bit(70) << 2 | bit(76) << 3 | bit(34) << 4]
        elif m2 == 1:
            wb, tb, direct = 7, 6, False
            w = [bits(5, 7), bits(15, 7), bits(25, 7)]; x = [bits(35, 6), bits(45, 6), bits(55, 6)]
            y = [bits(65, 6), bits(41, 4) | bit(24) << 4 | bit(2) << 5, bits(61, 4) | bit(14) << 4 | bit(22) << 5]
            z = [bits(71, 6), bits(51, 4) | bit(3) << 4 | bit(4) << 5,
                 bit(12) | bit(13) << 1 | bit(23) << 2 | bit(32) << 3 | bit(34) << 4 | bit(33) << 5]
        else:
            wb, tb, direct = 10, 5, False
            w = [bits(5, 10), bits(15, 10), bits(25, 10)]; x = [bits(35, 5), bits(45, 5), bits(55, 5)]
            y = [bits(65, 5), bits(41, 4) | bit(2) << 4, bits(61, 4) | bit(3) << 4]
            z = [bits(71, 5), bits(51, 4) | bit(40) << 4, bit(50) | bit(60) << 1 | bit(70) << 2 | bit(76) << 3 | bit(4) << 4]
        shape = bits(77, 5)
        ep = []
        for c in range(3):
            vals = [w[c]]
            for d in (x[c], y[c], z[c]):
                vals.append(d if direct else (w[c] + _sext(d, tb)) & ((1 << wb) - 1))
            if signed:
                vals = [_sext(q, wb) for q in vals]
            ep.append([_bc6_unq(q, wb, signed) for q in vals])
        pos = 82
        mask, anchor = _BC6_PART2[shape], _BC6_ANCHOR2[shape]
        for t in range(16):
            nb = 2 if (t == 0 or t == anchor) else 3
            idx = bits(pos, nb); pos += nb
            wgt = _BC6_W3[idx]
            r = (mask >> t) & 1
            for c in range(3):
                a, b_ = ep[c][2 * r], ep[c][2 * r + 1]
                out[t, c] = _bc6_finish((a * (64 - wgt) + b_ * wgt + 32) >> 6, signed)
        return out
    return None


def decode_bc6h(blocks, width, height, signed):
    """float32 [H, W, 3] (NaN where a block uses a mode outside 1, 2, 10..14)."""
    bx, by = (width + 3) // 4, (height + 3) // 4
    b = np.asarray(blocks, np.uint8).reshape(by, bx, 16)
    out = np.full((by * 4, bx * 4, 3), np.nan, np.float32)
    for j in range(by):
        for i in range(bx):
            h = decode_bc6h_block(b[j, i], signed)
            if h is not None:
                out[j * 4:j * 4 + 4, i * 4:i * 4 + 4] = h.reshape(4, 4, 3).view(np.float16).astype(np.float32)
    return out[:height, :width]
