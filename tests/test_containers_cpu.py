"""Container headers against the reference's real writers: tests/golden/containers/headers.npz holds the
first 148 / 64 bytes of the DDS / KTX files Texture::save() wrote for every block (format, type, colour space, alpha type)
pair with and without a mip chain (tools/pin/make_container_goldens.py, full libcuttlefish.so), or an empty entry where the
reference has no such file. Host code only: runs without a GPU."""
import os

import numpy as np

import cuttlefish_b200 as cfx

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "containers", "headers.npz")


def test_every_header_matches_the_reference():
    z = np.load(GOLD)
    keys = [k for k in z.files if not k.endswith("__size")]
    assert len(keys) >= 900
    for k in keys:
        fmt, typ, srgb, alpha, mips, ext = k.split("__")
        levels = cfx.mip_levels(40, 24) if mips == "1" else 1
        got = cfx.container_header(ext.upper(), fmt, 40, 24, mip_levels=levels, type=typ, srgb=bool(int(srgb)), alpha=alpha)
        want = z[k]
        if want.size == 0:
            assert got is None, "%s: the reference has no such file" % k
        else:
            assert got is not None and np.array_equal(got, want), k


def test_file_sizes_follow_from_the_headers():
    # DDS: header + levels back to back; KTX: header + per level (imageSize + blocks)
    z = np.load(GOLD)
    for k in z.files:
        if not k.endswith("__size"):
            continue
        fmt, typ, srgb, alpha, mips, ext, _ = k.split("__")
        levels = cfx.mip_levels(40, 24) if mips == "1" else 1
        blocks = sum(cfx.encoded_size(fmt, max(1, 40 >> i), max(1, 24 >> i)) for i in range(levels))
        assert int(z[k][0]) == (148 + blocks if ext == "dds" else 64 + 4 * levels + blocks), k


def test_srgb_texture_refuses_formats_without_srgb_variant():
    # Texture::convert(), lib/src/Texture.cpp:1542: sRGB textures convert only to formats with a native sRGB variant
    tex = cfx.Texture(16, 16, srgb=True)
    tex.setImage(np.zeros((16, 16, 4), np.float32))
    for fmt, typ in (("BC4", "UNorm"), ("BC5", "SNorm"), ("BC6H", "UFloat"), ("ETC1", "UNorm"), ("EAC_R11", "UNorm"), ("ASTC_6x6", "UFloat")):
        assert tex.convert(fmt, typ) is False and not tex.converted()
