"""CPU tests of the drop-in boundary: libcfx.so loads, exports every symbol include/cfx.h
declares, answers the size/format queries like the reference's tables, validates descriptors,
and fails loudly (no CPU fallback) when no GPU is present."""
import ctypes
import os
import re

import numpy as np
import pytest

import cuttlefish_b200 as cfx
from cuttlefish_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    header = open(os.path.join(ROOT, "include", "cfx.h")).read()
    declared = set(re.findall(r"\b(cfx_[a-z_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libcfx.so does not export %s" % name
    assert declared == {s[0] for s in _lib.SYMBOLS}


def test_block_tables_match_reference():
    # Texture::blockWidth/Height/Size, lib/src/Texture.cpp:529-773
    expect = {"BC1_RGB": (4, 4, 8), "BC1_RGBA": (4, 4, 8), "BC2": (4, 4, 16), "BC3": (4, 4, 16),
              "BC4": (4, 4, 8), "BC5": (4, 4, 16), "BC6H": (4, 4, 16), "BC7": (4, 4, 16),
              "ETC1": (4, 4, 8), "ETC2_R8G8B8": (4, 4, 8), "ETC2_R8G8B8A1": (4, 4, 8),
              "ETC2_R8G8B8A8": (4, 4, 16), "EAC_R11": (4, 4, 8), "EAC_R11G11": (4, 4, 16),
              "ASTC_4x4": (4, 4, 16), "ASTC_6x6": (6, 6, 16), "ASTC_10x5": (10, 5, 16),
              "ASTC_12x12": (12, 12, 16)}
    for fmt, v in expect.items():
        assert cfx.block_info(fmt) == v
    # TextureConvertSpecialTest: 16x16 -> blocksX*blocksY*blockSize (lib/test/TextureTest.cpp:847-868)
    assert cfx.encoded_size("BC1_RGB", 16, 16) == 16 * 8
    assert cfx.encoded_size("BC7", 16, 16) == 16 * 16
    assert cfx.encoded_size("ASTC_6x6", 16, 16) == 9 * 16
    assert cfx.encoded_size("ASTC_12x12", 16, 16) == 4 * 16
    assert cfx.encoded_size("BC7", 8192, 8192) == 64 << 20
    assert cfx.encoded_size("ASTC_6x6", 8192, 8192) == 29855296
    assert cfx.encoded_size("BC7", 30, 22) == 8 * 6 * 16
    assert cfx.encoded_size(14, 16, 16) == 0          # R8G8B8A8 is not block compressed


def test_unsupported_pairs_report_unsupported():
    assert not cfx.format_supported("BC4", "UInt")
    assert not cfx.format_supported(14, "UNorm")
    with pytest.raises(cfx.CfxError) as e:
        cfx.encode(np.zeros((4, 4, 4), np.uint8), "EAC_R11", type="UInt")
    assert e.value.code == -2
    with pytest.raises(cfx.CfxError) as e:
        cfx.encode(np.zeros((12, 12, 4), np.float16), "ASTC_12x12", type="SNorm")       # no such converter in the reference either
    assert e.value.code == -2


def test_descriptor_validation():
    lib = _lib.load()
    d = cfx.api.make_desc("BC4", 0, 4, "RGBA8", 16)
    buf = (ctypes.c_uint8 * 64)()
    assert lib.cfx_encode(ctypes.byref(d), buf, buf, 64) == -1
    d = cfx.api.make_desc("BC4", 8, 8, "RGBA8", 16)          # pitch smaller than a row
    assert lib.cfx_encode(ctypes.byref(d), buf, buf, 64) == -1
    assert b"pitch" in lib.cfx_last_error()
    d = cfx.api.make_desc("BC4", 8, 8, "RGBA8", 32)
    d.quality = 9
    assert lib.cfx_encode(ctypes.byref(d), buf, buf, 64) == -1
    assert lib.cfx_encode(None, buf, buf, 64) == -1


def test_shard_block_rows_partition():
    for h, bh, world in [(8192, 4, 8), (8192, 6, 8), (30, 4, 4), (5, 4, 8)]:
        rows = (h + bh - 1) // bh
        spans = [cfx.shard_block_rows(h, bh, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == rows
        for a, b in zip(spans, spans[1:]):
            assert a[1] == b[0] and a[3] == b[2] or a[1] == b[0]
        assert spans[-1][3] == h


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(cfx.CfxError) as e:
        cfx.encode(np.zeros((8, 8, 4), np.uint8), "BC4")
    assert e.value.code in (-3, -4)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_mip_levels_follow_the_reference():
    # Texture::maxMipmapLevels for 2D textures (lib/src/Texture.cpp:514-527): 32 - clz(max(w, h))
    assert cfx.mip_levels(4096, 4096) == 13
    assert cfx.mip_levels(5, 3) == 3
    assert cfx.mip_levels(1, 1) == 1
    assert cfx.mip_levels(1, 1024) == 11


def test_resize_and_mip_chain_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    img = np.zeros((8, 8, 4), np.float32)
    with pytest.raises(cfx.CfxError) as e:
        cfx.resize(img, 4, 4)
    assert e.value.code in (-3, -4)
    with pytest.raises(cfx.CfxError) as e:
        cfx.encode_mip_chain(img, "BC1_RGB")
    assert e.value.code in (-3, -4)
    # argument errors are reported before any device work
    lib = _lib.load()
    buf = (ctypes.c_float * 64)()
    assert lib.cfx_resize(buf, 0, 4, 64, buf, 2, 2, 32, 3, 0) == -1
    assert lib.cfx_resize(buf, 4, 4, 64, buf, 2, 2, 32, 9, 0) == -1          # no such filter
    assert lib.cfx_resize(buf, 4, 4, 8, buf, 2, 2, 32, 3, 0) == -1           # pitch smaller than a row
    d = cfx.api.make_desc("BC1_RGB", 8, 8, "RGBA16F", 64)
    sizes = (ctypes.c_size_t * 4)(32, 8, 8, 8)
    dst = (ctypes.c_void_p * 4)(*[ctypes.addressof(buf)] * 4)
    assert lib.cfx_encode_mip_chain(ctypes.byref(d), buf, 3, 4, dst, sizes, None) == -1    # level 0 must be RGBA32F or RGBA8
    assert b"RGBA32F" in lib.cfx_last_error()
