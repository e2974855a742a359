"""GPU parity of cfx_resize / cfx_encode_mip_chain (Image::resize, Texture::generateMipmaps + convert) with the
reference's FreeImage_Rescale: committed vectors, the numpy restatement on fresh inputs, and size-independent
properties at the BASELINE config-5 size. Linear images are held to bit-exactness; sRGB ones to 1 float ulp
(pow() is the only operation whose rounding may differ between glibc and CUDA)."""
import os

import numpy as np
import pytest

import cuttlefish_b200 as cfx
from oracle import resize as R

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resize", "cases.npz")
CASES = {"down_odd": (18, 11, False), "mixed": (40, 9, False), "y_only": (9, 16, False), "up": (20, 20, False),
         "to_1x1": (1, 1, False), "half": (16, 12, False)}


@pytest.mark.parametrize("name", sorted(CASES))
def test_resize_matches_reference_vectors(name):
    g = np.load(GOLDEN)
    dw, dh, srgb = CASES[name]
    for f in cfx.FILTERS:
        got = cfx.resize(g[name + "/src"], dw, dh, f, srgb)
        assert np.array_equal(got, g["%s/%s" % (name, f)]), (name, f)


def test_resize_srgb_within_one_ulp():
    g = np.load(GOLDEN)
    for f in cfx.FILTERS:
        got = cfx.resize(g["half_srgb/src"], 16, 12, f, True)
        want = g["half_srgb/" + f]
        ulp = np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
        assert ulp.max() <= 1, (f, int(ulp.max()))
        assert np.array_equal(got[..., 3], want[..., 3])          # alpha does not go through the transfer function


@pytest.mark.parametrize("shape", [(517, 389, 258, 194), (258, 194, 129, 97), (100, 3, 50, 1), (31, 200, 77, 50),
                                   (640, 480, 640, 100), (1, 64, 1, 32)])
def test_resize_matches_restatement(shape):
    sw, sh, dw, dh = shape
    rng = np.random.default_rng(sw * 1000 + sh)
    img = rng.random((sh, sw, 4), dtype=np.float32) * 4 - 1
    for f in ("CatmullRom", "Box", "Cubic"):
        assert np.array_equal(cfx.resize(img, dw, dh, f), R.resize_np(img, dw, dh, f)), (shape, f)


def test_resize_same_size_is_a_copy():
    img = np.random.default_rng(3).random((9, 7, 4), dtype=np.float32)
    assert np.array_equal(cfx.resize(img, 7, 9), img)


def test_resize_rejects_bad_arguments():
    img = np.zeros((4, 4, 4), np.float32)
    with pytest.raises(cfx.CfxError) as e:
        cfx.resize(img, 0, 2)
    assert e.value.code == -1
    with pytest.raises((cfx.CfxError, KeyError)):
        cfx.resize(img, 2, 2, filter=9)


def test_mip_chain_images_match_reference_vectors():
    g = np.load(GOLDEN)
    blocks, images = cfx.encode_mip_chain(g["chain/src"], "BC7", "CatmullRom", return_images=True)
    assert len(blocks) == len(images) == 6
    for k in range(1, 6):
        assert np.array_equal(images[k], g["chain/%d" % k]), k


@pytest.mark.parametrize("fmt", ["BC1_RGB", "BC7", "ETC2_R8G8B8A8", "ASTC_6x6", "BC4"])
def test_mip_chain_blocks_equal_per_level_encode(fmt):
    """The chain call must give, level by level, exactly what Texture::convert() gives on the reference's mip images."""
    rng = np.random.default_rng(11)
    img = rng.random((52, 84, 4), dtype=np.float32)
    blocks, images = cfx.encode_mip_chain(img, fmt, "CatmullRom", return_images=True)
    want_images = R.mip_chain(img, "CatmullRom")
    assert len(blocks) == len(want_images) == 7
    for k, (b, im, want) in enumerate(zip(blocks, images, want_images)):
        assert np.array_equal(im, want), k
        assert b.size == cfx.encoded_size(fmt, im.shape[1], im.shape[0])
        assert np.array_equal(b, cfx.encode(want, fmt)), (fmt, k)


def test_mip_chain_device_equals_host_chain():
    import torch
    rng = np.random.default_rng(31)
    img = rng.random((200, 333, 4), dtype=np.float32)
    want = cfx.encode_mip_chain(img, "BC7", "Cubic")
    got = cfx.encode_mip_chain_device(torch.from_numpy(img).cuda(), "BC7", "Cubic")
    torch.cuda.synchronize()
    assert len(got) == len(want) == 9
    for k, (g, w) in enumerate(zip(got, want)):
        assert np.array_equal(g.cpu().numpy(), w), k


def test_mip_chain_from_8bit_level0_equals_widened_float_chain():
    """An RGBA8 level 0 is taken as (float)v/255, FreeImage_ConvertToRGBAF's conversion (ConversionRGBAF.cpp:116-119)."""
    import torch
    rng = np.random.default_rng(41)
    img8 = rng.integers(0, 256, (77, 130, 4), dtype=np.uint8)
    imgf = img8.astype(np.float32) / np.float32(255.0)
    want_blocks, want_images = cfx.encode_mip_chain(imgf, "BC3", "CatmullRom", return_images=True)
    got_blocks, got_images = cfx.encode_mip_chain(img8, "BC3", "CatmullRom", return_images=True)
    dev_blocks = cfx.encode_mip_chain_device(torch.from_numpy(img8).cuda(), "BC3", "CatmullRom")
    torch.cuda.synchronize()
    assert len(got_blocks) == len(want_blocks) == 8
    for k in range(8):
        if k:
            assert np.array_equal(got_images[k], want_images[k]), k
            assert np.array_equal(got_images[k], R.mip_chain(imgf, "CatmullRom")[k]), k
        assert np.array_equal(got_blocks[k], want_blocks[k]), k
        assert np.array_equal(dev_blocks[k].cpu().numpy(), want_blocks[k]), k
    # sRGB: the tabulated 8-bit path must agree with the float path (same transfer function, same rounding)
    _, a = cfx.encode_mip_chain(img8, "BC1_RGB", "Cubic", return_images=True, srgb=True)
    _, b = cfx.encode_mip_chain(imgf, "BC1_RGB", "Cubic", return_images=True, srgb=True)
    for k in range(1, 8):
        assert np.array_equal(a[k], b[k]), k
    # a 1-texel-wide image only has the pass along y: the conversion must happen there too
    col8 = rng.integers(0, 256, (33, 1, 4), dtype=np.uint8)
    _, images = cfx.encode_mip_chain(col8, "BC1_RGB", "Box", return_images=True)
    want = R.mip_chain(col8.astype(np.float32) / np.float32(255.0), "Box")
    for k in range(1, len(want)):
        assert np.array_equal(images[k], want[k]), k


def test_mip_chain_level_limit_and_errors():
    img = np.random.default_rng(5).random((16, 16, 4), dtype=np.float32)
    assert len(cfx.encode_mip_chain(img, "BC1_RGB", levels=3)) == 3
    assert len(cfx.encode_mip_chain(img, "BC1_RGB", levels=99)) == 5
    assert len(cfx.encode_mip_chain(img, "BC1_RGB", levels=0)) == 1
    with pytest.raises(cfx.CfxError) as e:
        cfx.encode_mip_chain(img, "BC7", type="SNorm")
    assert e.value.code == -2


def test_texture_generate_mipmaps_then_convert():
    rng = np.random.default_rng(21)
    img = rng.random((40, 24, 4), dtype=np.float32)
    t = cfx.Texture(24, 40)
    assert not t.generateMipmaps()                     # no image yet
    assert t.setImage(img)
    assert t.generateMipmaps("Box")
    assert t.mip_levels == 6 and t.imagesComplete()
    want = R.mip_chain(img, "Box")
    assert t.convert("ETC2_R8G8B8")
    for k, im in enumerate(want):
        assert np.array_equal(t.data(k), cfx.encode(im, "ETC2_R8G8B8")), k


def test_full_size_chain_properties():
    """BASELINE config 5 shape (4096^2, 13 levels). A constant image stays that constant on every level (the weights of
    a window sum to one), and every level of a random image equals the restatement applied to the level above it."""
    n = 4096
    const = np.empty((n, n, 4), np.float32)
    const[:] = np.array([0.25, 0.5, 0.75, 1.0], np.float32)
    blocks, images = cfx.encode_mip_chain(const, "ETC2_R8G8B8A8", "CatmullRom", return_images=True)
    assert len(images) == 13 and images[12].shape == (1, 1, 4)
    for k in range(1, 13):
        assert np.abs(images[k] - const[0, 0]).max() <= 6e-8, k
    rng = np.random.default_rng(2)
    img = rng.random((n, n, 4), dtype=np.float32)
    blocks, images = cfx.encode_mip_chain(img, "ETC2_R8G8B8A8", "CatmullRom", return_images=True)
    for k in range(4, 13):                              # 256^2 and below: seconds for the numpy restatement
        want = R.resize_np(images[k - 1], images[k].shape[1], images[k].shape[0], "CatmullRom")
        assert np.array_equal(images[k], want), k
    # level 1 spot check: a 64-row band of the 2048^2 level against the restatement's window arithmetic
    left, count, weight = R.windows(3, 2048, 4096)
    ys = np.arange(100, 104)
    rows = img[::-1].astype(np.float64)
    for y in ys:                                        # y is a bottom-up row index of level 1
        acc_rows = []
        for x in (0, 1, 777, 2047):
            # x pass for every source row in the y window, then the y pass
            col = np.zeros((count[y], 4), np.float64)
            for j in range(count[y]):
                acc = np.zeros(4, np.float64)
                for k in range(count[x]):
                    acc = acc + weight[x, k] * rows[left[y] + j, left[x] + k]
                col[j] = acc.astype(np.float32)
            acc = np.zeros(4, np.float64)
            for j in range(count[y]):
                acc = acc + weight[y, j] * col[j]
            acc_rows.append(acc.astype(np.float32))
            assert np.array_equal(images[1][2047 - y, x], acc_rows[-1]), (y, x)
