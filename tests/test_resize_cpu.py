"""CPU tests of the resize / mip-chain oracle: the numpy restatement (oracle/resize.py) against the committed vectors
produced by the reference's real FreeImage_Rescale (tests/golden/make_resize_goldens.py), and, where the compiled
reference is present, against it directly on fresh inputs. Bit-exact: the arithmetic is double precision in a fixed
order."""
import os

import numpy as np
import pytest

from oracle import resize as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resize", "cases.npz")
CASES = {"down_odd": (18, 11, False), "mixed": (40, 9, False), "y_only": (9, 16, False), "up": (20, 20, False),
         "to_1x1": (1, 1, False), "half": (16, 12, False), "half_srgb": (16, 12, True)}


def same_bits(a, b):
    """Equal as floats everywhere (no NaNs expected; -0.0 == +0.0)."""
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b)


@pytest.mark.parametrize("name", sorted(CASES))
def test_restatement_matches_reference_vectors(name):
    g = np.load(GOLDEN)
    dw, dh, srgb = CASES[name]
    for f in R.FILTERS:
        got = R.resize_np(g[name + "/src"], dw, dh, f, srgb)
        assert same_bits(got, g["%s/%s" % (name, f)]), (name, f)


def test_mip_chain_matches_reference_vectors():
    g = np.load(GOLDEN)
    chain = R.mip_chain(g["chain/src"], "CatmullRom")
    assert [c.shape[:2] for c in chain] == [(20, 48), (10, 24), (5, 12), (2, 6), (1, 3), (1, 1)]
    for k, level in enumerate(chain[1:], 1):
        assert same_bits(level, g["chain/%d" % k]), k


def test_mip_sizes_follow_the_reference():
    # Texture::maxMipmapLevels (lib/src/Texture.cpp:514-527): 32 - clz(max(w, h))
    assert len(R.mip_sizes(4096, 4096)) == 13
    assert R.mip_sizes(5, 3) == [(5, 3), (2, 1), (1, 1)]
    assert R.mip_sizes(8, 8, levels=2) == [(8, 8), (4, 4)]
    assert R.mip_sizes(8, 8, levels=0) == [(8, 8)]


def test_windows_are_normalised_and_trimmed():
    for f in range(5):
        for (d, s) in [(18, 37), (40, 16), (1, 3), (2048, 4096)]:
            left, count, weight = R.windows(f, d, s)
            assert (count >= 1).all() and (left >= 0).all() and (left + count <= s).all()
            assert np.allclose(weight.sum(1), 1.0, atol=1e-12)
            last = weight[np.arange(d), count - 1]
            assert (last != 0).all() or (count == 1).any()


@pytest.mark.skipif(not R.ref_available(), reason="oracle/_ref/libfiresize.so not built")
def test_restatement_matches_compiled_reference():
    rng = np.random.default_rng(7)
    for (sw, sh, dw, dh) in [(61, 47, 30, 23), (30, 23, 15, 11), (13, 64, 50, 7), (64, 13, 64, 40), (2, 2, 1, 1), (1, 9, 1, 4)]:
        img = rng.random((sh, sw, 4), dtype=np.float32) * 4 - 1
        for f in R.FILTERS:
            assert same_bits(R.resize_np(img, dw, dh, f), R.resize_ref(img, dw, dh, f)), (sw, sh, dw, dh, f)
        img = rng.random((sh, sw, 4), dtype=np.float32)
        assert same_bits(R.resize_np(img, dw, dh, "CatmullRom", True), R.resize_ref(img, dw, dh, "CatmullRom", True))
