"""CPU tests: the oracle (reference encoders built from /root/reference) reproduces every
committed golden vector, and its decoders agree with the encoders (PSNR sanity)."""
import numpy as np
import pytest

from util import golden_cases, load_golden, src_as_float


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_golden(oracle, name):
    src, blocks, fmt, kw = load_golden(name)
    got = oracle.encode(src_as_float(src), fmt, threads=0, **kw)
    assert got.size == blocks.size
    assert np.array_equal(got, blocks), "oracle output drifted from committed golden %s" % name


def test_oracle_thread_count_invariant(oracle):
    img = oracle.gen_image("noise+grad", 64, 64)
    for fmt in ["BC1_RGB", "BC7", "ETC1", "ASTC_6x6"]:
        assert np.array_equal(oracle.encode(img, fmt, threads=1), oracle.encode(img, fmt, threads=4))


@pytest.mark.parametrize("fmt,floor", [("BC1_RGB", 40.0), ("BC3", 40.0), ("BC7", 45.0), ("ETC1", 35.0),
                                       ("ETC2_R8G8B8", 38.0), ("ASTC_6x6", 38.0)])
def test_oracle_decode_psnr_gradient(oracle, fmt, floor):
    img = oracle.gen_image("gradient", 128, 128)
    dec = oracle.decode(oracle.encode(img, fmt), fmt, 128, 128)
    assert oracle.psnr_rgb(img, dec) > floor


def test_generator_is_8bit_snapped(oracle):
    img = oracle.gen_image("noise+grad", 33, 17)
    u8 = oracle.to_rgba8(img)
    assert np.array_equal(u8.astype(np.float32) / np.float32(255.0), img)
