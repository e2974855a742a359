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


GLUE_FORMATS = ["BC1_RGB", "BC1_RGBA", "BC2", "BC3", "BC4", "BC5", "BC7", "ETC1", "ETC2_R8G8B8", "ETC2_R8G8B8A1",
                "ETC2_R8G8B8A8", "EAC_R11", "EAC_R11G11", "ASTC_4x4", "ASTC_6x6", "ASTC_10x6", "ASTC_12x12"]


@pytest.mark.parametrize("fmt", GLUE_FORMATS)
def test_restated_glue_matches_reference_glue(oracle, fmt):
    """Pins oracle/cfref.cpp (our restatement of the Converter glue) against the reference's REAL glue:
    lib/src/Converter.cpp + *Converter.cpp compiled from /root/reference over oracle/glue_stub.cpp.
    Byte-identical output is required for every format, ragged sizes, transparent texels, qualities."""
    if not oracle.glue_available():
        pytest.skip("oracle/_ref/libcfglue.so not built (needs /root/reference)")
    for kind, w, h in [("noise+grad", 48, 40), ("gradient", 30, 22)]:
        img = oracle.gen_image(kind, w, h, seed=77)
        img[:8, :8, 3] = 0.25                       # transparent corner: BC1_RGBA's squish path, EAC, ASTC alpha
        for q in (["Normal", "Low", "High"] if kind == "gradient" else ["Normal"]):
            assert np.array_equal(oracle.encode(img, fmt, quality=q), oracle.encode_glue(img, fmt, quality=q)), (fmt, kind, q)
    img = oracle.gen_image("noise+grad", 24, 24)
    assert np.array_equal(oracle.encode(img, fmt, color_mask=5), oracle.encode_glue(img, fmt, color_mask=5))
    assert np.array_equal(oracle.encode(img, fmt, srgb=True), oracle.encode_glue(img, fmt, srgb=True))


@pytest.mark.parametrize("fmt", ["BC4", "BC5", "EAC_R11", "EAC_R11G11"])
def test_restated_glue_matches_reference_glue_snorm(oracle, fmt):
    if not oracle.glue_available():
        pytest.skip("oracle/_ref/libcfglue.so not built (needs /root/reference)")
    img = (oracle.gen_image("noise+grad", 30, 22, seed=5)*2.0 - 1.0).astype(np.float32)
    img[..., 3] = 1.0
    assert np.array_equal(oracle.encode(img, fmt, type="SNorm"), oracle.encode_glue(img, fmt, type="SNorm"))


def test_restated_glue_matches_reference_glue_bc6h(oracle):
    if not oracle.glue_available():
        pytest.skip("oracle/_ref/libcfglue.so not built (needs /root/reference)")
    hdr = oracle.gen_image("hdr", 30, 22)
    for q in ("Normal", "Low"):
        assert np.array_equal(oracle.encode(hdr, "BC6H", type="UFloat", quality=q), oracle.encode_glue(hdr, "BC6H", type="UFloat", quality=q))


def test_bc6h_spec_decoder_pinned_on_reference_decoder(oracle):
    """tests/util.py decode_bc6h (used to check the SIGNED format, where the reference has no usable decode) agrees
    with the reference's decoder on unsigned blocks: exactly on the reference encoder's blocks, within one half ulp
    everywhere it covers."""
    from util import decode_bc6h
    covered = 0
    for name in golden_cases(["BC6H"]):
        src, blocks, fmt, kw = load_golden(name)
        h, w, _ = src.shape
        mine = decode_bc6h(blocks, w, h, False)
        ref = oracle.decode(blocks, "BC6H", w, h, **kw)[..., :3]
        ok = np.isfinite(mine)
        covered += int(ok.sum())
        assert np.array_equal(mine[ok], ref[ok])
    assert covered > 0
