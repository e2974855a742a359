"""generateMipmaps() + convert() + save() in one call (cfx_encode_mip_chain_to_file): whole files against the reference's
real Texture::save() output for two byte-exact formats (tests/golden/containers/, tools/pin/make_container_goldens.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "containers")


@pytest.mark.parametrize("fmt,name", [("BC4", "bc4_mips.dds"), ("BC1_RGB", "bc1_mips.ktx")])
def test_whole_file_equals_the_reference(cfx, tmp_path, fmt, name):
    if not cfx.format_is_exact(fmt, "UNorm", "Normal"):
        pytest.fail("%s is not byte-exact in this build (rgbcx tables missing?)" % fmt)
    img = np.load(os.path.join(GOLD, "source_52x36.npy"))
    out = tmp_path / name
    cfx.encode_mip_chain_to_file(img, fmt, str(out), filter="CatmullRom")
    got, want = np.fromfile(out, np.uint8), np.fromfile(os.path.join(GOLD, name), np.uint8)
    assert got.size == want.size
    assert np.array_equal(got, want), "first difference at byte %d" % int(np.argmax(got != want))


def test_unsupported_container_leaves_no_file(cfx, tmp_path):
    img = np.load(os.path.join(GOLD, "source_52x36.npy"))
    out = tmp_path / "astc.dds"                      # DDS knows no ASTC format (isValidForDds)
    with pytest.raises(cfx.CfxError):
        cfx.encode_mip_chain_to_file(img, "ASTC_6x6", str(out))
    assert not out.exists()
