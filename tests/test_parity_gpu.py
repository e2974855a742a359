"""GPU parity tests (run on the B200 box): every call goes through the C-ABI (cfx_encode /
cfx_encode_device) and is compared with the committed goldens and with the CPU oracle on the
same seeded inputs.  Bit-exact for BC4/BC5 (integer) ...; PSNR within 0.1 dB for search formats."""
import numpy as np
import pytest

from util import block_mismatches, golden_cases, load_golden, src_as_float

pytestmark = pytest.mark.gpu

EXACT_FORMATS = ["BC4", "BC5"]


@pytest.mark.parametrize("name", golden_cases(EXACT_FORMATS))
def test_exact_vs_golden(cfx, name):
    src, blocks, fmt, kw = load_golden(name)
    got = cfx.encode(src, fmt, **kw)
    bad = block_mismatches(got, blocks, cfx.block_info(fmt)[2])
    assert bad.size == 0, "%s: %d blocks differ, first %s" % (name, bad.size, bad[:8])


@pytest.mark.parametrize("fmt", EXACT_FORMATS)
@pytest.mark.parametrize("kind,w,h", [("noise+grad", 256, 256), ("gradient", 256, 128), ("noise+grad", 97, 61)])
def test_exact_vs_oracle(cfx, oracle, fmt, kind, w, h):
    img = oracle.gen_image(kind, w, h, seed=777)
    ref = oracle.encode(img, fmt)
    for src in (oracle.to_rgba8(img), img):                     # RGBA8 and RGBA32F source paths
        got = cfx.encode(src, fmt)
        bad = block_mismatches(got, ref, cfx.block_info(fmt)[2])
        assert bad.size == 0, "%d blocks differ, first %s" % (bad.size, bad[:8])


@pytest.mark.parametrize("fmt", EXACT_FORMATS)
def test_device_entry_matches_host_entry(cfx, oracle, fmt):
    import torch
    img = oracle.to_rgba8(oracle.gen_image("noise+grad", 128, 64, seed=5))
    host = cfx.encode(img, fmt)
    dev = cfx.encode_device(torch.from_numpy(img).cuda(), fmt)
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), host)
