"""The reference-side adapter as CODE: adapter/CudaConverter.cpp compiled against the reference's own
lib/src/Converter.h and hooked into the reference's real createConverter() (oracle/_ref/libcfglue_cuda.so, built by
oracle/Makefile from /root/reference where it lies). Converter::convert() -> CudaConverter -> cfx_encode() must give
exactly the bytes of a direct cfx_encode() call, for top-down and for bottom-up Image storage."""
import numpy as np
import pytest


def _need(oracle):
    if not oracle.glue_cuda_available():
        pytest.fail("oracle/_ref/libcfglue_cuda.so missing: run __graft_entry__.build() where /root/reference is mounted")


@pytest.mark.gpu
@pytest.mark.parametrize("fmt,type_,kind", [("BC7", "UNorm", "noise+grad"), ("BC1_RGB", "UNorm", "gradient"),
                                           ("BC3", "UNorm", "noise+grad"), ("BC4", "SNorm", "noise+grad"),
                                           ("BC6H", "UFloat", "hdr"), ("ETC2_R8G8B8A8", "UNorm", "noise+grad"),
                                           ("EAC_R11", "UNorm", "noise+grad"), ("ASTC_6x6", "UNorm", "noise+grad"),
                                           ("ASTC_10x8", "UNorm", "noise+grad")])
@pytest.mark.parametrize("bottom_up", [False, True])
def test_converter_convert_goes_through_cfx(cfx, oracle, fmt, type_, kind, bottom_up):
    _need(oracle)
    img = oracle.gen_image(kind, 203, 117, seed=21).astype(np.float32)
    cfx.init(0)
    want = cfx.encode(img, fmt, type=type_)
    got, secs, on_gpu = oracle.encode_glue_cuda(img, fmt, type=type_, bottom_up=bottom_up)
    assert on_gpu == 1, "the surface did not take the GPU path"
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_unsupported_pair_falls_through_to_the_cpu_converters(cfx, oracle):
    """R8G8B8A8 has no GPU encoder: createConverter() must carry on to the stock converter (cfx has no fallback of its own)."""
    _need(oracle)
    img = oracle.gen_image("noise+grad", 16, 8, seed=1)
    got, _, on_gpu = oracle.encode_glue_cuda(img, 14, out_bytes=16 * 8 * 4)
    assert on_gpu == 0
    assert np.array_equal(got, oracle.encode_glue(img, 14, out_bytes=16 * 8 * 4))


@pytest.mark.gpu
def test_quality_alpha_mask_and_srgb_reach_the_kernel(cfx, oracle):
    _need(oracle)
    img = oracle.gen_image("noise+grad", 64, 64, seed=2).astype(np.float32)
    img[..., 3] = np.linspace(0, 1, 64, dtype=np.float32)[None, :]
    for kw in (dict(quality="Lowest"), dict(quality="Highest"), dict(alpha="None"), dict(color_mask=5), dict(srgb=True)):
        want = cfx.encode(img, "BC7", **kw)
        got, _, on_gpu = oracle.encode_glue_cuda(img, "BC7", **kw)
        assert on_gpu == 1 and np.array_equal(got, want), kw


def test_adapter_falls_back_without_gpu(oracle):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    if not oracle.glue_cuda_available():
        pytest.skip("adapter library not built")
    img = oracle.gen_image("noise+grad", 32, 24, seed=4)
    got, _, on_gpu = oracle.encode_glue_cuda(img, "BC1_RGB", bottom_up=True)
    assert on_gpu == 0
    assert np.array_equal(got, oracle.encode_glue(img, "BC1_RGB"))
