"""GPU tests of the host-buffer pipeline of cfx_encode / cfx_encode_batch: the device pool (one call sharding a surface
by block row over several contexts -- two contexts on ONE GPU exercise the same code on a one-GPU box, two real GPUs
when the box has them), pinned vs pageable buffers, the narrowing of a pageable RGBA32F source on its way through the
staging slots, and bottom-up sources (cuttlefish::Image's storage order). Every variant must give the bytes of the
plain one-device, top-down, pinned-free call."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FORMATS = ["BC7", "BC1_RGB", "BC3", "BC5", "ETC2_R8G8B8A8", "ASTC_6x6"]


def _pinned(cfx, arr):
    """A pinned copy of arr (cfx_host_alloc), returned with a keep-alive handle."""
    from cuttlefish_b200 import _lib
    lib = _lib.load()
    ptr = lib.cfx_host_alloc(arr.nbytes)
    assert ptr
    buf = (ctypes.c_uint8 * arr.nbytes).from_address(ptr)
    out = np.frombuffer(buf, dtype=arr.dtype).reshape(arr.shape)
    out[...] = arr
    return out, ptr


@pytest.fixture()
def pool_reset(cfx):
    yield cfx
    cfx.init(0)


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("w,h", [(1024, 1024), (1030, 778)])
def test_two_contexts_equal_one(pool_reset, oracle, fmt, w, h):
    cfx = pool_reset
    img = oracle.to_rgba8(oracle.gen_image("noise+grad", w, h, seed=31))
    cfx.init(0)
    assert cfx.device_count() == 1
    want = cfx.encode(img, fmt)
    cfx.set_devices([0, 0])
    assert cfx.device_count() == 2
    got = cfx.encode(img, fmt)
    assert np.array_equal(got, want)
    cfx.set_devices([0, 0, 0])
    assert np.array_equal(cfx.encode(img, fmt), want)


def test_two_gpus_equal_one(pool_reset, oracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    cfx = pool_reset
    for fmt, (w, h) in [("BC7", (2048, 2048)), ("ASTC_6x6", (1030, 778)), ("BC1_RGB", (4096, 1024))]:
        img = oracle.to_rgba8(oracle.gen_image("noise+grad", w, h, seed=32))
        cfx.init(0)
        want = cfx.encode(img, fmt)
        cfx.init_devices(2)
        assert cfx.device_count() == 2
        assert np.array_equal(cfx.encode(img, fmt), want), fmt
        cfx.init(1)
        assert np.array_equal(cfx.encode(img, fmt), want), fmt + " on device 1"
    cfx.init_devices(0)
    assert cfx.device_count() == torch.cuda.device_count()


def test_device_pointer_calls_follow_the_pointer(pool_reset, oracle):
    """cfx_encode_device runs on the device that owns d_src, whatever the pool says, and leaves the caller's current
    device alone (one global context silently moved callers between GPUs)."""
    import torch
    cfx = pool_reset
    img = oracle.to_rgba8(oracle.gen_image("noise+grad", 256, 128, seed=5))
    cfx.init(0)
    want = cfx.encode(img, "BC7")
    last = torch.cuda.device_count() - 1
    cfx.init(last)
    for dev in sorted({0, last}):
        t = torch.from_numpy(img).to("cuda:%d" % dev)
        torch.cuda.set_device(0)
        got = cfx.encode_device(t, "BC7")
        assert torch.cuda.current_device() == 0
        torch.cuda.synchronize(dev)
        assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("fmt,type_,kind", [("BC7", "UNorm", "noise+grad"), ("BC1_RGB", "UNorm", "noise+grad"),
                                           ("BC5", "UNorm", "noise+grad"), ("BC5", "SNorm", "noise+grad"),
                                           ("BC6H", "UFloat", "hdr"), ("ETC2_R8G8B8", "UNorm", "noise+grad"),
                                           ("ASTC_6x6", "UNorm", "noise+grad")])
def test_pageable_float_source_equals_pinned(pool_reset, oracle, fmt, type_, kind):
    """A pageable RGBA32F source is narrowed on the host on its way through the staging slots (RGBA8 / RGBA16F where the
    kernel's load stage computes exactly that view); a pinned one is DMA'd as it is and narrowed in the kernel. Same
    blocks either way. The image is NOT 8-bit snapped and leaves [0, 1], so the rounding and the clamp both matter."""
    cfx = pool_reset
    rng = np.random.default_rng(7)
    img = oracle.gen_image(kind, 1536, 1024, seed=11).astype(np.float32)
    img[..., :3] += rng.uniform(-0.004, 0.004, img[..., :3].shape).astype(np.float32)
    img[100:110, 200:260, :3] = -0.25
    img[300:310, 200:260, :3] = 1.5 if kind != "hdr" else 70000.0
    cfx.init(0)
    pinned, ptr = _pinned(cfx, img)
    try:
        want = cfx.encode(pinned, fmt, type=type_).copy()
        got = cfx.encode(img, fmt, type=type_)
        assert np.array_equal(got, want)
        cfx.set_devices([0, 0])
        assert np.array_equal(cfx.encode(img, fmt, type=type_), want)
    finally:
        from cuttlefish_b200 import _lib
        _lib.load().cfx_host_free(ptr)


@pytest.mark.parametrize("fmt", ["BC7", "BC4", "ETC1", "ASTC_8x8", "BC6H"])
@pytest.mark.parametrize("w,h", [(512, 256), (97, 61)])
def test_bottom_up_source(pool_reset, oracle, fmt, w, h):
    """CFX_FLAG_BOTTOM_UP: the same image stored with its rows reversed gives the same blocks, from RGBA32F (what the
    adapter passes: pageable, narrowed on the host for BC7/BC4/BC6H) and from RGBA8, on one context and on two."""
    cfx = pool_reset
    kind, type_ = ("hdr", "UFloat") if fmt == "BC6H" else ("noise+grad", "UNorm")
    img = oracle.gen_image(kind, w, h, seed=3).astype(np.float32)
    variants = [img] if fmt == "BC6H" else [img, oracle.to_rgba8(img)]
    for src in variants:
        flipped = np.ascontiguousarray(src[::-1])
        for pool in ([0], [0, 0]):
            cfx.set_devices(pool)
            want = cfx.encode(src, fmt, type=type_)
            got = cfx.encode(flipped, fmt, type=type_, bottom_up=True)
            assert np.array_equal(got, want), (fmt, src.dtype, pool)
    import torch
    cfx.init(0)
    src = oracle.to_rgba8(img) if fmt != "BC6H" else img.astype(np.float16)
    want = cfx.encode(src, fmt, type=type_)
    dev = cfx.encode_device(torch.from_numpy(np.ascontiguousarray(src[::-1])).cuda(), fmt, type=type_, bottom_up=True)
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), want)


def test_batch_over_two_contexts(pool_reset, oracle):
    """A mip chain as one batch: big levels are cut across the pool, small ones go whole to one context."""
    cfx = pool_reset
    levels = []
    img = oracle.gen_image("noise+grad", 1024, 1024, seed=77)
    while True:
        levels.append(oracle.to_rgba8(img))
        if img.shape[0] == 1:
            break
        img = img.reshape(img.shape[0] // 2, 2, img.shape[1] // 2, 2, 4).mean(axis=(1, 3)).astype(np.float32)
    cfx.init(0)
    want = cfx.encode_batch(levels, "BC3")
    cfx.set_devices([0, 0])
    got = cfx.encode_batch(levels, "BC3")
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


def test_large_surface_chunks(pool_reset, oracle):
    """4096 x 4096: several chunks per context; blocks equal the device-entry encode of the whole surface."""
    import torch
    cfx = pool_reset
    img = oracle.to_rgba8(oracle.gen_image("noise+grad", 4096, 4096, seed=12345))
    cfx.init(0)
    want = cfx.encode_device(torch.from_numpy(img).cuda(), "BC1_RGB")
    torch.cuda.synchronize()
    want = want.cpu().numpy()
    for pool in ([0], [0, 0]):
        cfx.set_devices(pool)
        assert np.array_equal(cfx.encode(img, "BC1_RGB"), want)


def test_unaligned_device_pointer_is_rejected(pool_reset):
    import torch
    cfx = pool_reset
    from cuttlefish_b200 import _lib
    lib = _lib.load()
    t = torch.zeros(64 * 64 * 16 + 16, dtype=torch.uint8, device="cuda")
    out = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    d = cfx.api.make_desc("BC7", 16, 16, "RGBA32F", 16 * 16)
    assert lib.cfx_encode_device(ctypes.byref(d), t.data_ptr() + 4, out.data_ptr(), 4096, None) == -1
    assert b"texel aligned" in lib.cfx_last_error()
    assert lib.cfx_encode_device(ctypes.byref(d), t.data_ptr(), out.data_ptr(), 4096, None) == 0
    torch.cuda.synchronize()
