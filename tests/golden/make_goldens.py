"""Regenerates tests/golden/*.npz from the CPU oracle (the reference's encoders built by
oracle/Makefile from /root/reference). Run in the build container only:

    python tests/golden/make_goldens.py

Each .npz holds the RGBA8 (or RGBA16F bit pattern) input and the reference encoder's packed
blocks for one (format, input) case at small sizes, so parity can be checked on the GPU box
without /root/reference.  The reference's own tests hold no byte-level vectors for this path
(SURVEY.md section 4), so these outputs of the reference itself are the pin.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

LDR_FORMATS = ["BC1_RGB", "BC1_RGBA", "BC2", "BC3", "BC4", "BC5", "BC7", "ETC1", "ETC2_R8G8B8",
               "ETC2_R8G8B8A8", "ASTC_4x4", "ASTC_6x6", "ASTC_8x8"]


def alpha_variant(img):
    """Adds a varying alpha channel (ramp + a fully transparent and a fully opaque region)."""
    h, w, _ = img.shape
    out = img.copy()
    yy, xx = np.mgrid[0:h, 0:w]
    a = ((xx * 7 + yy * 13) % 256).astype(np.float32) / np.float32(255.0)
    a[: h // 4, : w // 4] = 0.0
    a[h // 2:, w // 2:] = 1.0
    out[..., 3] = a
    return out


def cases():
    for kind, w, h in [("noise+grad", 64, 64), ("gradient", 64, 64), ("noise+grad", 30, 22)]:
        img = oracle.gen_image(kind, w, h)
        for fmt in LDR_FORMATS:
            yield "%s_%s_%dx%d" % (fmt, kind.replace("+", ""), w, h), fmt, {}, img
    img = alpha_variant(oracle.gen_image("noise+grad", 32, 32))
    for fmt in ["BC1_RGBA", "BC2", "BC3", "BC7", "ETC2_R8G8B8A8", "ASTC_6x6"]:
        yield "%s_alpha_32x32" % fmt, fmt, {}, img
    img = oracle.gen_image("noise+grad", 32, 32)
    for q in ["Lowest", "Low", "High", "Highest"]:
        for fmt in ["BC4", "BC5", "BC1_RGB", "BC3", "BC7", "ETC1"]:
            yield "%s_%s_32x32" % (fmt, q), fmt, {"quality": q}, img
    hdr = oracle.gen_image("hdr", 64, 64)
    yield "BC6H_hdr_64x64", "BC6H", {"type": "UFloat"}, hdr
    hdr = oracle.gen_image("hdr", 30, 22)
    yield "BC6H_hdr_30x22", "BC6H", {"type": "UFloat"}, hdr
    # formats added later in round 1: the rest of the ETC/EAC family, signed BC4/BC5, the other ASTC footprints
    img = oracle.gen_image("noise+grad", 32, 32)
    for fmt in ["EAC_R11", "EAC_R11G11", "ETC2_R8G8B8A1", "ASTC_5x4", "ASTC_5x5", "ASTC_6x5", "ASTC_8x5", "ASTC_8x6",
                "ASTC_10x5", "ASTC_10x6", "ASTC_10x8", "ASTC_10x10", "ASTC_12x10", "ASTC_12x12"]:
        yield "%s_noisegrad_32x32" % fmt, fmt, {}, img
    punch = alpha_variant(img)
    punch[..., 3] = (punch[..., 3] >= 0.5).astype(np.float32)
    yield "ETC2_R8G8B8A1_alpha_32x32", "ETC2_R8G8B8A1", {}, punch
    signed = (img*np.float32(2.0) - np.float32(1.0)).astype(np.float16).astype(np.float32)
    signed[..., 3] = 1.0
    for fmt in ["BC4", "BC5", "EAC_R11", "EAC_R11G11"]:
        yield "%s_snorm_32x32" % fmt, fmt, {"type": "SNorm"}, signed
    ui = oracle.gen_image("ui", 48, 48)
    for fmt in ["ASTC_4x4", "ASTC_6x6", "ASTC_8x8", "BC7", "ETC2_R8G8B8"]:
        yield "%s_ui_48x48" % fmt, fmt, {}, ui


def main():
    n = 0
    for name, fmt, kw, img in cases():
        blocks = oracle.encode(img, fmt, threads=0, **kw)
        if kw.get("type") in ("UFloat", "SNorm"):
            src = img.astype(np.float16).view(np.uint16)      # RNE, same as the kernels' f32->f16
        else:
            src = oracle.to_rgba8(img)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), src=src, blocks=blocks,
                            format=fmt, kw=repr(kw))
        n += 1
    print("wrote %d golden cases" % n)


if __name__ == "__main__":
    main()
