"""Regenerates tests/golden/real/*.npz: 192x192 crops of the images the reference itself ships (astcenc's Small test
set, Compressonator's ruby.png) together with the reference CPU encoders' blocks for them, so that real-image parity
can be checked on the GPU box, where /root/reference does not exist. Run in the build container only:

    python tests/golden/make_real_goldens.py

Each .npz: `src` = the RGBA8 crop (or float16 bits for the HDR image) and one `blocks__<FORMAT>__<Quality>` array per
encoded variant (reference output through the reference's own Converter glue, oracle/_ref).
"""
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

REF = os.environ.get("CFX_REFERENCE", "/root/reference")
SMALL = REF + "/lib/astc-encoder/Test/Images/Small/"
N = 192

# (name, path, crop origin)
LDR = [("rgb00", SMALL + "LDR-RGB/ldr-rgb-00.png", (32, 32)), ("rgb03", SMALL + "LDR-RGB/ldr-rgb-03.png", (32, 32)),
       ("rgb05", SMALL + "LDR-RGB/ldr-rgb-05.png", (0, 64)), ("rgb07", SMALL + "LDR-RGB/ldr-rgb-07.png", (64, 0)),
       ("rgb09", SMALL + "LDR-RGB/ldr-rgb-09.png", (32, 32)),
       ("rgba00", SMALL + "LDR-RGBA/ldr-rgba-00.png", (32, 32)), ("rgba01", SMALL + "LDR-RGBA/ldr-rgba-01.png", (32, 32)),
       ("rgba02", SMALL + "LDR-RGBA/ldr-rgba-02.png", (32, 32)),
       ("ruby", REF + "/lib/compressonator/runtime/images/ruby.png", (200, 100))]
FORMATS_NORMAL = ["BC1_RGB", "BC3", "BC7", "ETC1", "ETC2_R8G8B8", "ETC2_R8G8B8A8", "ASTC_4x4", "ASTC_6x6", "ASTC_8x8", "ASTC_10x8"]
# the other Texture::Quality levels, on a subset (astcenc's exhaustive preset takes ~20 s per crop)
LEVEL_IMAGES = ["rgb00", "rgb07", "rgba01"]
LEVEL_FORMATS = ["BC7", "ASTC_6x6", "ETC2_R8G8B8A8", "BC1_RGB", "ETC1"]
LEVELS = ["Lowest", "Low", "High", "Highest"]


def read_rgbe(path):
    """Minimal Radiance .hdr reader (new-style RLE scanlines)."""
    with open(path, "rb") as f:
        data = f.read()
    pos = data.index(b"\n\n") + 2
    end = data.index(b"\n", pos)
    dims = data[pos:end].split()
    h, w = int(dims[1]), int(dims[3])
    pos = end + 1
    out = np.zeros((h, w, 4), np.uint8)
    for y in range(h):
        assert data[pos] == 2 and data[pos + 1] == 2
        pos += 4
        for c in range(4):
            x = 0
            while x < w:
                n = data[pos]; pos += 1
                if n > 128:
                    n -= 128
                    out[y, x:x + n, c] = data[pos]; pos += 1
                else:
                    out[y, x:x + n, c] = np.frombuffer(data[pos:pos + n], np.uint8); pos += n
                x += n
    e = out[..., 3].astype(np.int32)
    scale = np.where(e > 0, np.ldexp(1.0, e - 136), 0.0).astype(np.float32)
    img = np.ones((h, w, 4), np.float32)
    img[..., :3] = out[..., :3].astype(np.float32) * scale[..., None]
    return img


def main():
    os.makedirs(os.path.join(HERE, "real"), exist_ok=True)
    for name, path, (x0, y0) in LDR:
        rgba = np.array(Image.open(path).convert("RGBA"))
        src = np.ascontiguousarray(rgba[y0:y0 + N, x0:x0 + N])
        assert src.shape == (N, N, 4), (name, rgba.shape)
        img = src.astype(np.float32) / np.float32(255.0)
        arrays = {"src": src}
        # (variants already in the file are kept: adding a format does not re-run astcenc's exhaustive preset)
        path_out = os.path.join(HERE, "real", name + ".npz")
        if os.path.exists(path_out):
            old = np.load(path_out)
            if np.array_equal(old["src"], src):
                arrays.update({k: old[k] for k in old.files})
        for fmt in FORMATS_NORMAL:
            if "blocks__%s__Normal" % fmt not in arrays:
                arrays["blocks__%s__Normal" % fmt] = oracle.encode_glue(img, fmt, threads=0)
        if name in LEVEL_IMAGES:
            for fmt in LEVEL_FORMATS:
                for q in LEVELS:
                    if "blocks__%s__%s" % (fmt, q) not in arrays:
                        arrays["blocks__%s__%s" % (fmt, q)] = oracle.encode_glue(img, fmt, threads=0, quality=q)
        np.savez_compressed(os.path.join(HERE, "real", name + ".npz"), **arrays)
        print(name, len(arrays) - 1, "variants", flush=True)
    hdr = read_rgbe(SMALL + "HDR-RGB/hdr-rgb-00.hdr")
    src = np.ascontiguousarray(hdr[32:32 + N, 32:32 + N]).astype(np.float16)
    img = src.astype(np.float32)
    arrays = {"src": src.view(np.uint16)}
    for q in ["Low", "Normal", "High"]:
        arrays["blocks__BC6H__%s" % q] = oracle.encode_glue(img, "BC6H", threads=0, type="UFloat", quality=q)
    arrays["blocks__ASTC_6x6__Normal"] = oracle.encode_glue(img, "ASTC_6x6", threads=0, type="UFloat")
    np.savez_compressed(os.path.join(HERE, "real", "hdr00.npz"), **arrays)
    print("hdr00", len(arrays) - 1, "variants; max", float(img[..., :3].max()))


if __name__ == "__main__":
    main()
