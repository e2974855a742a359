"""Generates tests/golden/resize/cases.npz from the REFERENCE's Image::resize(): the real FreeImage_Rescale compiled
from /root/reference by oracle/Makefile (oracle/_ref/libfiresize.so, see oracle/fi_resize.cpp). Run in the build
container (needs /root/reference): python tests/golden/make_resize_goldens.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import resize as R  # noqa: E402

# (name, src w, src h, dst w, dst h, srgb)
CASES = [("down_odd", 37, 23, 18, 11, False), ("mixed", 16, 16, 40, 9, False), ("y_only", 9, 33, 9, 16, False),
         ("up", 7, 5, 20, 20, False), ("to_1x1", 3, 2, 1, 1, False), ("half", 32, 24, 16, 12, False),
         ("half_srgb", 32, 24, 16, 12, True)]


def source(name, w, h, srgb):
    rng = np.random.default_rng(sum(map(ord, name)) * 7919 + w * 31 + h)
    img = rng.random((h, w, 4), dtype=np.float32)
    if not srgb:
        img[..., :3] = img[..., :3] * 3.0 - 0.5          # HDR and negative values: the float path does not clamp
    return img


def main():
    out = {}
    for (name, sw, sh, dw, dh, srgb) in CASES:
        img = source(name, sw, sh, srgb)
        out[name + "/src"] = img
        for f in R.FILTERS:
            out["%s/%s" % (name, f)] = R.resize_ref(img, dw, dh, f, srgb)
    # a full mip chain (Texture::generateMipmaps): every level from the one above
    img = source("chain", 48, 20, False)
    out["chain/src"] = img
    for k, level in enumerate(R.mip_chain(img, "CatmullRom", fn=R.resize_ref)[1:], 1):
        out["chain/%d" % k] = level
    path = os.path.join(ROOT, "tests", "golden", "resize", "cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
