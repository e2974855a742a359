"""Developer tool: PSNR delta (emulated encoder - CPU oracle) of several BC7 candidate sets."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import emu  # noqa: E402
import oracle  # noqa: E402
from PIL import Image  # noqa: E402

R = "/root/reference/lib/astc-encoder/Test/Images/Small/"
IMAGES = [("noise", "noise+grad", 256), ("g256", "gradient", 256), ("g512", "gradient", 512), ("g1024", "gradient", 1024),
          ("rgb00", R + "LDR-RGB/ldr-rgb-00.png", 0), ("rgb03", R + "LDR-RGB/ldr-rgb-03.png", 0),
          ("rgb07", R + "LDR-RGB/ldr-rgb-07.png", 0), ("rgba00", R + "LDR-RGBA/ldr-rgba-00.png", 0),
          ("rgba02", R + "LDR-RGBA/ldr-rgba-02.png", 0),
          ("ruby", "/root/reference/lib/compressonator/runtime/images/ruby.png", 0)]


def load(kind, n):
    if os.path.exists(kind):
        src = np.ascontiguousarray(np.array(Image.open(kind).convert("RGBA")))
        return src.astype(np.float32) / np.float32(255), src
    img = oracle.gen_image(kind, n, n)
    return img, oracle.to_rgba8(img)


def psnr4(src, dec):
    d = src.astype(np.float64) - np.rint(dec * 255)
    return 10 * np.log10(255.0 ** 2 / max(np.mean(d * d), 1e-12))


def main():
    lib = emu.build("bc7")
    sets = sys.argv[1:]
    data = []
    for name, kind, n in IMAGES:
        img, src = load(kind, n)
        h, w = src.shape[:2]
        ref = oracle.encode(img, "BC7")
        dr = oracle.decode(ref, "BC7", w, h)
        data.append((name, img, src, oracle.psnr_rgb(img, dr), psnr4(src, dr)))
    print("%-44s" % "set" + "".join("%14s" % d[0] for d in data))
    print("%-44s" % "ref rgb/rgba dB" + "".join("%7.2f/%6.2f" % (d[3], d[4]) for d in data))
    for spec in sets:
        if "/" in spec:
            so, sa = spec.split("/")
        else:
            so, sa = spec, ""
        co = emu.parse_cands(so)
        ca = emu.parse_cands(sa) if sa else None
        line = "%-44s" % spec[:43]
        for name, img, src, pr, pr4 in data:
            h, w = src.shape[:2]
            got, _ = emu.bc7_encode(lib, src, co, ca)
            dg = oracle.decode(got, "BC7", w, h)
            line += "%+7.2f/%+6.2f" % (oracle.psnr_rgb(img, dg) - pr, psnr4(src, dg) - pr4)
        print(line, flush=True)


if __name__ == "__main__":
    main()
