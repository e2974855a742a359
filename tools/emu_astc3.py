"""Developer tool: drive tools/emu_astc3.cpp (two-phase ASTC search experiment) and compare with the CPU oracle.
    python tools/emu_astc3.py [--kind noise+grad,gradient,/path/to.png] [--size 192] [--fmt ASTC_6x6] [--all-modes]
"""
import argparse
import ctypes
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402


def build():
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    so = os.path.join(HERE, "_build", "libemu_astc3.so")
    src = os.path.join(HERE, "emu_astc3.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-march=native", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)


def load(kind, n):
    if os.path.exists(kind):
        from PIL import Image
        src = np.ascontiguousarray(np.array(Image.open(kind).convert("RGBA")))
        if n and (src.shape[0] > n or src.shape[1] > n):
            y0, x0 = max(0, (src.shape[0] - n) // 2), max(0, (src.shape[1] - n) // 2)
            src = np.ascontiguousarray(src[y0:y0 + n, x0:x0 + n])
        return src.astype(np.float32) / np.float32(255), os.path.basename(kind)
    return oracle.gen_image(kind, n, n), kind


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="noise+grad,gradient")
    ap.add_argument("--size", type=int, default=192)
    ap.add_argument("--fmt", default="ASTC_6x6")
    ap.add_argument("--ns", default="1,2,3,4,6,8,12,16,100000")
    ap.add_argument("--refine", type=int, default=2)
    ap.add_argument("--quality", type=int, default=2)
    ap.add_argument("--all-modes", action="store_true")
    ap.add_argument("--pick", type=int, default=5)
    a = ap.parse_args()
    lib = build()
    bw, bh = [int(x) for x in a.fmt.split("_")[1].split("x")]
    ns = np.array([int(x) for x in a.ns.split(",")], np.uint32)
    for kind in a.kind.split(","):
        img, name = load(kind, a.size)
        h, w, _ = img.shape
        img = np.ascontiguousarray(img, np.float32)
        nb = ((w + bw - 1) // bw) * ((h + bh - 1) // bh)
        sse = np.zeros(len(ns), np.float64)
        hist = np.zeros(64, np.float64)
        out = np.zeros(nb * 16, np.uint8)
        lib.emu3_run(img.ctypes.data_as(ctypes.c_void_p), w, h, bw, bh, ns.ctypes.data_as(ctypes.c_void_p), len(ns),
                     sse.ctypes.data_as(ctypes.c_void_p), hist.ctypes.data_as(ctypes.c_void_p),
                     out.ctypes.data_as(ctypes.c_void_p), a.pick, a.refine, a.quality, 1 if a.all_modes else 0)
        ref = oracle.encode(img, a.fmt)
        dr = oracle.decode(ref, a.fmt, w, h)
        pr = oracle.psnr_rgb(img, dr)
        has_alpha = bool((img[..., 3] != 1).any())
        nch = 4 if has_alpha else 3
        # the encoder's own SSE covers edge-clamped texels too; close enough for ranking N
        ps = 10 * np.log10(255.0 ** 2 * (nb * bw * bh * nch) / np.maximum(sse, 1e-9))
        dg = oracle.decode(out, a.fmt, w, h)
        pg = oracle.psnr_rgb(img, dg)
        print("%s %dx%d %s: ref %.3f dB | decoded (N=%d) %.3f dB (delta %+.3f)" % (name, w, h, a.fmt, pr, ns[a.pick], pg, pg - pr))
        print("   N:      " + " ".join("%7d" % n for n in ns))
        print("   psnr*:  " + " ".join("%7.3f" % p for p in ps))
        print("   d(all): " + " ".join("%+7.3f" % (p - ps[-1]) for p in ps))
        c = np.cumsum(hist) / max(hist.sum(), 1)
        print("   exact winner within top-k by estimate: " + " ".join("k=%d:%.2f" % (k, c[k - 1]) for k in (1, 2, 4, 8, 16, 32, 63)))


if __name__ == "__main__":
    main()
