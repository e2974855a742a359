"""Developer tool: drive the host emulation of the ASTC core and compare with the CPU oracle.
    python tools/emu_astc.py [--kind gradient,noise+grad] [--size 256] [--fmt ASTC_6x6]
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
    so = os.path.join(HERE, "_build", "libemu_astc.so")
    extra = os.environ.get("EMU_CXXFLAGS", "").split()
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=fast", "-mfma", "-o", so,
                           os.path.join(HERE, "emu_astc.cpp")] + extra)
    return ctypes.CDLL(so)


def encode(lib, img, bw, bh, slots=9, refine=2, quality=2):
    h, w, _ = img.shape
    nb = ((w + bw - 1) // bw) * ((h + bh - 1) // bh)
    out = np.zeros(nb * 16, np.uint8)
    dbg = np.zeros((nb, 4), np.uint32)
    img = np.ascontiguousarray(img, np.float32)
    lib.emu_astc_encode(img.ctypes.data_as(ctypes.c_void_p), w, h, bw, bh, out.ctypes.data_as(ctypes.c_void_p),
                        slots, refine, quality, dbg.ctypes.data_as(ctypes.c_void_p))
    return out, dbg


def load(kind, n):
    if os.path.exists(kind):
        from PIL import Image
        src = np.ascontiguousarray(np.array(Image.open(kind).convert("RGBA")))
        return src.astype(np.float32) / np.float32(255), os.path.basename(kind)
    return oracle.gen_image(kind, n, n), kind


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="gradient,noise+grad")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--fmt", default="ASTC_6x6")
    ap.add_argument("--slots", type=int, default=9)
    ap.add_argument("--refine", type=int, default=2)
    ap.add_argument("--quality", type=int, default=2)
    ap.add_argument("--modes", action="store_true")
    a = ap.parse_args()
    lib = build()
    bw, bh = [int(x) for x in a.fmt.split("_")[1].split("x")]
    for kind in a.kind.split(","):
        img, name = load(kind, a.size)
        h, w, _ = img.shape
        got, dbg = encode(lib, img, bw, bh, a.slots, a.refine, a.quality)
        ref = oracle.encode(img, a.fmt)
        dg, dr = oracle.decode(got, a.fmt, w, h), oracle.decode(ref, a.fmt, w, h)
        pg, pr = oracle.psnr_rgb(img, dg), oracle.psnr_rgb(img, dr)
        # self-check: error predicted by the encoder vs error of the oracle-decoded block
        pred = float(dbg[:, 2].sum())
        act = float((((dg - img) * 255.0) ** 2).sum())
        print("%s %dx%d %s: emu %.3f dB ref %.3f dB delta %+.3f | predicted SSE %.0f decoded SSE %.0f" % (
            name, w, h, a.fmt, pg, pr, pg - pr, pred, act))
        print("   slots chosen:", {int(k): int(v) for k, v in zip(*np.unique(dbg[:, 0], return_counts=True))},
              "partitions:", {int(k): int(v) for k, v in zip(*np.unique(dbg[:, 3], return_counts=True))})
        if a.modes:
            ids, cnt = np.unique(dbg[dbg[:, 0] != 9, 1], return_counts=True)
            info = (ctypes.c_uint32 * 4)()
            for i, c in sorted(zip(ids, cnt), key=lambda t: -t[1])[:24]:
                lib.emu_astc_mode_info(bw, bh, int(i), info)
                print("     mode %3d: grid %dx%d levels %2d wbits %2d  x%d" % (i, info[0], info[1], info[2], info[3], c))


if __name__ == "__main__":
    main()
