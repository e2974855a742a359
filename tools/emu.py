"""Developer tool: drive the host emulation of an encoder core and compare with the CPU oracle.
    python tools/emu.py bc7 [--kind gradient] [--size 256]
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


def build(name):
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    so = os.path.join(HERE, "_build", "libemu_%s.so" % name)
    src = os.path.join(HERE, "emu_%s.cpp" % name)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=fast", "-mfma", "-o", so, src])
    return ctypes.CDLL(so)


def bc7_encode(lib, src, G=16, mask=15):
    h, w, _ = src.shape
    nb = ((w + 3) // 4) * ((h + 3) // 4)
    out = np.empty(nb * 16, np.uint8)
    dbg = np.empty((nb, 4), np.uint32)
    lib.emu_bc7_encode(src.ctypes.data_as(ctypes.c_void_p), w, h, out.ctypes.data_as(ctypes.c_void_p), G, mask,
                       dbg.ctypes.data_as(ctypes.c_void_p))
    return out, dbg


def block_sse(img8, dec, bw=4, bh=4, channels=3):
    h, w, _ = img8.shape
    d = (img8[..., :channels].astype(np.int64) - np.rint(dec[..., :channels] * 255).astype(np.int64)) ** 2
    d = d.sum(axis=2)
    H, W = (h + bh - 1) // bh * bh, (w + bw - 1) // bw * bw
    pad = np.zeros((H, W), np.int64)
    pad[:h, :w] = d
    return pad.reshape(H // bh, bh, W // bw, bw).sum(axis=(1, 3)).ravel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("codec")
    ap.add_argument("--kind", default="gradient,noise+grad")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("-G", type=int, default=16)
    ap.add_argument("--show", type=int, default=0)
    a = ap.parse_args()
    lib = build(a.codec)
    for kind in a.kind.split(","):
        n = a.size
        img = oracle.gen_image(kind, n, n)
        src = oracle.to_rgba8(img)
        got, dbg = bc7_encode(lib, src, a.G)
        ref = oracle.encode(img, "BC7")
        dg, dr = oracle.decode(got, "BC7", n, n), oracle.decode(ref, "BC7", n, n)
        pg, pr = oracle.psnr_rgb(img, dg), oracle.psnr_rgb(img, dr)
        eg, er = block_sse(src, dg), block_sse(src, dr)
        eg4, er4 = block_sse(src, dg, channels=4), block_sse(src, dr, channels=4)
        print("%s %d: emu %.3f dB ref %.3f dB delta %+.3f | rgba sse emu %d ref %d | blocks worse %d better %d equal %d" % (
            kind, n, pg, pr, pg - pr, eg4.sum(), er4.sum(), (eg > er).sum(), (eg < er).sum(), (eg == er).sum()))
        modes = np.bincount(dbg[:, 0], minlength=8)
        print("   modes chosen:", {m: int(c) for m, c in enumerate(modes) if c})
        refmode = np.array([int(np.log2(b & -b)) if b else 8 for b in ref.reshape(-1, 16)[:, 0]])
        print("   ref modes:", {m: int(c) for m, c in enumerate(np.bincount(refmode, minlength=9)) if c})
        if a.show:
            worst = np.argsort(eg - er)[::-1][:a.show]
            bxn = (n + 3) // 4
            for b in worst:
                by, bx = divmod(int(b), bxn)
                print("   block", b, "emu sse", eg[b], "ref sse", er[b], "mode/shape/var/err", dbg[b], "refmode", refmode[b])
                print(src[by * 4:by * 4 + 4, bx * 4:bx * 4 + 4, :3].reshape(16, 3).T)


if __name__ == "__main__":
    main()
