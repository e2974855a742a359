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


def C(mode, rank=0, variant=0, rounds=2):
    return mode | (rank << 4) | (variant << 8) | (rounds << 12)


DEFAULT_OPAQUE = [C(6), C(6, 0, 1), C(1), C(3), C(1, 0, 1), C(3, 0, 1), C(1, 1), C(3, 1), C(1, 1, 1), C(3, 1, 1),
                  C(1, 2), C(3, 2), C(1, 2, 1), C(3, 2, 1), C(1, 3), C(3, 3)]
DEFAULT_ALPHA = [C(6), C(6, 0, 1)] + [C(7, r, v) for r in range(7) for v in (0, 1)]


def parse_cands(text):
    """'6,6v,1r0,3r0v,1r1x1' -> descriptors: mode, rN rank, v variant, xN rounds"""
    import re
    out = []
    for t in text.split(","):
        m = re.fullmatch(r"(\d)(?:r(\d+))?(v)?(c)?(d)?(?:x(\d))?", t.strip())
        out.append(C(int(m.group(1)), int(m.group(2) or 0), (1 if m.group(3) else 0) | (2 if m.group(4) else 0) |
                     (4 if m.group(5) else 0), int(m.group(6) or 2)))
    return out


def bc7_encode(lib, src, cands=None, cands_alpha=None, mask=15):
    cands = cands or DEFAULT_OPAQUE
    cands_alpha = cands_alpha or DEFAULT_ALPHA[:len(cands)]
    while len(cands_alpha) < len(cands):
        cands_alpha.append(cands_alpha[-1])
    G = len(cands)
    co = np.array(cands, np.uint16)
    ca = np.array(cands_alpha[:G], np.uint16)
    h, w, _ = src.shape
    nb = ((w + 3) // 4) * ((h + 3) // 4)
    out = np.empty(nb * 16, np.uint8)
    dbg = np.empty((nb, 4), np.uint32)
    lib.emu_bc7_encode(src.ctypes.data_as(ctypes.c_void_p), w, h, out.ctypes.data_as(ctypes.c_void_p), G, mask,
                       dbg.ctypes.data_as(ctypes.c_void_p), co.ctypes.data_as(ctypes.c_void_p),
                       ca.ctypes.data_as(ctypes.c_void_p))
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
    ap.add_argument("--cands", default="")
    ap.add_argument("--cands-alpha", default="")
    ap.add_argument("--show", type=int, default=0)
    a = ap.parse_args()
    lib = build(a.codec)
    cands = parse_cands(a.cands) if a.cands else None
    cands_alpha = parse_cands(a.cands_alpha) if a.cands_alpha else None
    for kind in a.kind.split(","):
        n = a.size
        if os.path.exists(kind):
            from PIL import Image
            src = np.ascontiguousarray(np.array(Image.open(kind).convert("RGBA")))
            img = src.astype(np.float32) / np.float32(255)
            kind = os.path.basename(kind)
        else:
            img = oracle.gen_image(kind, n, n)
            src = oracle.to_rgba8(img)
        hh, ww = src.shape[:2]
        got, dbg = bc7_encode(lib, src, cands, cands_alpha)
        ref = oracle.encode(img, "BC7")
        dg, dr = oracle.decode(got, "BC7", ww, hh), oracle.decode(ref, "BC7", ww, hh)
        pg, pr = oracle.psnr_rgb(img, dg), oracle.psnr_rgb(img, dr)
        eg, er = block_sse(src, dg), block_sse(src, dr)
        eg4, er4 = block_sse(src, dg, channels=4), block_sse(src, dr, channels=4)
        p4 = lambda e: 10 * np.log10(255.0 ** 2 * ww * hh * 4 / max(e.sum(), 1))
        print("%s %dx%d: emu %.3f dB ref %.3f dB delta %+.3f | rgba psnr emu %.3f ref %.3f delta %+.3f | blocks worse %d better %d equal %d" % (
            kind, ww, hh, pg, pr, pg - pr, p4(eg4), p4(er4), p4(eg4) - p4(er4), (eg > er).sum(), (eg < er).sum(), (eg == er).sum()))
        modes = np.bincount(dbg[:, 0], minlength=8)
        print("   modes chosen:", {m: int(c) for m, c in enumerate(modes) if c})
        refmode = np.array([int(np.log2(b & -b)) if b else 8 for b in ref.reshape(-1, 16)[:, 0]])
        print("   ref modes:", {m: int(c) for m, c in enumerate(np.bincount(refmode, minlength=9)) if c})
        if a.show:
            worst = np.argsort(eg - er)[::-1][:a.show]
            bxn = (ww + 3) // 4
            for b in worst:
                by, bx = divmod(int(b), bxn)
                print("   block", b, "emu sse", eg[b], "ref sse", er[b], "mode/shape/var/err", dbg[b], "refmode", refmode[b])
                print(src[by * 4:by * 4 + 4, bx * 4:bx * 4 + 4, :3].reshape(16, 3).T)


if __name__ == "__main__":
    main()
