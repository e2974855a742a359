"""Developer tool (GPU box): where does our ASTC encoder lose to the reference?  Per-block SSE of both on the
screenshot-like probe image, with the partition count / end point mode / block mode each side chose.
    python tools/astc_block_diff.py ASTC_8x8 [size]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuttlefish_b200 as cfx
import oracle

fmt = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 144
bw, bh, _ = cfx.block_info(fmt)
n = n // (bw * bh) * (bw * bh) if (n % bw or n % bh) else n
img = oracle.gen_image("ui", n, n)
cfx.init(0)
got = cfx.encode(oracle.to_rgba8(img), fmt)
ref = oracle.encode(img, fmt)
dg, dr = oracle.decode(got, fmt, n, n), oracle.decode(ref, fmt, n, n)
nbx, nby = n // bw, n // bh


def blk_sse(d):
    return (((d[..., :3] - img[..., :3]) * 255) ** 2).reshape(nby, bh, nbx, bw, 3).sum(axis=(1, 3, 4))


def info(b):
    v = int.from_bytes(bytes(b), "little")
    mode = v & 0x7FF
    if (mode & 0x1FF) == 0x1FC:
        return "void"
    pc = ((v >> 11) & 3) + 1
    if pc == 1:
        return "pc1 cem%d mode%03x" % ((v >> 13) & 0xF, mode)
    return "pc%d cemfield%02x mode%03x" % (pc, (v >> 23) & 0x3F, mode)


eg, er = blk_sse(dg), blk_sse(dr)
print("%s ui %dx%d: ours %.3f dB ref %.3f dB | total sse ours %.0f ref %.0f" % (
    fmt, n, n, oracle.psnr_rgb(img, dg), oracle.psnr_rgb(img, dr), eg.sum(), er.sum()))
for name, blocks in (("ref", ref), ("ours", got)):
    cnt = {}
    for b in blocks.reshape(-1, 16):
        k = info(b).split(" mode")[0]
        cnt[k] = cnt.get(k, 0) + 1
    print(" %s kinds: %s" % (name, dict(sorted(cnt.items(), key=lambda kv: -kv[1]))))
d = eg - er
# excess grouped by the reference's choice
grp = {}
for o in range(nbx * nby):
    k = info(ref.reshape(-1, 16)[o]).split(" mode")[0]
    grp[k] = grp.get(k, 0.0) + float(d.ravel()[o])
print(" excess SSE by the reference's block kind:", {k: int(v) for k, v in sorted(grp.items(), key=lambda kv: -kv[1])})
for o in np.argsort(-d.ravel())[:12]:
    by, bx = divmod(int(o), nbx)
    print("  blk(%d,%d) ours %.0f ref %.0f | ref: %s | ours: %s" % (by, bx, eg[by, bx], er[by, bx],
          info(ref.reshape(-1, 16)[o]), info(got.reshape(-1, 16)[o])))
