"""Developer tool: histogram of what an ASTC encoder chose (block modes, partitions, CEMs)."""
import collections
import sys
import numpy as np


def decode_mode(b):
    if (b & 3) != 0:
        R = ((b >> 4) & 1) | ((b & 3) << 1)
        A, B = (b >> 5) & 3, (b >> 7) & 3
        D, H = (b >> 10) & 1, (b >> 9) & 1
        k = (b >> 2) & 3
        if k == 0: W, Hh = B + 4, A + 2
        elif k == 1: W, Hh = B + 8, A + 2
        elif k == 2: W, Hh = A + 2, B + 8
        elif B & 2: W, Hh = (B & 1) + 2, A + 2
        else: W, Hh = A + 2, (B & 1) + 6
    else:
        if ((b >> 2) & 3) == 0: return None
        R = ((b >> 4) & 1) | (((b >> 2) & 3) << 1)
        A = (b >> 5) & 3
        D, H = (b >> 10) & 1, (b >> 9) & 1
        k = (b >> 7) & 3
        if k == 0: W, Hh = 12, A + 2
        elif k == 1: W, Hh = A + 2, 12
        elif k == 3:
            if (b >> 5) & 2: return None
            W, Hh = (10, 6) if (b >> 5) & 1 else (6, 10)
        else: W, Hh, D, H = A + 6, ((b >> 9) & 3) + 6, 0, 0
    if R < 2: return None
    n = [2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32][(R - 2) + 6 * H]
    return W, Hh, n, D


def stats(blocks):
    blocks = np.asarray(blocks, np.uint8).reshape(-1, 16)
    c = collections.Counter()
    pcs = collections.Counter(); cems = collections.Counter(); duals = collections.Counter()
    for blk in blocks:
        v = int.from_bytes(blk.tobytes(), "little")
        if (v & 0x1FF) == 0x1FC:
            c["void"] += 1; continue
        m = decode_mode(v & 0x7FF)
        pc = ((v >> 11) & 3) + 1
        if pc == 1:
            cem = (v >> 13) & 15
        else:
            cf = (v >> 23) & 0x3F
            cem = (cf >> 2) & 15 if (cf & 3) == 0 else "multi"
        pcs[pc] += 1; cems[cem] += 1; duals[m[3] if m else -1] += 1
        c[m] += 1
    return c, pcs, cems, duals


if __name__ == "__main__":
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    import oracle
    kind, n = sys.argv[1], int(sys.argv[2])
    fmt = sys.argv[3] if len(sys.argv) > 3 else "ASTC_6x6"
    if os.path.exists(kind):
        from PIL import Image
        img = np.array(Image.open(kind).convert("RGBA")).astype(np.float32) / np.float32(255)
    else:
        img = oracle.gen_image(kind, n, n)
    c, pcs, cems, duals = stats(oracle.encode(img, fmt))
    print("partitions", dict(pcs), "cems", dict(cems), "dual", dict(duals))
    for m, k in c.most_common(16):
        print("  ", m, k)
