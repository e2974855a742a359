"""Developer tool (CPU): where do OUR ASTC blocks (gpurun_out/real_blocks.npz) lose against the reference's on the real
crops?  Per-block SSE of both, grouped by what the reference chose for the block."""
import collections, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import oracle
from astc_stats import decode_mode


def info(blk):
    v = int.from_bytes(blk.tobytes(), "little")
    if (v & 0x1FF) == 0x1FC:
        return ("void",)
    m = decode_mode(v & 0x7FF)
    pc = ((v >> 11) & 3) + 1
    if pc == 1:
        cem = (v >> 13) & 15
    else:
        cf = (v >> 23) & 0x3F
        cem = (cf >> 2) & 15 if (cf & 3) == 0 else "mix"
    return (pc, cem, m[3] if m else -1, "%dx%d" % (m[0], m[1]) if m else "?", m[2] if m else 0)


def main():
    name, fmt = sys.argv[1], sys.argv[2]
    q = sys.argv[3] if len(sys.argv) > 3 else "Normal"
    bw, bh = [int(x) for x in fmt.split("_")[1].split("x")]
    z = np.load(os.path.join(HERE, "..", "tests", "golden", "real", name + ".npz"))
    ours = np.load(os.path.join(HERE, "..", "gpurun_out", "real_blocks.npz"))["%s__%s__%s" % (name, fmt, q)]
    ref = z["blocks__%s__%s" % (fmt, q)]
    src = z["src"]; img = src.astype(np.float32) / np.float32(255)
    h, w, _ = src.shape
    dg, dr = oracle.decode(ours, fmt, w, h), oracle.decode(ref, fmt, w, h)
    nch = 4 if (src[..., 3] != 255).any() else 3

    def bsse(d):
        e = ((d[..., :nch].astype(np.float64) - img[..., :nch]) ** 2).sum(axis=2) * 65025
        H, W = (h + bh - 1) // bh * bh, (w + bw - 1) // bw * bw
        pad = np.zeros((H, W)); pad[:h, :w] = e
        return pad.reshape(H // bh, bh, W // bw, bw).sum(axis=(1, 3)).ravel()
    eg, er = bsse(dg), bsse(dr)
    print("%s %s %s: total SSE ours %.0f ref %.0f  (%.3f dB)" % (name, fmt, q, eg.sum(), er.sum(), 10 * np.log10(er.sum() / eg.sum())))
    rb, ob = ref.reshape(-1, 16), ours.reshape(-1, 16)
    for label, key in (("pc", 0), ("cem", 1), ("dual", 2), ("grid", 3), ("wlevels", 4)):
        groups = collections.defaultdict(lambda: [0, 0.0, 0.0])
        for i in range(rb.shape[0]):
            inf = info(rb[i])
            k = inf[key] if len(inf) > key else "void"
            g = groups[k]; g[0] += 1; g[1] += eg[i]; g[2] += er[i]
        print(" by reference's %s:" % label)
        for k, (n, a, b) in sorted(groups.items(), key=lambda kv: -(kv[1][1] - kv[1][2])):
            print("    %-8s n=%4d  ours %9.0f  ref %9.0f  excess %+9.0f" % (k, n, a, b, a - b))
    # what WE chose on the blocks where we lose most
    lose = np.argsort(er - eg)[:12]
    print(" worst blocks (ours vs ref):")
    for i in lose:
        print("    blk %4d ours %7.0f %s | ref %7.0f %s" % (i, eg[i], info(ob[i]), er[i], info(rb[i])))
    ours_groups = collections.Counter(info(b)[:3] for b in ob)
    ref_groups = collections.Counter(info(b)[:3] for b in rb)
    print(" (pc, cem, dual) ours:", dict(ours_groups))
    print(" (pc, cem, dual) ref: ", dict(ref_groups))


if __name__ == "__main__":
    main()
