"""Developer tool (GPU box): PSNR delta (GPU - reference) for every committed real-image golden, as a table."""
import glob, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import cuttlefish_b200 as cfx
import oracle
from util import decode_any
cfx.init(0)
only = sys.argv[1:]
D = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "real")
rows = {}
for path in sorted(glob.glob(os.path.join(D, "*.npz"))):
    name = os.path.basename(path)[:-4]
    z = np.load(path)
    src = z["src"]
    hdr = src.dtype == np.uint16
    if hdr:
        src = src.view(np.float16)
    img = src.astype(np.float32) if hdr else src.astype(np.float32) / np.float32(255)
    h, w, _ = src.shape
    for key in z.files:
        if not key.startswith("blocks__"):
            continue
        _, fmt, q = key.split("__")
        if only and not any(fmt.startswith(o) for o in only):
            continue
        kw = dict(quality=q)
        dkw = {}
        if hdr:
            kw["type"] = "UFloat"; dkw["type"] = "UFloat"
        got = cfx.encode(src, fmt, **kw)
        ref = z[key]
        if np.array_equal(got, ref):
            rows.setdefault((fmt, q), []).append((name, "exact"))
            continue
        dg, dr = decode_any(oracle, got, fmt, w, h, dkw), decode_any(oracle, ref, fmt, w, h, dkw)
        ps = lambda d, n: 10 * np.log10(1.0 / max(float(np.mean((d[..., :n].astype(np.float64) - img[..., :n]) ** 2)), 1e-12))
        rows.setdefault((fmt, q), []).append((name, "%+.2f/%+.2f" % (ps(dg, 3) - ps(dr, 3), ps(dg, 4) - ps(dr, 4))))
for (fmt, q), vals in sorted(rows.items()):
    print("%-14s %-8s " % (fmt, q) + "  ".join("%s %s" % v for v in vals))
