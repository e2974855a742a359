"""Developer tool (GPU box): ETC2 / EAC partial edge blocks -- mean squared error over the visible texels of the blocks
that overhang the image, ours against the reference (which hands etc2comp a smaller image for them)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuttlefish_b200 as cfx
import oracle
cfx.init(0)
for fmt, q in (("ETC2_R8G8B8", "Normal"), ("ETC2_R8G8B8A8", "Normal"), ("ETC2_R8G8B8", "High"), ("ETC1", "High"), ("ETC2_R8G8B8A1", "Normal")):
    for w, h in ((97, 61), (130, 67), (33, 18), (13, 7)):
        img = oracle.gen_image("noise+grad", w, h, seed=w*100 + h)
        if fmt.endswith("A8"):
            img[..., 3] = np.clip(oracle.gen_image("gradient", w, h)[..., 0]*1.3, 0, 1)
        src = oracle.to_rgba8(img)
        got = cfx.encode(src, fmt, quality=q)
        ref = oracle.encode(img, fmt, quality=q)
        dg, dr = oracle.decode(got, fmt, w, h), oracle.decode(ref, fmt, w, h)
        edge = np.zeros((h, w), bool)
        if w % 4: edge[:, w - w % 4:] = True
        if h % 4: edge[h - h % 4:, :] = True
        nch = 4 if fmt.endswith("A8") else 3
        e = lambda d, m: float(np.mean((d[..., :nch].astype(np.float64)[m] - img[..., :nch][m]) ** 2))
        print("%s %s %dx%d edge mse ours %.4g ref %.4g ratio %.3f | interior ratio %.3f" % (fmt, q, w, h, e(dg, edge), e(dr, edge),
            e(dg, edge)/max(e(dr, edge), 1e-12), e(dg, ~edge)/max(e(dr, ~edge), 1e-12)), flush=True)
