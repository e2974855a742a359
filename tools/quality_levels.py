"""Developer tool (GPU box): PSNR of every Texture::Quality level against the CPU oracle at the same level.
    python tools/quality_levels.py BC7 ASTC_6x6 ..."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuttlefish_b200 as cfx
import oracle
cfx.init(0)
n = 192
for fmt in sys.argv[1:]:
    for kind in ("noise+grad", "ui"):
        img = oracle.gen_image(kind, n, n)
        src = oracle.to_rgba8(img)
        line = "%s %s:" % (fmt, kind)
        for q in ("Lowest", "Low", "Normal", "High", "Highest"):
            got = cfx.encode(src, fmt, quality=q)
            t = time.time(); ref = oracle.encode(img, fmt, quality=q); dt = time.time() - t
            pg = oracle.psnr_rgb(img, oracle.decode(got, fmt, n, n)); pr = oracle.psnr_rgb(img, oracle.decode(ref, fmt, n, n))
            line += " %s %+.2f (%.2f, cpu %.1fs)" % (q, pg - pr, pg, dt)
        print(line, flush=True)
