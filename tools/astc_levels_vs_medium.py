"""Developer tool (GPU box): OUR ASTC PSNR at every quality level against the REFERENCE AT NORMAL (astcenc medium),
on the real crops: tells how much of a gap more exact candidates / refinement rounds would close."""
import glob, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuttlefish_b200 as cfx
import oracle
cfx.init(0)
fmts = sys.argv[1:] or ["ASTC_6x6"]
D = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "real")
for fmt in fmts:
    for path in sorted(glob.glob(os.path.join(D, "*.npz"))):
        name = os.path.basename(path)[:-4]
        z = np.load(path)
        key = "blocks__%s__Normal" % fmt
        if key not in z.files or z["src"].dtype != np.uint8:
            continue
        src = z["src"]; img = src.astype(np.float32) / np.float32(255)
        pr = oracle.psnr_rgb(img, oracle.decode(z[key], fmt, 192, 192))
        line = "%s %-7s ref(medium) %.2f |" % (fmt, name, pr)
        for q in ("Lowest", "Low", "Normal", "High", "Highest"):
            pg = oracle.psnr_rgb(img, oracle.decode(cfx.encode(src, fmt, quality=q), fmt, 192, 192))
            line += " %s %+.2f" % (q, pg - pr)
        print(line, flush=True)
