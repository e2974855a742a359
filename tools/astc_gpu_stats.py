"""Developer tool (GPU box): what OUR ASTC encoder chose (partition counts, end point modes, dual plane) on the bench
generator and on the real crops, next to the reference's choice for the crops.
    python tools/astc_gpu_stats.py [ASTC_6x6]"""
import glob, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cuttlefish_b200 as cfx
from cuttlefish_b200 import synth
from astc_stats import stats
fmt = sys.argv[1] if len(sys.argv) > 1 else "ASTC_6x6"
cfx.init(0)
for kind in ("noise+grad", "ui", "gradient"):
    img = synth.gen_image(kind, 576, 576)
    c, pcs, cems, duals = stats(cfx.encode(synth.to_rgba8(img), fmt))
    print(kind, "partitions", dict(pcs), "cems", dict(cems), "dual", dict(duals))
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "real")
for f in sorted(glob.glob(os.path.join(root, "*.npz"))):
    d = np.load(f)
    key = "blocks__%s__Normal" % fmt
    if key not in d.files or "hdr" in f:
        continue
    src = d["src"]
    if src.dtype != np.uint8:
        src = synth.to_rgba8(src)
    c, pcs, cems, duals = stats(cfx.encode(src, fmt))
    print(os.path.basename(f), "ours partitions", dict(pcs), "cems", dict(cems), "dual", dict(duals))
    c, pcs, cems, duals = stats(d[key])
    print(os.path.basename(f), "ref  partitions", dict(pcs), "cems", dict(cems), "dual", dict(duals))
