"""Developer tool (GPU box): a few small encodes of the kernels changed this round, to run under compute-sanitizer.
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuttlefish_b200 as cfx
from cuttlefish_b200 import synth
cfx.init(0)
z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "real", "rgba00.npz"))["src"][:96, :96]
small = synth.to_rgba8(synth.gen_image("noise+grad", 97, 61))
for fmt in ("ASTC_4x4", "ASTC_6x6", "ASTC_8x8", "ASTC_10x8", "ASTC_12x12"):
    for src in (small, z):
        cfx.encode(src, fmt)
hdr = synth.gen_image("hdr", 64, 64).astype(np.float32)
cfx.encode(hdr, "ASTC_6x6", type="UFloat")
for fmt in ("ETC2_R8G8B8", "ETC2_R8G8B8A8", "ETC2_R8G8B8A1", "EAC_R11", "EAC_R11G11", "ETC1"):
    cfx.encode(small, fmt); cfx.encode(small, fmt, quality="High")
for q in ("Lowest", "Low", "Normal", "High", "Highest"):
    cfx.encode(small, "BC1_RGB", quality=q); cfx.encode(z, "BC3", quality=q)
print("sanitize_small done")
