"""Developer tool (GPU box): FNV-1a of ASTC encodes of fixed inputs (footprints x quality levels, LDR with and without
alpha, HDR, a ragged surface) -- to prove that a kernel change is output-neutral (run before and after, compare)."""
import glob, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuttlefish_b200 as cfx
from cuttlefish_b200 import synth
cfx.init(0)


def fnv(a):
    h = 0xcbf29ce484222325
    for x in np.frombuffer(a.tobytes(), np.uint64).tolist():
        h = ((h ^ x) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "real")
inputs = [("noise+grad 576", synth.to_rgba8(synth.gen_image("noise+grad", 576, 576))), ("ui 288", synth.to_rgba8(synth.gen_image("ui", 288, 288))),
          ("ragged 97x61", synth.to_rgba8(synth.gen_image("noise+grad", 97, 61)))]
for f in ("rgba00", "rgba01", "rgb09", "rgb05"):
    inputs.append((f, np.load(os.path.join(root, f + ".npz"))["src"]))
for name, src in inputs:
    for fmt in ("ASTC_4x4", "ASTC_6x6", "ASTC_8x8", "ASTC_10x8", "ASTC_12x12"):
        qs = ("Lowest", "Normal", "Highest") if fmt == "ASTC_6x6" else ("Normal",)
        print(name, fmt, " ".join("%s=%s" % (q, fnv(cfx.encode(src, fmt, quality=q))) for q in qs), flush=True)
hdr = synth.gen_image("hdr", 192, 192).astype(np.float32)
for fmt in ("ASTC_4x4", "ASTC_6x6"):
    print("hdr", fmt, fnv(cfx.encode(hdr, fmt, type="UFloat")), flush=True)
