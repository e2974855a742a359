"""Developer tool (GPU box): FNV-1a of BC7 encodes of fixed inputs at every quality -- to prove that a kernel change is
output-neutral (run before and after, compare the lines)."""
import os, sys
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


z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "real", "rgba00.npz"))
inputs = [("noise+grad 1024", synth.to_rgba8(synth.gen_image("noise+grad", 1024, 1024))), ("gradient 512", synth.to_rgba8(synth.gen_image("gradient", 512, 512))),
          ("ui 384", synth.to_rgba8(synth.gen_image("ui", 384, 384))), ("rgba00", z["src"]), ("ragged 97x61", synth.to_rgba8(synth.gen_image("noise+grad", 97, 61)))]
for name, src in inputs:
    print(name, " ".join("%s=%s" % (q, fnv(cfx.encode(src, "BC7", quality=q))) for q in ("Lowest", "Normal", "High", "Highest")), flush=True)
