"""Developer tool (GPU box): small BC4 / BC5 / BC3 encodes (plain and TMA-staged kernels, UNorm and SNorm, every quality)
to run under compute-sanitizer.
    compute-sanitizer --tool racecheck python tools/sanitize_bc45.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuttlefish_b200 as cfx
from cuttlefish_b200 import synth
cfx.init(0)
ragged = synth.to_rgba8(synth.gen_image("noise+grad", 97, 61))
tiled = synth.to_rgba8(synth.gen_image("noise+grad", 512, 64))      # block rows of 128 blocks: the TMA kernel
for src in (ragged, tiled):
    for fmt in ("BC4", "BC5", "BC3"):
        for q in ("Lowest", "Normal", "High"):
            cfx.encode(src, fmt, quality=q)
    for fmt in ("BC4", "BC5"):
        cfx.encode(src, fmt, type="SNorm")
print("sanitize_bc45 done")
