import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import cuttlefish_b200 as cfx
from cuttlefish_b200 import synth
cfx.init(0)
n = 4096
src = torch.from_numpy(synth.to_rgba8(synth.gen_image("noise+grad", n, n))).cuda()
for fmt in ("BC1_RGB", "BC3"):
    out = torch.empty(cfx.encoded_size(fmt, n, n), dtype=torch.uint8, device="cuda")
    for q in ("Lowest", "Low", "Normal", "High", "Highest"):
        for _ in range(2): cfx.encode_device(src, fmt, out=out, quality=q)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): cfx.encode_device(src, fmt, out=out, quality=q)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/3
        print("%s %s %d^2: %.2f ms %.1f Mtexel/s" % (fmt, q, n, ms, n*n/ms/1e3), flush=True)
