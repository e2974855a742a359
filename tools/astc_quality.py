"""Developer tool (GPU box): ASTC PSNR vs the CPU oracle on the probe images + device-resident speed.
    python tools/astc_quality.py ASTC_6x6 ASTC_8x8 ..."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import cuttlefish_b200 as cfx
import oracle
cfx.init(0)
for fmt in sys.argv[1:]:
    line = fmt + ":"
    for kind, n in (("ui", 288), ("noise+grad", 288), ("gradient", 288)):
        img = oracle.gen_image(kind, n, n)
        got = cfx.encode(oracle.to_rgba8(img), fmt)
        ref = oracle.encode(img, fmt)
        pg = oracle.psnr_rgb(img, oracle.decode(got, fmt, n, n)); pr = oracle.psnr_rgb(img, oracle.decode(ref, fmt, n, n))
        line += " %s %+.3f (%.2f)" % (kind, pg - pr, pg)
    n = 2048
    img = oracle.gen_image("noise+grad", n, n)
    src = torch.from_numpy(oracle.to_rgba8(img)).cuda()
    out = torch.empty(cfx.encoded_size(fmt, n, n), dtype=torch.uint8, device="cuda")
    for _ in range(2): cfx.encode_device(src, fmt, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): cfx.encode_device(src, fmt, out=out)
    e1.record(); torch.cuda.synchronize()
    print(line + " | %.0f Mtexel/s" % (n*n/(e0.elapsed_time(e1)/5)/1e3), flush=True)
