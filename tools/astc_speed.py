"""Developer tool (GPU box): device-resident ASTC throughput + PSNR vs oracle for a list of footprints.
    python tools/astc_speed.py ASTC_6x6:4096 ASTC_4x4:2048:High ..."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import cuttlefish_b200 as cfx
import oracle
cfx.init(0)
for spec in sys.argv[1:]:
    parts = spec.split(":")
    fmt, n = parts[0], int(parts[1])
    kw = {"quality": parts[2]} if len(parts) > 2 else {}
    img = oracle.gen_image("noise+grad", n, n)
    src = torch.from_numpy(oracle.to_rgba8(img)).cuda()
    out = torch.empty(cfx.encoded_size(fmt, n, n), dtype=torch.uint8, device="cuda")
    for _ in range(2): cfx.encode_device(src, fmt, out=out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): cfx.encode_device(src, fmt, out=out, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/5
    m = 512
    p = oracle.psnr_rgb(img[:m, :m], oracle.decode(cfx.encode(oracle.to_rgba8(img[:m, :m]), fmt, **kw), fmt, m, m))
    print("%s %s %d^2: %.2f ms %.1f Mtexel/s | psnr(512 crop) %.3f" % (fmt, kw.get("quality", ""), n, ms, n*n/ms/1e3, p), flush=True)
