"""Developer tool (GPU box): device-resident ETC1 throughput at every quality level."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import cuttlefish_b200 as cfx
from cuttlefish_b200 import synth
cfx.init(0)
n = 4096
src = torch.from_numpy(synth.to_rgba8(synth.gen_image("noise+grad", n, n))).cuda()
out = torch.empty(cfx.encoded_size("ETC1", n, n), dtype=torch.uint8, device="cuda")
for q in ("Lowest", "Normal", "High", "Highest"):
    for _ in range(2): cfx.encode_device(src, "ETC1", out=out, quality=q)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): cfx.encode_device(src, "ETC1", out=out, quality=q)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/3
    print("ETC1 %s %d^2: %.2f ms %.1f Mtexel/s" % (q, n, ms, n*n/ms/1e3), flush=True)
