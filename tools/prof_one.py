"""Developer tool (GPU box): a few device-resident encodes of one format, for ncu to capture.
    ncu --set full -k regex:bc7 -s 1 -c 1 -o gpurun_out/prof python tools/prof_one.py BC7 4096
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import cuttlefish_b200 as cfx  # noqa: E402
from cuttlefish_b200 import synth  # noqa: E402

fmt = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
kind = sys.argv[3] if len(sys.argv) > 3 else "noise+grad"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
img = synth.gen_image(kind, n, n)
kw = {}
if kind == "hdr":
    src = torch.from_numpy(img.astype(np.float16)).cuda()
    kw["type"] = "UFloat"
else:
    src = torch.from_numpy(synth.to_rgba8(img)).cuda()
cfx.init(0)
out = torch.empty(cfx.encoded_size(fmt, n, n), dtype=torch.uint8, device="cuda")
for _ in range(reps):
    cfx.encode_device(src, fmt, out=out, **kw)
torch.cuda.synchronize()
print("done", fmt, n, kind)
