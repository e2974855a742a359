"""Developer tool (GPU box): encode every committed real-image golden and save OUR blocks to gpurun_out/real_blocks.npz
(key = image__format__quality), for block-level analysis on the CPU side (tools/real_blockdiff.py)."""
import glob, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuttlefish_b200 as cfx
cfx.init(0)
only = sys.argv[1:]
D = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "real")
out = {}
for path in sorted(glob.glob(os.path.join(D, "*.npz"))):
    name = os.path.basename(path)[:-4]
    z = np.load(path)
    src = z["src"]
    hdr = src.dtype == np.uint16
    if hdr:
        src = src.view(np.float16)
    for key in z.files:
        if not key.startswith("blocks__"):
            continue
        _, fmt, q = key.split("__")
        if only and not any(fmt.startswith(o) for o in only):
            continue
        kw = dict(quality=q)
        if hdr:
            kw["type"] = "UFloat"
        out["%s__%s__%s" % (name, fmt, q)] = cfx.encode(src, fmt, **kw).copy()
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/real_blocks.npz", **out)
print("saved", len(out))
