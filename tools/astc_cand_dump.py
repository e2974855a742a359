"""Developer tool (GPU box, CFX_ASTC3_TUNE build): print (estimate terms, exact error) of every kept candidate of blocks.
    CFX_ASTC3_FLAGS=768 python tools/astc_cand_dump.py ASTC_6x6 40      (40 blocks drawn from every LDR real crop)"""
import glob, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cuttlefish_b200 as cfx
import torch
cfx.init(0)
fmt, n = sys.argv[1], int(sys.argv[2])
bw, bh = [int(x) for x in fmt.split("_")[1].split("x")]
rng = np.random.default_rng(1)
paths = sorted(glob.glob(os.path.join(HERE, "..", "tests", "golden", "real", "*.npz")))
for path in paths:
    name = os.path.basename(path)[:-4]
    src = np.load(path)["src"]
    if src.dtype != np.uint8 or (len(sys.argv) > 3 and name not in sys.argv[3:]):
        continue
    bx_n, by_n = 192 // bw, 192 // bh
    for b in rng.choice(bx_n * by_n, size=max(1, n // 9), replace=False):
        by, bx = divmod(int(b), bx_n)
        tile = np.ascontiguousarray(src[by * bh:(by + 1) * bh, bx * bw:(bx + 1) * bw])
        print("=== block %s %d" % (name, b), flush=True)
        cfx.encode(tile, fmt)
        torch.cuda.synchronize()
