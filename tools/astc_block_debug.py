"""Developer tool (GPU box, library built with CFX_ASTC3_TUNE=1): encode single blocks of a real crop under restrictions
(CFX_ASTC3_SLOT / CFX_ASTC3_LEVEL / CFX_ASTC3_NW are read once per process, so every variant is its own process).
    python tools/astc_block_debug.py rgb09 ASTC_6x6 482 29 852"""
import os, subprocess, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)

if os.environ.get("CFX_DEBUG_CHILD"):
    import cuttlefish_b200 as cfx
    import oracle
    from real_blockdiff import info
    cfx.init(0)
    name, fmt = sys.argv[1], sys.argv[2]
    bw, bh = [int(x) for x in fmt.split("_")[1].split("x")]
    z = np.load(os.path.join(HERE, "..", "tests", "golden", "real", name + ".npz"))
    src = z["src"]; ref = z["blocks__%s__Normal" % fmt].reshape(-1, 16)
    bx_n = (192 + bw - 1) // bw
    out = []
    for b in [int(x) for x in sys.argv[3:]]:
        by, bx = divmod(b, bx_n)
        tile = np.ascontiguousarray(src[by * bh:(by + 1) * bh, bx * bw:(bx + 1) * bw])
        img = tile.astype(np.float32) / np.float32(255)
        got = cfx.encode(tile, fmt, quality=os.environ.get("CFX_DEBUG_QUALITY", "Normal"))
        e = lambda blk: float((((oracle.decode(blk, fmt, bw, bh)[..., :3].astype(np.float64) - img[..., :3]) ** 2).sum()) * 65025)
        out.append("blk %d: ours %.0f %s | ref %.0f %s" % (b, e(got), info(got.reshape(-1, 16)[0]), e(ref[b]), info(ref[b])))
    print("\n".join(out))
    sys.exit(0)

variants = [("free", {}), ("slot13", {"CFX_ASTC3_SLOT": "13"}), ("slot13 L6 nw36", {"CFX_ASTC3_SLOT": "13", "CFX_ASTC3_LEVEL": "6", "CFX_ASTC3_NW": "36"}),
            ("slot0", {"CFX_ASTC3_SLOT": "0"}), ("slot0 nw30", {"CFX_ASTC3_SLOT": "0", "CFX_ASTC3_NW": "30"}),
            ("slot14", {"CFX_ASTC3_SLOT": "14"}), ("slot1", {"CFX_ASTC3_SLOT": "1"}), ("free Highest", {"CFX_DEBUG_QUALITY": "Highest"})]
if os.environ.get("CFX_DEBUG_VARIANTS"):      # "label:K=V,K=V;label2:..."
    variants = []
    for item in os.environ["CFX_DEBUG_VARIANTS"].split(";"):
        label, _, kv = item.partition(":")
        variants.append((label, dict(x.split("=") for x in kv.split(",") if x)))
for label, env in variants:
    e = dict(os.environ, CFX_DEBUG_CHILD="1", **env)
    r = subprocess.run([sys.executable, __file__] + sys.argv[1:], env=e, capture_output=True, text=True)
    print("== " + label)
    print(r.stdout.strip() or r.stderr[-400:])
