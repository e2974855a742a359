import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases(prefixes=None):
    out = []
    for path in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        name = os.path.basename(path)[:-4]
        if prefixes is None or any(name.startswith(p + "_") for p in prefixes):
            out.append(name)
    return out


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    kw = eval(str(z["kw"]))  # written by make_goldens.py: a dict literal
    src = z["src"]
    if src.dtype == np.uint16:
        src = src.view(np.float16)
    return src, z["blocks"], str(z["format"]), kw


def src_as_float(src):
    if src.dtype == np.uint8:
        return src.astype(np.float32) / np.float32(255.0)
    return src.astype(np.float32)


def block_mismatches(a, b, block_bytes):
    a = np.asarray(a).reshape(-1, block_bytes)
    b = np.asarray(b).reshape(-1, block_bytes)
    return np.nonzero((a != b).any(axis=1))[0]
