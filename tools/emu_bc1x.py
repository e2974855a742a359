"""Developer tool (CPU): the byte-exact BC1 restatement (csrc/bc1_exact.cuh) compiled for the host, against the
reference encoder (oracle) on the synthetic generators and the vendored real crops, at every quality level.
    python tools/emu_bc1x.py"""
import ctypes, glob, os, subprocess, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import oracle
out = os.path.join(ROOT, "tools", "_build")
os.makedirs(out, exist_ok=True)
so = os.path.join(out, "libemu_bc1x.so")
subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-DCFX_HAVE_RGBCX_TABLES=1", "-x", "c++",
                       os.path.join(ROOT, "tools", "emu_bc1x.cpp"), "-o", so])
lib = ctypes.CDLL(so)
inputs = [(k + " %d" % n, oracle.to_rgba8(oracle.gen_image(k, n, n))) for k, n in (("noise+grad", 256), ("gradient", 256), ("ui", 288))]
inputs.append(("ragged", oracle.to_rgba8(oracle.gen_image("noise+grad", 97, 61))))
for f in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "real", "*.npz"))):
    src = np.load(f)["src"]
    if src.dtype == np.uint8:
        inputs.append((os.path.basename(f), src))
bad = 0
for name, src in inputs:
    h, w, _ = src.shape
    for q, quality in ((0, "Lowest"), (1, "Low"), (2, "Normal"), (3, "High"), (4, "Highest")):
        for fmt, a3, ab in (("BC1_RGB", 1, 1), ("BC3", 0, 0)):
            got = np.zeros(((h + 3)//4)*((w + 3)//4)*8, np.uint8)
            s = np.ascontiguousarray(src).copy()
            if fmt == "BC1_RGB":
                s[..., 3] = 255
            lib.emu_bc1x_encode(s.ctypes.data_as(ctypes.c_void_p), w, h, got.ctypes.data_as(ctypes.c_void_p), a3, ab, q)
            ref = oracle.encode(s.astype(np.float32)/np.float32(255), fmt, quality=quality)
            ref = ref.reshape(-1, 16)[:, 8:].reshape(-1) if fmt == "BC3" else ref
            n = int(np.sum(np.any(got.reshape(-1, 8) != ref.reshape(-1, 8), axis=1)))
            bad += n
            print("%-18s %-7s %-7s mismatching blocks %d of %d" % (name, fmt, quality, n, got.size//8), flush=True)
print("TOTAL mismatches", bad)
