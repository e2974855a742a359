"""Developer tool (CPU): the byte-exact ETC1 restatement (csrc/etc1_exact.cuh) compiled for the host, against the
reference encoder (oracle) on the synthetic generators and the vendored real crops, at every quality level.
    python tools/emu_etc1x.py [High Highest]"""
import ctypes, glob, os, subprocess, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import oracle
out = os.path.join(ROOT, "tools", "_build")
os.makedirs(out, exist_ok=True)
so = os.path.join(out, "libemu_etc1x.so")
subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++",
                       os.path.join(ROOT, "tools", "emu_etc1x.cpp"), "-o", so])
lib = ctypes.CDLL(so)
lib.emu_etc1x_encode.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_float, ctypes.c_int]
EFFORT = {"Lowest": 0.0, "Low": 20.0, "Normal": 40.0, "High": 70.0, "Highest": 100.0}
levels = sys.argv[1:] or list(EFFORT)
inputs = [(k + " %d" % n, oracle.gen_image(k, n, n)) for k, n in (("noise+grad", 128), ("gradient", 128), ("ui", 96))]
inputs.append(("ragged", oracle.gen_image("noise+grad", 97, 61)))
for f in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "real", "*.npz"))):
    src = np.load(f)["src"]
    if src.dtype == np.uint8:
        inputs.append((os.path.basename(f), src[:96, :96].astype(np.float32)/np.float32(255)))
bad = 0
for name, img in inputs:
    img = np.ascontiguousarray(img, np.float32)
    h, w, _ = img.shape
    for quality in levels:
      for srgb in (False, True):
        got = np.zeros(((h + 3)//4)*((w + 3)//4)*8, np.uint8)
        lib.emu_etc1x_encode(img.ctypes.data, w, h, got.ctypes.data, EFFORT[quality], 1 if srgb else 0)
        ref = oracle.encode(img, "ETC1", quality=quality, srgb=srgb)
        n = int(np.sum(np.any(got.reshape(-1, 8) != ref.reshape(-1, 8), axis=1)))
        bad += n
        print("%-18s ETC1 %-7s %-6s mismatching blocks %d of %d" % (name, quality, "sRGB" if srgb else "linear", n, got.size//8), flush=True)
print("TOTAL mismatches", bad)
