"""CPU: the byte-exact restatements (csrc/bc1_exact.cuh, csrc/etc1_exact.cuh) are plain C++ through hostdev.h; compiled for
the host they must reproduce the committed reference outputs (tests/golden/real/*.npz: crops of the reference's own images
encoded by rgbcx / etc2comp at the five quality levels) byte for byte -- the same source the GPU kernels compile."""
import ctypes
import glob
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
REAL = os.path.join(ROOT, "tests", "golden", "real")
BUILD = os.path.join(ROOT, "tools", "_build")
TABLES = os.path.join(ROOT, "cuttlefish_b200", "csrc", "generated", "rgbcx_tables.inc")
EFFORT = {"Lowest": 0.0, "Low": 20.0, "Normal": 40.0, "High": 70.0, "Highest": 100.0}
QUALITY = {"Lowest": 0, "Low": 1, "Normal": 2, "High": 3, "Highest": 4}


def _compile(name, defines=()):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "libtest_%s.so" % name)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", *defines, "-x", "c++",
                           os.path.join(ROOT, "tools", "%s.cpp" % name), "-o", so])
    return ctypes.CDLL(so)


def _cases(prefix):
    out = []
    for f in sorted(glob.glob(os.path.join(REAL, "*.npz"))):
        z = np.load(f)
        if z["src"].dtype != np.uint8:
            continue
        for k in z.files:
            if k.startswith("blocks__%s__" % prefix):
                out.append((os.path.basename(f)[:-4], z["src"], k.split("__")[2], z[k]))
    return out


def test_bc1_restatement_matches_rgbcx_goldens_at_every_level():
    if not os.path.exists(TABLES):
        pytest.skip("csrc/generated/rgbcx_tables.inc not generated (needs /root/reference at build time)")
    lib = _compile("emu_bc1x", ["-DCFX_HAVE_RGBCX_TABLES=1"])
    cases = _cases("BC1_RGB")
    assert len(cases) >= 20 and {q for _, _, q, _ in cases} == set(QUALITY)
    for name, src, q, ref in cases:
        h, w, _ = src.shape
        s = np.ascontiguousarray(src).copy(); s[..., 3] = 255
        got = np.zeros(ref.size, np.uint8)
        lib.emu_bc1x_encode(s.ctypes.data_as(ctypes.c_void_p), w, h, got.ctypes.data_as(ctypes.c_void_p), 1, 1, QUALITY[q])
        bad = int(np.sum(np.any(got.reshape(-1, 8) != np.asarray(ref).reshape(-1, 8), axis=1)))
        assert bad == 0, "BC1_RGB %s %s: %d blocks differ" % (name, q, bad)


def test_etc1_restatement_matches_etc2comp_goldens_at_every_level():
    lib = _compile("emu_etc1x")
    lib.emu_etc1x_encode.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_float, ctypes.c_int]
    cases = _cases("ETC1")
    assert len(cases) >= 21 and {q for _, _, q, _ in cases} == set(EFFORT)
    for name, src, q, ref in cases:
        h, w, _ = src.shape
        img = np.ascontiguousarray(src.astype(np.float32)/np.float32(255))
        got = np.zeros(ref.size, np.uint8)
        lib.emu_etc1x_encode(img.ctypes.data, w, h, got.ctypes.data, EFFORT[q], 0)
        bad = int(np.sum(np.any(got.reshape(-1, 8) != np.asarray(ref).reshape(-1, 8), axis=1)))
        assert bad == 0, "ETC1 %s %s: %d blocks differ" % (name, q, bad)
