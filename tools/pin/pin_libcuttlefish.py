"""DEVELOPER TOOL (this container only): build the reference's full libcuttlefish.so out of
tree, run its real Texture::convert() on generator-G images and compare the bytes with oracle/cfref (the oracle every
golden and parity test uses).  Nothing here is imported by the product, the tests or bench.py.

    cmake -S /root/reference -B /tmp/cfbuild -G Ninja -DCMAKE_BUILD_TYPE=Release -DCUTTLEFISH_BUILD_TESTS=OFF \
          -DCUTTLEFISH_BUILD_DOCS=OFF -DCUTTLEFISH_BUILD_PVRTC=OFF -DCUTTLEFISH_SHARED=ON && cmake --build /tmp/cfbuild -j16
    python tools/pin/pin_libcuttlefish.py /tmp/cfbuild > profiles/r02_libcuttlefish_pin.txt
"""
import hashlib, os, subprocess, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle
from cuttlefish_b200.api import FORMATS, QUALITY, TYPES

build = sys.argv[1] if len(sys.argv) > 1 else "/tmp/cfbuild"
libdir = os.path.join(build, "output")
exe = os.path.join(tempfile.gettempdir(), "pin_libcuttlefish")
subprocess.check_call(["g++", "-O2", "-std=c++14", os.path.join(HERE, "pin_libcuttlefish.cpp"), "-I/root/reference/lib/include",
                       "-I" + os.path.join(build, "lib", "include"), "-L" + libdir, "-lcuttlefish", "-Wl,-rpath," + libdir, "-o", exe])


def fnv1a64(b):
    h = 0xcbf29ce484222325
    for x in np.frombuffer(b, np.uint8).tolist():
        h = ((h ^ x) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


CASES = [("BC1_RGB", "UNorm", "Normal", "gradient", 256, 256, 0), ("BC1_RGB", "UNorm", "Highest", "noise+grad", 128, 128, 0),
         ("BC3", "UNorm", "Normal", "noise+grad", 128, 128, 0), ("BC4", "UNorm", "Normal", "noise+grad", 128, 128, 0),
         ("BC5", "UNorm", "High", "noise+grad", 64, 64, 0), ("BC7", "UNorm", "Normal", "noise+grad", 128, 128, 0),
         ("BC7", "UNorm", "Normal", "noise+grad", 97, 61, 1), ("BC6H", "UFloat", "Normal", "hdr", 64, 64, 0),
         ("ETC1", "UNorm", "Normal", "gradient", 128, 128, 0), ("ETC2_R8G8B8A8", "UNorm", "Normal", "noise+grad", 64, 64, 0),
         ("ETC2_R8G8B8", "UNorm", "Low", "noise+grad", 50, 30, 1), ("ASTC_6x6", "UNorm", "Normal", "noise+grad", 96, 96, 0),
         ("ASTC_4x4", "UNorm", "Low", "gradient", 64, 64, 0), ("ASTC_8x8", "UNorm", "Normal", "noise+grad", 100, 52, 1)]
ok = 0
for fmt, typ, q, kind, w, h, srgb in CASES:
    img = oracle.gen_image(kind, w, h, seed=5)
    with tempfile.TemporaryDirectory() as d:
        src, dst = os.path.join(d, "in.f32"), os.path.join(d, "out.bin")
        np.ascontiguousarray(img, np.float32).tofile(src)
        rc = subprocess.call([exe, src, str(w), str(h), str(FORMATS[fmt]), str(TYPES[typ]), str(QUALITY[q]), str(srgb), dst])
        real = np.fromfile(dst, np.uint8) if rc == 0 else None
    ours = oracle.encode(img, fmt, type=typ, quality=q, srgb=bool(srgb))
    same = real is not None and real.size == ours.size and np.array_equal(real, ours)
    ok += same
    print("%-14s %-6s %-8s %-10s %4dx%-4d srgb=%d  libcuttlefish %s  cfref %s  %s" % (
        fmt, typ, q, kind, w, h, srgb, fnv1a64(real.tobytes()) if real is not None else "rc=%d" % rc, fnv1a64(ours.tobytes()),
        "IDENTICAL" if same else "DIFFERENT"))
print("%d of %d cases byte-identical between the real Texture::convert() and oracle/cfref" % (ok, len(CASES)))
