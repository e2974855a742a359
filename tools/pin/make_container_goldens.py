"""DEVELOPER TOOL (this container only): golden DDS / KTX headers of every block (format, type, colour space, alpha type)
pair and two whole files, written by the reference's real Texture::save() (full libcuttlefish.so built out of tree, see
pin_libcuttlefish.py).  Output: tests/golden/containers/headers.npz (+ bc4_mips.dds, bc1_mips.ktx and their source)."""
import os, subprocess, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
import oracle
from cuttlefish_b200.api import ALPHA, FORMATS, QUALITY, TYPES

build = sys.argv[1] if len(sys.argv) > 1 else "/tmp/cfbuild"
libdir = os.path.join(build, "output")
exe = os.path.join(tempfile.gettempdir(), "make_container_goldens")
subprocess.check_call(["g++", "-O2", "-std=c++14", os.path.join(HERE, "make_container_goldens.cpp"), "-I/root/reference/lib/include",
                       "-I" + os.path.join(build, "lib", "include"), "-L" + libdir, "-lcuttlefish", "-Wl,-rpath," + libdir, "-o", exe])
OUT = os.path.join(ROOT, "tests", "golden", "containers")
os.makedirs(OUT, exist_ok=True)
TYPES_OF = {"BC4": ["UNorm", "SNorm"], "BC5": ["UNorm", "SNorm"], "BC6H": ["UFloat", "Float"], "EAC_R11": ["UNorm", "SNorm"],
            "EAC_R11G11": ["UNorm", "SNorm"]}


def run(img, fmt, typ, srgb, alpha, mips, ftype, path, quality="Lowest"):
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "in.f32")
        np.ascontiguousarray(img, np.float32).tofile(src)
        h, w, _ = img.shape
        return subprocess.call([exe, src, str(w), str(h), str(FORMATS[fmt]), str(TYPES[typ]), str(QUALITY[quality]), str(int(srgb)),
                                str(ALPHA[alpha]), str(int(mips)), str(ftype), path])


heads = {}
w, h = 40, 24
for fmt in FORMATS:
    types = TYPES_OF.get(fmt, ["UNorm"]) + (["UFloat"] if fmt.startswith("ASTC") else [])
    for typ in types:
        kind = "hdr" if typ in ("UFloat", "Float") else "noise+grad"
        img = oracle.gen_image(kind, w, h, seed=3)
        for srgb in (0, 1):
            for alpha in (["Standard"] if srgb else ["None", "Standard", "PreMultiplied", "Encoded"]):
                for mips in (0, 1):
                    for ftype, ext, hdr in ((1, "dds", 148), (2, "ktx", 64)):
                        with tempfile.TemporaryDirectory() as d:
                            p = os.path.join(d, "t." + ext)
                            rc = run(img, fmt, typ, srgb, alpha, mips, ftype, p)
                            key = "%s__%s__%d__%s__%d__%s" % (fmt, typ, srgb, alpha, mips, ext)
                            if rc in (6, 10):      # convert() refuses the pair (e.g. sRGB BC4) / the container has no such format
                                heads[key] = np.zeros(0, np.uint8)
                                continue
                            assert rc == 0, (key, rc)
                            data = np.fromfile(p, np.uint8)
                            heads[key] = data[:hdr].copy()
                            heads[key + "__size"] = np.array([data.size], np.int64)
np.savez_compressed(os.path.join(OUT, "headers.npz"), **heads)
print("headers:", len(heads))
# two whole files from byte-exact formats at every quality (BC4) / at Normal (BC1), with the full generated mip chain
img = oracle.gen_image("noise+grad", 52, 36, seed=9)
np.save(os.path.join(OUT, "source_52x36.npy"), img.astype(np.float32))
assert run(img, "BC4", "UNorm", 0, "Standard", 1, 1, os.path.join(OUT, "bc4_mips.dds"), "Normal") == 0
assert run(img, "BC1_RGB", "UNorm", 0, "Standard", 1, 2, os.path.join(OUT, "bc1_mips.ktx"), "Normal") == 0
print("files written")
