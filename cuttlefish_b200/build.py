"""Build cuttlefish_b200/lib/libcfx.so (the C-ABI encoder library) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
"""
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, os.environ.get("CFX_BUILD_DIR", "lib"))      # (developer: variant builds side by side)
OBJ = os.path.join(OUT, "obj")
LIB = os.path.join(OUT, "libcfx.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--use_fast_math=false", "-Xptxas", "-v",
]

# (source, extra flags, macro that advertises it to cfx.cu)
UNITS = [
    ("cfx.cu", [], None),
    ("host_stage.cpp", [], None),
    ("containers.cpp", [], None),
    ("resize.cu", ["-fmad=false", "-Xcompiler", "-ffp-contract=off"], None),
    ("bc4_bc5.cu", [], None),
    ("bc7.cu", [], "CFX_HAVE_BC7"),
    ("bc1_bc3.cu", ["-fmad=false"], "CFX_HAVE_BC1"),
    ("etc.cu", ["-fmad=false"], "CFX_HAVE_ETC"),
    ("bc6h.cu", [], "CFX_HAVE_BC6H"),
    ("astc_host.cu", [], "CFX_HAVE_ASTC"),
    ("astc3.cu", (["-DCFX_ASTC3_TUNE=1"] if os.environ.get("CFX_ASTC3_TUNE") else []) + os.environ.get("CFX_ASTC3_DEFS", "").split(), None),
]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _digest(paths, flags):
    h = hashlib.sha256(" ".join(flags).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _compile(src, flags, defines, headers, verbose):
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    stamp = obj + ".sha"
    cmd = [nvcc()] + [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + flags + defines + \
        ["-c", os.path.join(CSRC, src), "-o", obj]
    digest = _digest([os.path.join(CSRC, src)] + headers, cmd)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
        return obj, ""
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(obj + ".log", "w") as f:
        f.write(r.stdout + r.stderr)
    with open(stamp, "w") as f:
        f.write(digest)
    return obj, (r.stdout + r.stderr) if verbose else ""


GENERATED = os.path.join(CSRC, "generated", "rgbcx_tables.inc")


def _generate_tables():
    """The byte-exact BC1 path needs rgbcx's data tables, dumped from the reference at build time
    (never committed). Without /root/reference and without a previously generated file, BC1 falls back
    to our own search (PSNR parity) and cfx_format_is_exact() says so."""
    if os.path.exists(GENERATED):
        return True
    ref = os.environ.get("CFX_REFERENCE", "/root/reference")
    tool = os.path.join(HERE, "..", "tools", "gen_rgbcx_tables.py")
    if os.path.exists(os.path.join(ref, "lib", "bc7enc_rdo", "rgbcx.cpp")) and os.path.exists(tool):
        r = subprocess.run([sys.executable, tool, ref], capture_output=True, text=True)
        if r.returncode != 0:
            print("warning: rgbcx table generation failed:\n" + r.stdout + r.stderr)
    return os.path.exists(GENERATED)


def build(verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    have_tables = _generate_tables()
    units = [(s, f, m) for (s, f, m) in UNITS if os.path.exists(os.path.join(CSRC, s))]
    defines = ["-D%s=1" % m for (_, _, m) in units if m]
    headers = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".cuh", ".h", ".hpp", ".inc"))]
    headers.append(os.path.join(HERE, "..", "include", "cfx.h"))
    if have_tables:
        headers.append(GENERATED)
    ref = os.environ.get("CFX_REFERENCE", "/root/reference")
    objs = []
    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        extra = ["-DCFX_HAVE_RGBCX_TABLES=1"] if have_tables else []
        futs = [ex.submit(_compile, s, f + extra + ["-DCFX_REFERENCE_DIR=\"%s\"" % ref], defines if s == "cfx.cu" else [],
                          headers, verbose) for (s, f, _) in units]
        for fu in futs:
            obj, log = fu.result()
            objs.append(obj)
            if log:
                print(log)
    newest = max(os.path.getmtime(o) for o in objs)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [nvcc(), "-shared", "-o", LIB, "-cudart", "static"] + objs + ["-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
