"""Developer tool: host emulation of the ETC core vs the CPU oracle."""
import ctypes, os, subprocess, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402
so = os.path.join(HERE, "_build", "libemu_etc.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "emu_etc.cpp")])
lib = ctypes.CDLL(so)
FMT = {"ETC1": 37, "ETC2_R8G8B8": 38, "ETC2_R8G8B8A1": 39, "ETC2_R8G8B8A8": 40}


def encode(img, fmt, rounds=2):
    h, w, _ = img.shape
    out = np.zeros(((w + 3) // 4) * ((h + 3) // 4) * (16 if fmt == "ETC2_R8G8B8A8" else 8), np.uint8)
    img = np.ascontiguousarray(img, np.float32)
    lib.emu_etc_encode(img.ctypes.data_as(ctypes.c_void_p), w, h, out.ctypes.data_as(ctypes.c_void_p), FMT[fmt], rounds)
    return out


if __name__ == "__main__":
    from PIL import Image
    R = "/root/reference/lib/astc-encoder/Test/Images/Small/"
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    cases = [("noise+grad", 128), ("gradient", 128), ("gradient", 512), (R + "LDR-RGB/ldr-rgb-00.png", 0), (R + "LDR-RGBA/ldr-rgba-00.png", 0)]
    for fmt in ["ETC1", "ETC2_R8G8B8", "ETC2_R8G8B8A8"]:
        for kind, n in cases:
            if os.path.exists(kind):
                src = np.ascontiguousarray(np.array(Image.open(kind).convert("RGBA")))
                img = src.astype(np.float32) / np.float32(255)
            else:
                img = oracle.gen_image(kind, n, n)
            h, w = img.shape[:2]
            got = encode(img, fmt, rounds)
            ref = oracle.encode(img, fmt)
            dg, dr = oracle.decode(got, fmt, w, h), oracle.decode(ref, fmt, w, h)
            pg, pr = oracle.psnr_rgb(img, dg), oracle.psnr_rgb(img, dr)
            ag = 10 * np.log10(1 / max(np.mean((dg[..., 3] - img[..., 3]) ** 2), 1e-12))
            ar = 10 * np.log10(1 / max(np.mean((dr[..., 3] - img[..., 3]) ** 2), 1e-12))
            print("%s %s %dx%d: emu %.3f dB ref %.3f dB delta %+.3f | alpha emu %.2f ref %.2f" % (fmt, os.path.basename(kind), w, h, pg, pr, pg - pr, ag, ar))
