"""Developer tool: host emulation of the BC1 core vs the CPU oracle."""
import ctypes, os, subprocess, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402
so = os.path.join(HERE, "_build", "libemu_bc1.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "emu_bc1.cpp")])
lib = ctypes.CDLL(so)


def encode(src, flags=3, descent=2):
    h, w, _ = src.shape
    out = np.zeros(((w + 3) // 4) * ((h + 3) // 4) * 8, np.uint8)
    lib.emu_bc1_encode(src.ctypes.data_as(ctypes.c_void_p), w, h, out.ctypes.data_as(ctypes.c_void_p), flags, descent)
    return out


if __name__ == "__main__":
    from PIL import Image
    R = "/root/reference/lib/astc-encoder/Test/Images/Small/"
    descent = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    cases = [("noise+grad", 256), ("gradient", 256), ("gradient", 1024), (R + "LDR-RGB/ldr-rgb-00.png", 0), (R + "LDR-RGB/ldr-rgb-03.png", 0),
             ("/root/reference/lib/compressonator/runtime/images/ruby.png", 0)]
    for kind, n in cases:
        if os.path.exists(kind):
            src = np.ascontiguousarray(np.array(Image.open(kind).convert("RGBA")))
            img = src.astype(np.float32) / np.float32(255)
        else:
            img = oracle.gen_image(kind, n, n); src = oracle.to_rgba8(img)
        h, w = src.shape[:2]
        got = encode(src, 3, descent)
        ref = oracle.encode(img, "BC1_RGB")
        pg = oracle.psnr_rgb(img, oracle.decode(got, "BC1_RGB", w, h))
        pr = oracle.psnr_rgb(img, oracle.decode(ref, "BC1_RGB", w, h))
        same = np.mean(np.all(got.reshape(-1, 8) == ref.reshape(-1, 8), axis=1))
        print("%s %dx%d: emu %.3f dB ref %.3f dB delta %+.3f identical %.1f%%" % (os.path.basename(kind), w, h, pg, pr, pg - pr, 100 * same))
