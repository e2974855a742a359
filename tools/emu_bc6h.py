"""Developer tool: host emulation of the BC6H core vs the CPU oracle."""
import ctypes, os, subprocess, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402

so = os.path.join(HERE, "_build", "libemu_bc6h.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=fast", "-mfma", "-o", so, os.path.join(HERE, "emu_bc6h.cpp")])
lib = ctypes.CDLL(so)


def encode(img16, quality=2, signed=False):
    h, w, _ = img16.shape
    out = np.zeros(((w + 3) // 4) * ((h + 3) // 4) * 16, np.uint8)
    src = np.ascontiguousarray(img16.view(np.uint16))
    lib.emu_bc6h_encode(src.ctypes.data_as(ctypes.c_void_p), w, h, out.ctypes.data_as(ctypes.c_void_p), quality, 1 if signed else 0)
    return out


def hdr_noise(n, seed=7):
    rng = np.random.default_rng(seed)
    base = oracle.gen_image("hdr", n, n)
    img = base.copy()
    img[..., :3] *= (1.0 + 0.5 * rng.random((n, n, 3), dtype=np.float32))
    # some hard edges
    img[n // 3: n // 2, :, :3] *= 4.0
    return img


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    signed_img = hdr_noise(n)
    signed_img[..., :3] -= np.array([16.0, 2.0, 0.25], np.float32)
    for name, img, typ in [("hdr", oracle.gen_image("hdr", n, n), "UFloat"), ("hdr+noise", hdr_noise(n), "UFloat"),
                           ("signed hdr+noise", signed_img, "Float")]:
        img16 = img.astype(np.float16)
        imgf = img16.astype(np.float32)
        got = encode(img16, signed=typ == "Float")
        ref = oracle.encode(imgf, "BC6H", type=typ)
        dg = oracle.decode(got, "BC6H", n, n, type=typ)
        dr = oracle.decode(ref, "BC6H", n, n, type=typ)
        peak = 64.0
        pg, pr = oracle.psnr_rgb(imgf, dg, peak), oracle.psnr_rgb(imgf, dr, peak)
        lg = lambda d: float(np.sqrt(np.mean((np.log2(np.maximum(d[..., :3], 1e-4)) - np.log2(np.maximum(imgf[..., :3], 1e-4))) ** 2)))
        modes = np.bincount([(b & 0x1F) if (b & 2) else (b & 1) for b in got.reshape(-1, 16)[:, 0]], minlength=32)
        print("%s %d: emu %.3f dB ref %.3f dB delta %+.3f | log2 rmse emu %.5f ref %.5f | modes %s" % (
            name, n, pg, pr, pg - pr, lg(dg), lg(dr), {i: int(c) for i, c in enumerate(modes) if c}))
