"""Developer tool (GPU box): ASTC HDR (Type::UFloat) PSNR against the CPU oracle.
    python tools/astc_hdr_eval.py ASTC_6x6 ..."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuttlefish_b200 as cfx
import oracle
cfx.init(0)
n = 192
rng = np.random.default_rng(7)
ramp = oracle.gen_image("hdr", n, n)
noisy = ramp.copy(); noisy[..., :3] *= (1.0 + 0.5*rng.random((n, n, 3), dtype=np.float32)); noisy[n//3:n//2, :, :3] *= 4.0
for fmt in sys.argv[1:]:
    for name, img in (("ramp", ramp), ("ramp+noise", noisy)):
        img16 = img.astype(np.float16); imgf = img16.astype(np.float32)
        ref = oracle.encode(imgf, fmt, type="UFloat")
        dr = oracle.decode(ref, fmt, n, n, type="UFloat")
        line = "%s %s: ref %.3f dB" % (fmt, name, oracle.psnr_rgb(imgf, dr, 64.0))
        for src in (img16, imgf):
            got = cfx.encode(src, fmt, type="UFloat")
            dg = oracle.decode(got, fmt, n, n, type="UFloat")
            lg = lambda d: float(np.sqrt(np.mean((np.log2(np.maximum(d[..., :3], 1e-4)) - np.log2(np.maximum(imgf[..., :3], 1e-4)))**2)))
            line += " | gpu(%s) %.3f dB (delta %+.3f) log2rmse %.4f vs %.4f alpha ok %s" % (src.dtype, oracle.psnr_rgb(imgf, dg, 64.0),
                oracle.psnr_rgb(imgf, dg, 64.0) - oracle.psnr_rgb(imgf, dr, 64.0), lg(dg), lg(dr), bool(np.allclose(dg[..., 3], 1.0)))
        print(line, flush=True)
