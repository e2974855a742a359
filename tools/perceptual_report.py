"""Developer tool (GPU box): sRGB textures -- error of OUR blocks against the reference's blocks IN THE REFERENCE'S OWN
perceptual metric (the metric its encoder minimised for an sRGB descriptor): bc7enc's weighted YCbCr distance
(lib/bc7enc_rdo/bc7enc.cpp:505-529 with the perceptual weights 128 / 64 / 16), etc2comp's REC709 error
(EtcBlock4x4Encoding.cpp:157-180), astcenc's channel weights 0.30 / 0.59 / 0.11 (astcenc_entry.cpp:644-649).
Prints 10 log10(reference error / our error): >= -0.1 dB holds the north_star bar in that metric."""
import glob, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuttlefish_b200 as cfx
import oracle
cfx.init(0)


def err_bc7(d, x):
    a, b = np.rint(d[..., :3].astype(np.float64)*255).astype(np.int64), np.rint(x[..., :3].astype(np.float64)*255).astype(np.int64)
    def ycc(p):
        l = p[..., 0]*109 + p[..., 1]*366 + p[..., 2]*37
        return l, (p[..., 0] << 9) - l, (p[..., 2] << 9) - l
    l1, cr1, cb1 = ycc(a); l2, cr2, cb2 = ycc(b)
    dl, dcr, dcb = (l1 - l2) >> 8, (cr1 - cr2) >> 8, (cb1 - cb2) >> 8
    return float(np.mean(128*dl*dl + 64*dcr*dcr + 16*dcb*dcb))


def err_rec709(d, x):
    def lcc(p):
        p = p.astype(np.float64)
        l = p[..., 0]*0.2126 + p[..., 1]*0.7152 + p[..., 2]*0.0722
        return l, 0.5*(p[..., 0] - l)/(1 - 0.2126), 0.5*(p[..., 2] - l)/(1 - 0.0722)
    l1, r1, b1 = lcc(x); l2, r2, b2 = lcc(d)
    return float(np.mean(3*(l1 - l2)**2 + (r1 - r2)**2 + 0.5*(b1 - b2)**2))


def err_astc(d, x):
    e = (d[..., :3].astype(np.float64) - x[..., :3])**2
    return float(np.mean(e[..., 0]*0.30 + e[..., 1]*0.59 + e[..., 2]*0.11))


METRIC = {"BC7": err_bc7, "ETC2_R8G8B8": err_rec709, "ETC2_R8G8B8A8": err_rec709, "ASTC_4x4": err_astc, "ASTC_6x6": err_astc, "ASTC_8x8": err_astc}
inputs = [("noise+grad", oracle.gen_image("noise+grad", 256, 256)), ("ui", oracle.gen_image("ui", 288, 288))]
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "real")
for f in sorted(glob.glob(os.path.join(root, "*.npz"))):
    src = np.load(f)["src"]
    if src.dtype == np.uint8:
        inputs.append((os.path.basename(f)[:-4], src.astype(np.float32)/np.float32(255)))
for fmt, metric in METRIC.items():
    line = "%-14s" % fmt
    for name, img in inputs:
        h, w, _ = img.shape
        x = oracle.to_rgba8(img).astype(np.float32)/np.float32(255)
        got = cfx.encode(oracle.to_rgba8(img), fmt, srgb=True)
        ref = oracle.encode(x, fmt, srgb=True)
        eg, er = metric(oracle.decode(got, fmt, w, h), x), metric(oracle.decode(ref, fmt, w, h), x)
        line += "  %s %+.2f" % (name, 10*np.log10(max(er, 1e-30)/max(eg, 1e-30)))
    print(line, flush=True)
