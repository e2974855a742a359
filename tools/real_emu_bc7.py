"""Developer tool: host emulation of the BC7 search over the committed real-image goldens (PSNR delta vs reference)."""
import glob, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import emu, oracle
lib = emu.build("bc7")
specs = sys.argv[1:] or ["6,1,3,1r1/6,7,7r1,7r2"]
for spec in specs:
    so, sa = spec.split("/") if "/" in spec else (spec, "")
    line = "%-40s" % spec[:39]
    for path in sorted(glob.glob(os.path.join(HERE, "..", "tests", "golden", "real", "*.npz"))):
        z = np.load(path)
        if "blocks__BC7__Normal" not in z.files:
            continue
        src = z["src"]; img = src.astype(np.float32) / np.float32(255)
        ref = z["blocks__BC7__Normal"]
        got, _ = emu.bc7_encode(lib, src, emu.parse_cands(so), emu.parse_cands(sa) if sa else None)
        dg, dr = oracle.decode(got, "BC7", 192, 192), oracle.decode(ref, "BC7", 192, 192)
        p4 = lambda d: 10*np.log10(1/np.mean((d.astype(np.float64)-img)**2))
        line += " %s %+.2f/%+.2f" % (os.path.basename(path)[:-4], oracle.psnr_rgb(img, dg) - oracle.psnr_rgb(img, dr), p4(dg)-p4(dr))
    print(line)
