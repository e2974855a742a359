"""Developer probe: end-to-end time of cfx_encode for the four kinds of host buffers (run on the GPU box)."""
import ctypes
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import cuttlefish_b200 as cfx
from cuttlefish_b200 import _lib, synth


def pinned_like(arr):
    lib = _lib.load()
    ptr = lib.cfx_host_alloc(arr.nbytes)
    buf = (ctypes.c_uint8 * arr.nbytes).from_address(ptr)
    out = np.frombuffer(buf, dtype=arr.dtype).reshape(arr.shape)
    out[...] = arr
    return out


def main():
    fmt = sys.argv[1] if len(sys.argv) > 1 else "BC7"
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    pools = [[0]]
    import torch
    n = torch.cuda.device_count()
    if n > 1:
        pools.append(list(range(n)))
    img8 = synth.to_rgba8(synth.gen_image("noise+grad", size, size))
    imgf = img8.astype(np.float32) / np.float32(255.0)
    out_n = cfx.encoded_size(fmt, size, size)
    out_pageable = np.empty(out_n, np.uint8)
    out_pinned = pinned_like(out_pageable)
    srcs = {"rgba8 pinned": pinned_like(img8), "rgba8 pageable": img8, "rgba32f pinned": pinned_like(imgf), "rgba32f pageable": imgf}
    for pool in pools:
        cfx.set_devices(pool)
        ref = None
        for name, src in srcs.items():
            for oname, out in (("pinned out", out_pinned), ("pageable out", out_pageable)):
                for _ in range(2):
                    cfx.encode(src, fmt, out=out)
                t = []
                for _ in range(5):
                    t0 = time.perf_counter()
                    cfx.encode(src, fmt, out=out)
                    t.append(time.perf_counter() - t0)
                if ref is None:
                    ref = out.copy()
                same = np.array_equal(out, ref)
                print("%s pool=%d %-17s %-12s best %.2f ms  median %.2f ms  %.0f Mtexels/s  same=%s" %
                      (fmt, len(pool), name, oname, min(t) * 1e3, sorted(t)[2] * 1e3, size * size / sorted(t)[2] / 1e6, same), flush=True)


if __name__ == "__main__":
    main()
