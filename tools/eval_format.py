"""Developer tool (runs on the GPU box): PSNR of our encoder vs the CPU oracle on generator-G
images, plus device-resident kernel throughput.  Usage:
    python tools/eval_format.py BC7 [--size 1024] [--big 4096] [--quality Normal] [--type UNorm]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import cuttlefish_b200 as cfx  # noqa: E402
import oracle  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("format")
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--big", type=int, default=4096)
    ap.add_argument("--quality", default="Normal")
    ap.add_argument("--type", default="UNorm")
    ap.add_argument("--kinds", default="noise+grad,gradient")
    ap.add_argument("--no-oracle", action="store_true")
    a = ap.parse_args()
    fmt = a.format
    cfx.init(0)
    kw = dict(quality=a.quality, type=a.type)
    for kind in a.kinds.split(","):
        n = a.size
        img = oracle.gen_image(kind, n, n)
        src = img.astype(np.float16) if kind == "hdr" else oracle.to_rgba8(img)
        peak = 64.0 if kind == "hdr" else 1.0
        got = cfx.encode(src, fmt, **kw)
        p_gpu = oracle.psnr_rgb(img, oracle.decode(got, fmt, n, n, type=a.type), peak)
        line = "%s %s %dx%d q=%s: gpu %.3f dB" % (fmt, kind, n, n, a.quality, p_gpu)
        if not a.no_oracle:
            t = time.time()
            ref = oracle.encode(img, fmt, **kw)
            dt = time.time() - t
            p_ref = oracle.psnr_rgb(img, oracle.decode(ref, fmt, n, n, type=a.type), peak)
            same = float(np.mean(np.all(got.reshape(-1, cfx.block_info(fmt)[2]) == ref.reshape(-1, cfx.block_info(fmt)[2]), axis=1)))
            line += " | ref %.3f dB (delta %+.3f) identical blocks %.1f%% | cpu %.2f Mtexel/s (%d thr)" % (
                p_ref, p_gpu - p_ref, 100 * same, n * n / dt / 1e6, oracle.hardware_threads())
        print(line, flush=True)
    # device-resident throughput
    n = a.big
    kind = a.kinds.split(",")[0]
    img = oracle.gen_image(kind, n, n)
    src = torch.from_numpy(img.astype(np.float16) if kind == "hdr" else oracle.to_rgba8(img)).cuda()
    out = torch.empty(cfx.encoded_size(fmt, n, n), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        cfx.encode_device(src, fmt, out=out, **kw)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    ev0.record()
    for _ in range(reps):
        cfx.encode_device(src, fmt, out=out, **kw)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    print("%s device-resident %dx%d: %.3f ms  -> %.1f Mtexel/s" % (fmt, n, n, ms, n * n / ms / 1e3), flush=True)


if __name__ == "__main__":
    main()
