#!/usr/bin/env python
"""bench.py -- Mtexels/s of the block-encode hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--format BC7] [--impl reference]

A step = one pass of the encoder over one batch of synthetic input (generator G).
Default workload = BASELINE.json configs[1]: BC7 UNorm, 8192x8192 RGBA8 noise+grad, quality
Normal.  At N GPUs the batch is an N-layer array texture of that size (Converter::convert's
depth/face loop, lib/src/Converter.cpp:521-527); every layer is sharded by block row across the
ranks, so per-GPU work is fixed ("weak"), and the packed blocks are gathered on
rank 0 with one NCCL gather per step.

  value   device-timed (CUDA events, max over ranks): inputs resident in HBM, kernels + gather.
  e2e     the same work through cfx_encode() with HOST buffers: pinned host -> device copy,
          kernels, device -> pinned host copy all inside the timed region.
  roofline  algorithmic bytes (source read once + blocks written once) / kernel duration, vs the
          measured HBM copy peak in MEASURED_PEAKS.json.
  cpu_baseline  the reference's own CPU encoder (oracle/_ref, built from /root/reference) on a
          bounded crop of the same image, all host threads.  The ONLY use of oracle/ here.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# bytes per texel read / written
WORKLOADS = {
    "BC7": dict(size=8192, kind="noise+grad", src="RGBA8", type="UNorm", read=4.0, write=1.0),
    "BC1_RGB": dict(size=8192, kind="noise+grad", src="RGBA8", type="UNorm", read=4.0, write=0.5),
    "BC3": dict(size=8192, kind="noise+grad", src="RGBA8", type="UNorm", read=4.0, write=1.0),
    "BC4": dict(size=8192, kind="noise+grad", src="RGBA8", type="UNorm", read=4.0, write=0.5),
    "BC5": dict(size=8192, kind="noise+grad", src="RGBA8", type="UNorm", read=4.0, write=1.0),
    "ETC1": dict(size=4096, kind="noise+grad", src="RGBA8", type="UNorm", read=4.0, write=0.5),
    "ETC2_R8G8B8": dict(size=4096, kind="noise+grad", src="RGBA8", type="UNorm", read=4.0, write=0.5),
    "ETC2_R8G8B8A8": dict(size=4096, kind="noise+grad", src="RGBA8", type="UNorm", read=4.0, write=1.0),
    "BC6H": dict(size=4096, kind="hdr", src="RGBA16F", type="UFloat", read=8.0, write=1.0),
    "ASTC_6x6": dict(size=8192, kind="noise+grad", src="RGBA8", type="UNorm", read=4.0, write=16.0 / 36.0),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--format", default="BC7")
    ap.add_argument("--quality", default="Normal")
    ap.add_argument("--size", type=int, default=0, help="override the square image size")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=4096, help="CPU baseline crop is SxS texels (4096: 10-30 s of CPU work)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--mips", action="store_true",
                    help="encode the full mip chain of the image (BASELINE config 5 shape): every rank encodes one "
                         "texture, level by level, through cfx_encode_device / cfx_encode_batch")
    ap.add_argument("--mipgen", action="store_true",
                    help="with --mips: build the chain on the GPU too (Texture::generateMipmaps, Catmull-Rom) from a float32 "
                         "level 0 through cfx_encode_mip_chain(_device); the reference arm then times FreeImage's chain + "
                         "Converter::convert")
    return ap.parse_args()


def measured_traffic(fmt, texels):
    """DRAM bytes of one launch from the committed ncu capture (profiles/r01_dram_traffic.json), scaled to the
    texels of the launch timed here; None when no capture exists for the format."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_dram_traffic.json")) as f:
            t = json.load(f)[fmt]
        return int((t["dram_read_bytes"] + t["dram_write_bytes"]) * (texels / t["texels"])), t["csv"]
    except Exception:
        return None, None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = the upper half of the samples (idle samples before/after the loop drop out)
        load = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference(a, wl, size, steps, warmup):
    """Times the reference's CPU encoders (oracle/_ref/libcfref.so) on a bounded crop."""
    import oracle  # the checker / CPU baseline: never on the product path
    from cuttlefish_b200 import synth
    s = min(a.cpu_sample, size)
    img = synth.gen_image(wl["kind"], size, size, rows=(0, s))[:, :s].copy()
    threads = oracle.hardware_threads()
    kw = dict(type=wl["type"], quality=a.quality)
    # the reference's own Converter::convert (std::thread pool, real converter glue, real encoders) when
    # oracle/_ref/libcfglue.so was built; else the same encoders under our byte-identical glue restatement
    enc = oracle.encode_glue if oracle.glue_available() else oracle.encode
    for _ in range(warmup):
        enc(img[: max(64, s // 8)], a.format, threads=0, **kw)
    times = []
    for _ in range(steps):
        t = time.perf_counter()
        enc(img, a.format, threads=0, **kw)
        times.append(time.perf_counter() - t)
    dt = float(np.mean(times))
    return {"value": s * s / dt / 1e6, "unit": "Mtexels/s", "cores": threads, "kind": "reference",
            "sample": "%dx%d top-left crop of the %dx%d %s image, %s quality=%s, oracle/_ref "
                      "(reference Converter::convert + encoders compiled from source), %d threads, %.2f s/step" %
                      (s, s, size, size, wl["kind"], a.format, a.quality, threads, dt)}, dt


def cpu_reference_mips(a, wl, size, steps, warmup):
    """Times the reference's Texture::generateMipmaps + convert on a bounded level 0: the real FreeImage_Rescale chain
    (oracle/_ref/libfiresize.so; single-threaded, as in the reference) and the real Converter::convert per level."""
    import oracle
    from oracle import resize as oresize
    from cuttlefish_b200 import synth
    s = min(a.cpu_sample // 2, size)
    img = synth.gen_image(wl["kind"], size, size, rows=(0, s))[:, :s].copy().astype(np.float32)
    threads = oracle.hardware_threads()
    kw = dict(type=wl["type"], quality=a.quality)
    enc = oracle.encode_glue if oracle.glue_available() else oracle.encode

    def chain():
        levels = oresize.mip_chain(img, "CatmullRom", fn=oresize.resize_ref)
        t_mid = time.perf_counter()
        for l in levels:
            enc(l, a.format, threads=0, **kw)
        return levels, t_mid

    for _ in range(warmup):
        enc(img[:64], a.format, threads=0, **kw)
    times, resize_times = [], []
    for _ in range(steps):
        t = time.perf_counter()
        levels, t_mid = chain()
        times.append(time.perf_counter() - t)
        resize_times.append(t_mid - t)
    dt = float(np.mean(times))
    texels = sum(l.shape[0] * l.shape[1] for l in levels)
    return {"value": texels / dt / 1e6, "unit": "Mtexels/s", "cores": threads, "kind": "reference",
            "sample": "%dx%d top-left crop of the %dx%d %s image as level 0, %d-level Catmull-Rom chain by the reference's "
                      "FreeImage_Rescale (1 thread, %.2f s) + %s quality=%s Converter::convert per level (%d threads), "
                      "%.2f s/step" % (s, s, size, size, wl["kind"], len(levels), float(np.mean(resize_times)), a.format,
                                       a.quality, threads, dt)}, dt


def main():
    a = parse()
    wl = dict(WORKLOADS[a.format])
    size = a.size or wl["size"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "%s %s encode, %dx%d %s synthetic %s (generator G), quality=%s" %
              (a.format, wl["type"], size, size, wl["src"], wl["kind"], a.quality),
              "format": a.format, "quality": a.quality, "width": size, "height": size,
              "layers": max(world, 1), "sharding": "block-row slabs of every layer across %d rank(s)" % world,
              "l2": "inputs (%d MiB per rank) larger than L2; no flush needed" %
                    (size * size * (4 if wl["src"] == "RGBA8" else 8) >> 20)}

    if a.impl == "reference":
        if rank != 0:
            return 0
        if a.mips and a.mipgen:
            base, dt = cpu_reference_mips(a, wl, size, a.steps, min(a.warmup, 1))
            config = dict(config, workload=config["workload"] + " + generateMipmaps(CatmullRom) + full mip chain")
        else:
            base, dt = cpu_reference(a, wl, size, a.steps, min(a.warmup, 1))
        print(json.dumps({"impl": "reference", "metric": "Mtexels/s encode", "value": base["value"],
                          "unit": "Mtexels/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": base,
                          "e2e": {"value": base["value"], "unit": "Mtexels/s", "h2d_bytes_per_step": 0,
                                  "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    import torch.distributed as dist
    import cuttlefish_b200 as cfx
    from cuttlefish_b200 import synth

    if a.mips:
        return bench_mips(a, wl, size, rank, world, local, config)

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfx.init(local)
    kw = dict(type=wl["type"], quality=a.quality)
    bw, bh, bbytes = cfx.block_info(a.format)

    # this rank's slab of every layer (layers differ by seed)
    r0, r1, y0, y1 = cfx.shard_block_rows(size, bh, rank, world)
    layers = max(world, 1)
    slab_rows = y1 - y0
    np_dtype = np.uint8 if wl["src"] == "RGBA8" else np.float16
    host = torch.empty((layers, slab_rows, size, 4), dtype=torch.uint8 if np_dtype == np.uint8 else torch.float16,
                       pin_memory=True)
    for l in range(layers):
        img = synth.gen_image(wl["kind"], size, size, seed=12345 + l, rows=(y0, y1))
        host[l] = torch.from_numpy(synth.to_rgba8(img) if np_dtype == np.uint8 else img.astype(np.float16))
    d_src = host.to(dev)
    slab_bytes = (r1 - r0) * ((size + bw - 1) // bw) * bbytes
    d_out = torch.empty((layers, slab_bytes), dtype=torch.uint8, device=dev)
    gathered = [torch.empty_like(d_out) for _ in range(world)] if (world > 1 and rank == 0) else None
    host_out = torch.empty((layers, slab_bytes), dtype=torch.uint8, pin_memory=True)
    texels_per_rank = layers * slab_rows * size
    total_texels = layers * size * size

    def device_step(kernel_events=None):
        for l in range(layers):
            if kernel_events is not None:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            cfx.encode_device(d_src[l], a.format, out=d_out[l], **kw)
            if kernel_events is not None:
                e1.record()
                kernel_events.append((e0, e1))
        if world > 1:
            dist.gather(d_out, gathered, dst=0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        device_step()
    barrier()
    launches0 = cfx.kernel_launches()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    kev = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(a.steps):
        device_step(kev)
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = cfx.kernel_launches() - launches0
    kernel_ms = float(np.mean([x.elapsed_time(y) for x, y in kev]))
    kernel_texels = slab_rows * size

    # ---- end to end through cfx_encode with host buffers
    hsrc = host.numpy()
    hout = host_out.numpy()

    def e2e_step():
        for l in range(layers):
            cfx.encode(hsrc[l], a.format, out=hout[l], **kw)

    for _ in range(max(a.warmup, 3)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([dev_ms, e2e_s * 1e3, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, kernel_ms = [float(x) for x in t.tolist()]
    lt = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(lt)
    launches = int(lt.item())

    if rank == 0:
        peak, how = peaks()
        ms_per_step = dev_ms / a.steps
        value = total_texels / (ms_per_step * 1e-3) / 1e6
        e2e_value = total_texels / (e2e_ms / a.steps * 1e-3) / 1e6
        bpt = wl["read"] + wl["write"]
        achieved = kernel_texels * bpt / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic(a.format, kernel_texels) if a.quality == "Normal" else (None, None)
        out = {"metric": "Mtexels/s encode", "value": value, "unit": "Mtexels/s", "n_gpus": world,
               "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "u8" if wl["src"] == "RGBA8" else "f16", "data": "synthetic", "config": config,
               "e2e": {"value": e2e_value, "unit": "Mtexels/s",
                       "h2d_bytes_per_step": int(host.numel() * host.element_size()) * world,
                       "d2h_bytes_per_step": int(host_out.numel()) * world},
               "gpu_launches": launches,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                            "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write)",
                            "traffic_source": traffic_src, "algorithmic_bytes": int(kernel_texels * bpt), "peak_source": how,
                            "kernel": "%s encode kernel, %.3f ms per launch over %d texels, %.3f B/texel" %
                                      (a.format, kernel_ms, kernel_texels, bpt)},
               "clocks": clocks}
        if world == 1 and not a.no_cpu:
            out["cpu_baseline"], _ = cpu_reference(a, wl, size, 1, 1)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def bench_mips(a, wl, size, rank, world, local, config):
    """BASELINE config 5 shape: one texture WITH its full mip chain per rank (box-filtered levels of generator G --
    the reference builds them on the host with FreeImage, outside this path), level by level through the encoder."""
    import torch
    import torch.distributed as dist
    import cuttlefish_b200 as cfx
    from cuttlefish_b200 import synth
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfx.init(local)
    kw = dict(type=wl["type"], quality=a.quality)
    img = synth.gen_image(wl["kind"], size, size, seed=12345 + rank)
    if a.mipgen:
        return bench_mipgen(a, wl, size, rank, world, local, config, img, kw)
    levels = []
    while True:
        levels.append(np.ascontiguousarray(synth.to_rgba8(img) if wl["src"] == "RGBA8" else img.astype(np.float16)))
        if img.shape[0] == 1 and img.shape[1] == 1:
            break
        h2, w2 = max(img.shape[0]//2, 1), max(img.shape[1]//2, 1)
        img = img[:h2*2, :w2*2].reshape(h2, img.shape[0]//h2, w2, img.shape[1]//w2, 4).mean(axis=(1, 3)).astype(np.float32)
    host = [torch.from_numpy(l).pin_memory() for l in levels]
    d_src = [h.to(dev) for h in host]
    d_out = [torch.empty(cfx.encoded_size(a.format, l.shape[1], l.shape[0]), dtype=torch.uint8, device=dev) for l in levels]
    texels = sum(l.shape[0]*l.shape[1] for l in levels)
    out_bytes = sum(int(o.numel()) for o in d_out)
    gathered = [torch.empty(out_bytes, dtype=torch.uint8, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None

    def device_step():
        for s_, o_ in zip(d_src, d_out):
            cfx.encode_device(s_, a.format, out=o_, **kw)
        if world > 1:
            dist.gather(torch.cat(d_out), gathered, dst=0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        device_step()
    barrier()
    l0 = cfx.kernel_launches()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(a.steps):
        device_step()
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = cfx.kernel_launches() - l0
    hlevels = [h.numpy() for h in host]
    for _ in range(max(a.warmup, 3)):
        cfx.encode_batch(hlevels, a.format, **kw)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cfx.encode_batch(hlevels, a.format, **kw)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0)*1e3
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    lt = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt)
    if rank == 0:
        dev_ms, e2e_ms = [float(x) for x in t.tolist()]
        peak, how = peaks()
        ms = dev_ms/a.steps
        config = dict(config, mip_levels=len(levels), layers=world, workload=config["workload"] + " + full mip chain (%d levels, box filter)" % len(levels),
                      sharding="one texture with its chain per rank",
                      l2="base level (%d MiB) larger than L2; the tail levels are launch bound" % (levels[0].nbytes >> 20))
        bpt = wl["read"] + wl["write"]
        print(json.dumps({"metric": "Mtexels/s encode", "value": world*texels/(ms*1e-3)/1e6, "unit": "Mtexels/s", "n_gpus": world,
                          "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "u8" if wl["src"] == "RGBA8" else "f16", "data": "synthetic", "config": config,
                          "e2e": {"value": world*texels/(e2e_ms/a.steps*1e-3)/1e6, "unit": "Mtexels/s",
                                  "h2d_bytes_per_step": int(sum(l.nbytes for l in levels))*world, "d2h_bytes_per_step": out_bytes*world},
                          "gpu_launches": int(lt.item()),
                          "roofline": {"bound": "hbm", "achieved": texels*bpt/(ms*1e-3)/1e9, "peak": peak, "unit": "GB/s",
                                       "frac": texels*bpt/(ms*1e-3)/1e9/peak, "traffic": None, "peak_source": how,
                                       "kernel": "%s encode kernels of one mip chain (%d launches, step time)" % (a.format, len(levels))},
                          "clocks": clocks}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def bench_mipgen(a, wl, size, rank, world, local, config, img, kw):
    """Texture::generateMipmaps(CatmullRom) + convert per rank, the chain generated on the GPU: `value` with the float32
    level 0 resident (cfx_encode_mip_chain_device), `e2e` from a host level 0 (cfx_encode_mip_chain, 16 B/texel H2D)."""
    import torch
    import torch.distributed as dist
    import cuttlefish_b200 as cfx
    dev = torch.device("cuda", local)
    from cuttlefish_b200 import synth
    # an 8-bit workload goes up as 8-bit texels: the library takes them as v/255, Image::convert(RGBAF) of an 8-bit image
    img = np.ascontiguousarray(synth.to_rgba8(img) if wl["src"] == "RGBA8" else img.astype(np.float32))
    texel0 = float(img.dtype.itemsize*4)
    host = torch.from_numpy(img).pin_memory()
    d_src = host.to(dev)
    sizes = [(max(1, size >> k), max(1, size >> k)) for k in range(cfx.mip_levels(size, size))]
    d_out = [torch.empty(cfx.encoded_size(a.format, w, h), dtype=torch.uint8, device=dev) for (w, h) in sizes]
    texels = sum(w*h for (w, h) in sizes)
    out_bytes = sum(int(o.numel()) for o in d_out)
    gathered = [torch.empty(out_bytes, dtype=torch.uint8, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None

    def device_step():
        cfx.encode_mip_chain_device(d_src, a.format, "CatmullRom", outs=d_out, **kw)
        if world > 1:
            dist.gather(torch.cat(d_out), gathered, dst=0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        device_step()
    barrier()
    l0 = cfx.kernel_launches()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(a.steps):
        device_step()
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = cfx.kernel_launches() - l0
    himg = host.numpy()
    houts = [torch.empty(int(o.numel()), dtype=torch.uint8).pin_memory().numpy() for o in d_out]
    for _ in range(max(a.warmup, 3)):
        cfx.encode_mip_chain(himg, a.format, "CatmullRom", outs=houts, **kw)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cfx.encode_mip_chain(himg, a.format, "CatmullRom", outs=houts, **kw)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0)*1e3
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    lt = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt)
    if rank == 0:
        dev_ms, e2e_ms = [float(x) for x in t.tolist()]
        peak, how = peaks()
        ms = dev_ms/a.steps
        # algorithmic bytes of one chain: every level read once by its encoder (16 B/texel) and written as blocks; every
        # resize reads the level above, writes and re-reads the x-filtered intermediate, and writes the level (16 B each)
        alg = 0.0
        for k, (w, h) in enumerate(sizes):
            alg += w*h*((16.0 if k else texel0) + wl["write"])
            if k:
                pw, ph = sizes[k - 1]
                alg += (16.0 if k > 1 else texel0)*pw*ph + 16.0*(2*w*ph + w*h)
        config = dict(config, mip_levels=len(sizes), layers=world,
                      workload=config["workload"] + " + generateMipmaps(CatmullRom) on the GPU + full mip chain (%d levels)" % len(sizes),
                      sharding="one texture with its chain per rank",
                      l2="level 0 (%d MiB) larger than L2; the tail levels are launch bound" % (img.nbytes >> 20))
        out = {"metric": "Mtexels/s encode", "value": world*texels/(ms*1e-3)/1e6, "unit": "Mtexels/s", "n_gpus": world,
               "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64 filter / u8 encode", "data": "synthetic", "config": config,
               "e2e": {"value": world*texels/(e2e_ms/a.steps*1e-3)/1e6, "unit": "Mtexels/s",
                       "h2d_bytes_per_step": int(img.nbytes)*world, "d2h_bytes_per_step": out_bytes*world},
               "gpu_launches": int(lt.item()),
               "roofline": {"bound": "hbm", "achieved": alg/(ms*1e-3)/1e9, "peak": peak, "unit": "GB/s",
                            "frac": alg/(ms*1e-3)/1e9/peak, "traffic": None, "peak_source": how, "algorithmic_bytes": alg,
                            "kernel": "resize passes + %s encode kernels of one mip chain (%d launches, step time)" %
                                      (a.format, int(lt.item())//max(a.steps, 1)//max(world, 1))},
               "clocks": clocks}
        if not a.no_cpu and world == 1:
            out["cpu_baseline"], _ = cpu_reference_mips(a, wl, size, 1, 1)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
