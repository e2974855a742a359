#!/usr/bin/env python
"""bench.py -- Mtexels/s of the block-encode hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--format BC7] [--impl reference]

A step = one pass of the encoder over one batch of synthetic input (generator G).
The headline workload is BASELINE.json configs[1]: BC7 UNorm, 8192x8192 RGBA8 noise+grad, quality
Normal; without --format the same JSON line also carries, under "secondary", the other GPU configs
of BASELINE.json: ASTC 6x6 8192^2 (config 3), BC6H UF16 4096^2 hdr ramp (config 4) and ETC2 RGBA8
4096^2 with its full mip chain generated on the GPU (config 5).

Per workload:
  value     device-timed (CUDA events, max over ranks), inputs resident in HBM. At N ranks the batch
            is an N-layer array texture (Converter::convert's depth/face loop,
            lib/src/Converter.cpp:521-527); every layer is sharded by block row across the ranks
            (SURVEY.md 8e), so per-GPU work is fixed ("weak"); the assembled output lives in rank 0's HBM
            and every rank's encode kernel stores its packed blocks straight into it over NVLink (peer
            memory, include/cfx.h cfx_ipc_*): the gather is fused into the kernel. --gather nccl uses one
            NCCL gather per layer instead, overlapped with the next layer's kernels.
  scaling_strong   ONE image of the same size over the N ranks (BASELINE config 3's shape), same gather.
  e2e       the same batch through ONE cfx_encode_batch() call with HOST buffers (pinned), issued by
            rank 0 alone with libcfx's device pool set to the N GPUs: the library shards every layer by
            block row, each GPU uploads its rows, encodes, and writes its blocks straight into the one
            caller-owned output. H2D, kernels, D2H all inside the timed region. This is what a C++
            Cuttlefish host gets from Converter::convert() through adapter/CudaConverter.cpp.
  e2e_rgbaf the drop-in shape: ONE image as PAGEABLE RGBA32F (what cuttlefish::Image holds) -> blocks in
            pageable memory, same call.
  roofline  algorithmic bytes (source read once + blocks written once) / kernel duration, vs the
            measured HBM copy peak in MEASURED_PEAKS.json.
  cpu_baseline / psnr   (N=1 only) the reference's own CPU encoder (oracle/_ref, built from
            /root/reference) on a bounded crop of the same image, all host threads, and the RGB PSNR of
            both encoders' blocks of that crop against the source, decoded by the reference's decoders.
            The ONLY use of oracle/ here.
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

PSNR_TOLERANCE_DB = 0.1         # BASELINE.json north_star

# bytes per texel read / written; crop = side of the CPU-baseline / PSNR sample
WORKLOADS = {
    "BC7": dict(size=8192, kind="noise+grad", src="RGBA8", type="UNorm", read=4.0, write=1.0, crop=4096, peak=1.0),
    "BC1_RGB": dict(size=8192, kind="noise+grad", src="RGBA8", type="UNorm", read=4.0, write=0.5, crop=4096, peak=1.0),
    "BC3": dict(size=8192, kind="noise+grad", src="RGBA8", type="UNorm", read=4.0, write=1.0, crop=4096, peak=1.0),
    "BC4": dict(size=8192, kind="noise+grad", src="RGBA8", type="UNorm", read=4.0, write=0.5, crop=2048, peak=1.0),
    "BC5": dict(size=8192, kind="noise+grad", src="RGBA8", type="UNorm", read=4.0, write=1.0, crop=2048, peak=1.0),
    "ETC1": dict(size=4096, kind="noise+grad", src="RGBA8", type="UNorm", read=4.0, write=0.5, crop=2048, peak=1.0),
    "ETC2_R8G8B8": dict(size=4096, kind="noise+grad", src="RGBA8", type="UNorm", read=4.0, write=0.5, crop=2048, peak=1.0),
    "ETC2_R8G8B8A8": dict(size=4096, kind="noise+grad", src="RGBA8", type="UNorm", read=4.0, write=1.0, crop=2048, peak=1.0),
    "BC6H": dict(size=4096, kind="hdr", src="RGBA16F", type="UFloat", read=8.0, write=1.0, crop=2048, peak=64.0),
    "ASTC_6x6": dict(size=8192, kind="noise+grad", src="RGBA8", type="UNorm", read=4.0, write=16.0 / 36.0, crop=2046, peak=1.0),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--format", default=None, help="bench only this format (default: BC7 + the secondary workloads)")
    ap.add_argument("--quality", default="Normal")
    ap.add_argument("--size", type=int, default=0, help="override the square image size")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="override the side of the CPU baseline crop")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how the packed blocks reach rank 0 (peer: stored by the encode kernels over NVLink; nccl: "
                         "one NCCL gather per layer)")
    ap.add_argument("--mips", action="store_true",
                    help="with --format: encode the image WITH its full mip chain generated on the GPU "
                         "(Texture::generateMipmaps(CatmullRom) + convert, BASELINE config 5's shape), one texture per rank")
    return ap.parse_args()


def measured_traffic(fmt, texels):
    """DRAM bytes of one launch from the committed ncu captures (profiles/r0N_dram_traffic.json, newest first), scaled
    to the texels of the launch timed here; None when no capture exists for the format."""
    for name in ("r02_dram_traffic.json", "r01_dram_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)[fmt]
            return int((t["dram_read_bytes"] + t["dram_write_bytes"]) * (texels / t["texels"])), t["csv"]
        except Exception:
            continue
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
        # "under load" = the upper half of the samples (idle samples before/after the loops drop out)
        load = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- the CPU side: reference encoders on a bounded crop (oracle/ -- the checker and the baseline, never the product) ----

def crop_side(a, wl, size):
    return min(a.cpu_sample or wl["crop"], size)


def cpu_reference(a, fmt, wl, size, steps, warmup):
    """Times the reference's CPU encoders (oracle/_ref) on a bounded crop. Returns (baseline dict, seconds per step,
    the crop as float32, the reference's blocks of it)."""
    import oracle  # the checker / CPU baseline: never on the product path
    from cuttlefish_b200 import synth
    s = crop_side(a, wl, size)
    img = synth.gen_image(wl["kind"], size, size, rows=(0, s))[:, :s].copy()
    if wl["src"] == "RGBA16F":
        img = img.astype(np.float16).astype(np.float32)
    threads = oracle.hardware_threads()
    kw = dict(type=wl["type"], quality=a.quality)
    # the reference's own Converter::convert (std::thread pool, real converter glue, real encoders) when
    # oracle/_ref/libcfglue.so was built; else the same encoders under our byte-identical glue restatement
    enc = oracle.encode_glue if oracle.glue_available() else oracle.encode
    for _ in range(warmup):
        enc(img[: max(64, s // 8)], fmt, threads=0, **kw)
    times, blocks = [], None
    for _ in range(steps):
        t = time.perf_counter()
        blocks = enc(img, fmt, threads=0, **kw)
        times.append(time.perf_counter() - t)
    dt = float(np.mean(times))
    return {"value": s * s / dt / 1e6, "unit": "Mtexels/s", "cores": threads, "kind": "reference",
            "sample": "%dx%d top-left crop of the %dx%d %s image, %s quality=%s, oracle/_ref "
                      "(reference Converter::convert + encoders compiled from source), %d threads, %.2f s/step" %
                      (s, s, size, size, wl["kind"], fmt, a.quality, threads, dt)}, dt, img, blocks


def psnr_pair(a, fmt, wl, img, ref_blocks, gpu_blocks):
    """RGB PSNR of both encoders' blocks of the crop against the source, decoded by the reference's decoders."""
    import oracle
    h, w, _ = img.shape
    kw = dict(type=wl["type"])
    p_gpu = oracle.psnr_rgb(img, oracle.decode(gpu_blocks, fmt, w, h, **kw), wl["peak"])
    p_ref = oracle.psnr_rgb(img, oracle.decode(ref_blocks, fmt, w, h, **kw), wl["peak"])
    ok = bool(p_gpu >= p_ref - PSNR_TOLERANCE_DB)
    if not ok:
        sys.stderr.write("bench: %s PSNR %.3f dB is more than %.1f dB under the reference's %.3f dB\n" %
                         (fmt, p_gpu, PSNR_TOLERANCE_DB, p_ref))
    return {"gpu": p_gpu, "reference": p_ref, "delta_db": p_gpu - p_ref, "tolerance_db": -PSNR_TOLERANCE_DB, "ok": ok,
            "sample": "RGB PSNR (peak %g) over the %dx%d CPU-baseline crop; both outputs decoded by the reference's decoder" %
                      (wl["peak"], w, h)}


def cpu_reference_mips(a, fmt, wl, size, steps, warmup):
    """Times the reference's Texture::generateMipmaps + convert on a bounded level 0: the real FreeImage_Rescale chain
    (oracle/_ref/libfiresize.so; single-threaded, as in the reference) and the real Converter::convert per level."""
    import oracle
    from oracle import resize as oresize
    from cuttlefish_b200 import synth
    s = min((a.cpu_sample or wl["crop"]), size)
    img = synth.gen_image(wl["kind"], size, size, rows=(0, s))[:, :s].copy().astype(np.float32)
    threads = oracle.hardware_threads()
    kw = dict(type=wl["type"], quality=a.quality)
    enc = oracle.encode_glue if oracle.glue_available() else oracle.encode
    for _ in range(warmup):
        enc(img[:64], fmt, threads=0, **kw)
    times, resize_times, levels, blocks0 = [], [], None, None
    for _ in range(steps):
        t = time.perf_counter()
        levels = oresize.mip_chain(img, "CatmullRom", fn=oresize.resize_ref)
        t_mid = time.perf_counter()
        for k, l in enumerate(levels):
            b = enc(l, fmt, threads=0, **kw)
            if k == 0:
                blocks0 = b
        times.append(time.perf_counter() - t)
        resize_times.append(t_mid - t)
    dt = float(np.mean(times))
    texels = sum(l.shape[0] * l.shape[1] for l in levels)
    return {"value": texels / dt / 1e6, "unit": "Mtexels/s", "cores": threads, "kind": "reference",
            "sample": "%dx%d top-left crop of the %dx%d %s image as level 0, %d-level Catmull-Rom chain by the reference's "
                      "FreeImage_Rescale (1 thread, %.2f s) + %s quality=%s Converter::convert per level (%d threads), "
                      "%.2f s/step" % (s, s, size, size, wl["kind"], len(levels), float(np.mean(resize_times)), fmt,
                                       a.quality, threads, dt)}, dt, img, blocks0


def workload_name(a, fmt, wl, size, mips=False):
    s = "%s %s encode, %dx%d %s synthetic %s (generator G), quality=%s" % (fmt, wl["type"], size, size, wl["src"], wl["kind"], a.quality)
    return s + (" + generateMipmaps(CatmullRom) + full mip chain" if mips else "")


def reference_arm(a):
    """bench.py --impl reference: the reference's own CPU implementation of the path on the box's host cores, each step a
    bounded sample (crop) of the workload; rank 0 alone."""
    fmt = a.format or "BC7"
    wl = dict(WORKLOADS[fmt])
    size = a.size or wl["size"]

    def line(fmt, wl, size, steps, mips):
        if mips:
            base, dt, _, _ = cpu_reference_mips(a, fmt, wl, size, steps, min(a.warmup, 1))
        else:
            base, dt, _, _ = cpu_reference(a, fmt, wl, size, steps, min(a.warmup, 1))
        s = crop_side(a, wl, size)
        return {"metric": "Mtexels/s encode", "value": base["value"], "unit": "Mtexels/s", "steps": steps,
                "ms_per_step": dt * 1e3, "config": {"workload": workload_name(a, fmt, wl, size, mips), "format": fmt,
                                                    "quality": a.quality, "width": size, "height": size,
                                                    "timed_sample": "each step encodes the %dx%d top-left crop of the image "
                                                                    "(a rate, not the whole image's time)" % (s, s)},
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "Mtexels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}

    out = line(fmt, wl, size, a.steps, a.mips)
    out.update({"impl": "reference", "n_gpus": a.gpus, "warmup": a.warmup, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic"})
    if a.format is None and not a.no_secondary:
        sec = []
        for f2, mips in (("ASTC_6x6", False), ("BC6H", False), ("ETC2_R8G8B8A8", True)):
            w2 = dict(WORKLOADS[f2])
            sec.append(line(f2, w2, w2["size"], min(a.steps, 3), mips))
        out["secondary"] = sec
    print(json.dumps(out))
    return 0


# ---- the GPU side ---------------------------------------------------------------------------------------------------

class Env:
    """Process-group plumbing shared by the workloads."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.cpu_group = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            # host-side barriers for the phases in which rank 0 alone drives every GPU: an NCCL barrier would park a
            # spinning kernel on the other ranks' GPUs
            self.cpu_group = dist.new_group(backend="gloo")
        self.images = {}

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def cpu_barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.cpu_group)

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def sum_over_ranks(self, value):
        t = self.torch.tensor([value], dtype=self.torch.int64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t)
        return int(t.item())

    def image_rows(self, kind, size, y0, y1):
        """Rows [y0, y1) of generator G's image (cached: BC7 and ASTC share the 8192^2 one)."""
        from cuttlefish_b200 import synth
        key = (kind, size, y0, y1)
        if key not in self.images:
            self.images[key] = synth.gen_image(kind, size, size, rows=(y0, y1))
        return self.images[key]

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def to_src(img, wl):
    from cuttlefish_b200 import synth
    return synth.to_rgba8(img) if wl["src"] == "RGBA8" else img.astype(np.float16)


def layer_variant(arr, l):
    """Layer l of the array texture: layer 0 with its columns rotated (distinct content per layer, no second pass of
    the generator)."""
    return arr if l == 0 else np.roll(arr, l * 64, axis=1)


def time_device(env, fn, steps, warmup):
    """W warm-up calls, then K timed ones bracketed by barrier + synchronize; returns the max over ranks in ms."""
    torch = env.torch
    for _ in range(warmup):
        fn(None)
    env.barrier()
    kev = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    ev0.record()
    for _ in range(steps):
        fn(kev)
    ev1.record()
    env.barrier()
    kernel_ms = float(np.mean([x.elapsed_time(y) for x, y in kev])) if kev else 0.0
    ms, kernel_ms = env.max_over_ranks([ev0.elapsed_time(ev1), kernel_ms])
    return ms, kernel_ms


def time_host_call(env, fn, steps, warmup):
    """Rank 0 alone times `fn` (a blocking host-buffer library call) with the wall clock; the other ranks wait on the host."""
    env.cpu_barrier()
    dt = 0.0
    if env.rank == 0:
        for _ in range(warmup):
            fn()
        env.torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        dt = time.perf_counter() - t0
    env.cpu_barrier()
    return dt * 1e3


def run_encode(a, env, fmt, steps, warmup, want_cpu):
    """One square single-surface workload: weak + strong device-timed, e2e, e2e_rgbaf, roofline, CPU baseline, PSNR."""
    import cuttlefish_b200 as cfx
    torch, dist = env.torch, env.dist
    rank, world, dev = env.rank, env.world, env.dev
    wl = dict(WORKLOADS[fmt])
    size = a.size or wl["size"]
    kw = dict(type=wl["type"], quality=a.quality)
    bw, bh, bbytes = cfx.block_info(fmt)
    cfx.init(env.local)

    # this rank's slab of every layer
    r0, r1, y0, y1 = cfx.shard_block_rows(size, bh, rank, world)
    layers = max(world, 1)
    slab_rows = y1 - y0
    slab0 = to_src(env.image_rows(wl["kind"], size, y0, y1), wl)
    tdtype = torch.uint8 if slab0.dtype == np.uint8 else torch.float16
    host = torch.empty((layers, slab_rows, size, 4), dtype=tdtype, pin_memory=True)
    for l in range(layers):
        host[l] = torch.from_numpy(layer_variant(slab0, l))
    d_src = host.to(dev)
    del host
    blocks_x = (size + bw - 1) // bw
    slab_bytes = (r1 - r0) * blocks_x * bbytes
    layer_bytes = ((size + bh - 1) // bh) * blocks_x * bbytes
    my_off = r0 * blocks_x * bbytes
    # The assembled output (every layer, every rank's slab at its byte offset) lives in rank 0's HBM. Default gather:
    # rank 0 exports the buffer, the other ranks map it (cfx_ipc_open) and their encode kernels store their packed
    # blocks STRAIGHT into it over NVLink -- the gather is fused into the kernel; one tiny all-reduce per step is the
    # stream-ordered "every slab has landed" signal. --gather nccl: blocks go to a local buffer and one NCCL gather per
    # layer (overlapping the next layer's kernels) moves them.
    whole = torch.empty((layers, layer_bytes), dtype=torch.uint8, device=dev) if rank == 0 else None
    d_out = torch.empty((layers, slab_bytes), dtype=torch.uint8, device=dev) if (world == 1 or a.gather == "nccl") else None
    base_ptr, handle, flag = None, None, None
    max_rows = max(cfx.shard_block_rows(size, bh, r, world)[1] - cfx.shard_block_rows(size, bh, r, world)[0] for r in range(world))
    pad_bytes = max_rows * blocks_x * bbytes

    def nccl_gather_layer0():
        """Layer 0 through a local buffer + NCCL gather (slab sizes differ by at most one block row: pad to the largest)."""
        send = torch.zeros(pad_bytes, dtype=torch.uint8, device=dev)
        cfx.encode_device(d_src[0], fmt, out=send[:slab_bytes], **kw)
        parts = [torch.empty(pad_bytes, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
        dist.gather(send, parts, dst=0)
        if rank != 0:
            return None
        sizes = [(cfx.shard_block_rows(size, bh, r, world)[1] - cfx.shard_block_rows(size, bh, r, world)[0]) * blocks_x * bbytes
                 for r in range(world)]
        return torch.cat([p_[:n_] for p_, n_ in zip(parts, sizes)])

    if world > 1 and a.gather == "peer":
        handles = [cfx.ipc_export(whole) if rank == 0 else None]
        dist.broadcast_object_list(handles, src=0)
        handle = handles[0]
        base_ptr = whole.data_ptr() if rank == 0 else cfx.ipc_open(handle, env.local)
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
    d_send = torch.empty((layers, pad_bytes), dtype=torch.uint8, device=dev) if (world > 1 and a.gather == "nccl") else None
    gathered = [[torch.empty(pad_bytes, dtype=torch.uint8, device=dev) for _ in range(world)] for _ in range(layers)] \
        if (world > 1 and a.gather == "nccl" and rank == 0) else None

    def device_step(n_layers, kev):
        works = []
        for l in range(n_layers):
            if kev is not None:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            if world == 1:
                cfx.encode_device(d_src[l], fmt, out=d_out[l], **kw)
            elif a.gather == "peer":
                cfx.encode_device(d_src[l], fmt, out_ptr=base_ptr + l * layer_bytes + my_off, out_bytes=slab_bytes, **kw)
            else:
                cfx.encode_device(d_src[l], fmt, out=d_send[l][:slab_bytes], **kw)
            if kev is not None:
                e1.record()
                kev.append((e0, e1))
            if world > 1 and a.gather == "nccl":
                # layer l's gather runs on NCCL's stream beside layer l+1's kernels
                works.append(dist.gather(d_send[l], gathered[l] if rank == 0 else None, dst=0, async_op=True))
        for w in works:
            w.wait()
        if world > 1 and a.gather == "peer":
            dist.all_reduce(flag)          # stream-ordered after this rank's kernels: done everywhere = all slabs landed

    gather_checked = None
    if world > 1 and a.gather == "peer":
        # once, before timing: the fused peer-store gather must give the bytes of the NCCL gather
        device_step(1, None)
        env.barrier()
        want = nccl_gather_layer0()
        env.barrier()
        if rank == 0:
            gather_checked = bool(torch.equal(whole[0], want))
            assert gather_checked, "peer-store gather differs from the NCCL gather"
        del want

    launches0 = cfx.kernel_launches()
    weak_ms, kernel_ms = time_device(env, lambda kev: device_step(layers, kev), steps, max(warmup, 3))
    launches = env.sum_over_ranks(cfx.kernel_launches() - launches0) * steps // (steps + max(warmup, 3))
    strong_ms, _ = time_device(env, lambda kev: device_step(1, None), steps, max(warmup, 3)) if world > 1 else (weak_ms, 0.0)
    if world > 1 and a.gather == "peer" and rank != 0:
        cfx.ipc_close(base_ptr, handle)
    env.barrier()
    del gathered, d_send, whole
    torch.cuda.empty_cache()

    # ---- end to end: ONE host-buffer call from rank 0, the library's device pool = all N GPUs
    e2e_ms = e2e_strong_ms = rgbaf_ms = 0.0
    h2d = d2h = rgbaf_bytes = 0
    n_rgbaf = min(steps, 5)
    gpu_crop_blocks = None
    if rank == 0:
        cfx.init_devices(world)
        full0 = to_src(env.image_rows(wl["kind"], size, 0, size), wl)
        hsrc = torch.empty((layers, size, size, 4), dtype=tdtype, pin_memory=True)
        for l in range(layers):
            hsrc[l] = torch.from_numpy(layer_variant(full0, l))
        out_bytes = cfx.encoded_size(fmt, size, size)
        hout = torch.empty((layers, out_bytes), dtype=torch.uint8, pin_memory=True)
        srcs = [hsrc[l].numpy() for l in range(layers)]
        outs = [hout[l].numpy() for l in range(layers)]
        h2d, d2h = int(hsrc.numel() * hsrc.element_size()), int(hout.numel())
    e2e_ms = time_host_call(env, lambda: cfx.encode_batch(srcs, fmt, outs=outs, **kw), steps, max(warmup, 3))
    if world > 1:
        e2e_strong_ms = time_host_call(env, lambda: cfx.encode(srcs[0], fmt, out=outs[0], **kw), steps, max(warmup, 3))
    if rank == 0:
        pageable = np.ascontiguousarray(full0.astype(np.float32) / np.float32(255.0)) if full0.dtype == np.uint8 \
            else full0.astype(np.float32)
        pout = np.empty(out_bytes, np.uint8)
        rgbaf_bytes = int(pageable.nbytes)
    rgbaf_ms = time_host_call(env, lambda: cfx.encode(pageable, fmt, out=pout, **kw), n_rgbaf, 2)
    if rank == 0:
        if want_cpu and world == 1:
            s = crop_side(a, wl, size)
            gpu_crop_blocks = cfx.encode(np.ascontiguousarray(full0[:s, :s]), fmt, **kw).copy()
        del hsrc, hout, srcs, outs, pageable, pout
        cfx.init(env.local)

    if rank != 0:
        return None
    peak, how = peaks()
    total_texels = layers * size * size
    ms_per_step = weak_ms / steps
    bpt = wl["read"] + wl["write"]
    kernel_texels = slab_rows * size
    achieved = kernel_texels * bpt / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic(fmt, kernel_texels) if a.quality == "Normal" else (None, None)
    out = {"metric": "Mtexels/s encode", "value": total_texels / (ms_per_step * 1e-3) / 1e6, "unit": "Mtexels/s",
           "n_gpus": world, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms_per_step,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u8" if wl["src"] == "RGBA8" else "f16", "data": "synthetic",
           "config": {"workload": workload_name(a, fmt, wl, size), "format": fmt, "quality": a.quality, "width": size,
                      "height": size, "layers": layers,
                      "sharding": "block-row slabs of every layer across %d rank(s)" % world,
                      "gather": "none (one rank)" if world == 1 else
                                ("fused into the encode kernel: every rank's kernel stores its packed blocks straight into rank "
                                 "0's HBM over NVLink (cfx_ipc_open peer memory), one tiny all-reduce per step as the completion "
                                 "signal; checked once against the NCCL gather: %s" % gather_checked) if a.gather == "peer" else
                                "one NCCL gather per layer, overlapped with the next layer's kernels",
                      "l2": "inputs (%d MiB per rank) larger than L2; no flush needed" % (int(d_src.numel() * d_src.element_size()) >> 20)},
           "e2e": {"value": total_texels / (e2e_ms / steps * 1e-3) / 1e6, "unit": "Mtexels/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / steps,
                   "how": "one cfx_encode_batch() call per step from one process over %d pinned %s layer(s); libcfx shards "
                          "each layer by block row over its pool of %d GPU(s) and every GPU writes its blocks straight into "
                          "the caller's output buffer" % (layers, wl["src"], world)},
           "e2e_rgbaf": {"value": size * size / (rgbaf_ms / n_rgbaf * 1e-3) / 1e6, "unit": "Mtexels/s", "steps": n_rgbaf,
                         "ms_per_step": rgbaf_ms / n_rgbaf, "h2d_source_bytes": rgbaf_bytes,
                         "how": "ONE %dx%d image as PAGEABLE RGBA32F (what cuttlefish::Image holds; the adapter's call) -> "
                                "blocks in pageable memory, cfx_encode() over %d GPU(s)" % (size, size, world)},
           "gpu_launches": launches,
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write)",
                        "traffic_source": traffic_src, "algorithmic_bytes": int(kernel_texels * bpt), "peak_source": how,
                        "kernel": "%s encode kernel, %.3f ms per launch over %d texels, %.3f B/texel" %
                                  (fmt, kernel_ms, kernel_texels, bpt)}}
    if world > 1:
        out["scaling_strong"] = {"value": size * size / (strong_ms / steps * 1e-3) / 1e6, "unit": "Mtexels/s",
                                 "ms_per_step": strong_ms / steps,
                                 "e2e": {"value": size * size / (e2e_strong_ms / steps * 1e-3) / 1e6, "unit": "Mtexels/s",
                                         "ms_per_step": e2e_strong_ms / steps},
                                 "workload": "ONE %dx%d image over %d GPUs (total work fixed)" % (size, size, world)}
    if want_cpu and world == 1:
        base, _, crop, ref_blocks = cpu_reference(a, fmt, wl, size, 1, 1)
        out["cpu_baseline"] = base
        out["psnr"] = psnr_pair(a, fmt, wl, crop, ref_blocks, gpu_crop_blocks)
    return out


def run_mipgen(a, env, fmt, steps, warmup, want_cpu):
    """Texture::generateMipmaps(CatmullRom) + convert, one texture per rank, the chain generated on the GPU: `value`
    with level 0 resident (cfx_encode_mip_chain_device), `e2e` from a host level 0 (cfx_encode_mip_chain)."""
    import cuttlefish_b200 as cfx
    torch, dist = env.torch, env.dist
    rank, world, dev = env.rank, env.world, env.dev
    wl = dict(WORKLOADS[fmt])
    size = a.size or wl["size"]
    kw = dict(type=wl["type"], quality=a.quality)
    cfx.init(env.local)
    from cuttlefish_b200 import synth
    img = env.image_rows(wl["kind"], size, 0, size)
    # an 8-bit workload goes up as 8-bit texels: the library takes them as v/255, Image::convert(RGBAF) of an 8-bit image
    img = np.ascontiguousarray(layer_variant(synth.to_rgba8(img) if wl["src"] == "RGBA8" else img.astype(np.float32), rank))
    texel0 = float(img.dtype.itemsize * 4)
    host = torch.from_numpy(img).pin_memory()
    d_src = host.to(dev)
    sizes = [(max(1, size >> k), max(1, size >> k)) for k in range(cfx.mip_levels(size, size))]
    d_out = [torch.empty(cfx.encoded_size(fmt, w, h), dtype=torch.uint8, device=dev) for (w, h) in sizes]
    texels = sum(w * h for (w, h) in sizes)
    out_bytes = sum(int(o.numel()) for o in d_out)
    gathered = [torch.empty(out_bytes, dtype=torch.uint8, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None

    def device_step(kev):
        cfx.encode_mip_chain_device(d_src, fmt, "CatmullRom", outs=d_out, **kw)
        if world > 1:
            dist.gather(torch.cat(d_out), gathered, dst=0)

    l0 = cfx.kernel_launches()
    dev_ms, _ = time_device(env, device_step, steps, max(warmup, 3))
    launches = env.sum_over_ranks(cfx.kernel_launches() - l0) * steps // (steps + max(warmup, 3))
    himg = host.numpy()
    houts = [torch.empty(int(o.numel()), dtype=torch.uint8).pin_memory().numpy() for o in d_out]
    # every rank drives its own GPU here (one texture per rank; a chain does not shard)
    for _ in range(max(warmup, 3)):
        cfx.encode_mip_chain(himg, fmt, "CatmullRom", outs=houts, **kw)
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        cfx.encode_mip_chain(himg, fmt, "CatmullRom", outs=houts, **kw)
    torch.cuda.synchronize()
    e2e_ms = env.max_over_ranks([(time.perf_counter() - t0) * 1e3])[0]
    gpu_crop_blocks = None
    if rank == 0 and want_cpu and world == 1:
        s = crop_side(a, wl, size)
        gpu_crop_blocks = cfx.encode(np.ascontiguousarray(img[:s, :s]), fmt, **kw).copy()
    if rank != 0:
        return None
    peak, how = peaks()
    ms = dev_ms / steps
    # algorithmic bytes of one chain: every level read once by its encoder (16 B/texel) and written as blocks; every
    # resize reads the level above, writes and re-reads the x-filtered intermediate, and writes the level (16 B each)
    alg = 0.0
    for k, (w, h) in enumerate(sizes):
        alg += w * h * ((16.0 if k else texel0) + wl["write"])
        if k:
            pw, ph = sizes[k - 1]
            alg += (16.0 if k > 1 else texel0) * pw * ph + 16.0 * (2 * w * ph + w * h)
    out = {"metric": "Mtexels/s encode", "value": world * texels / (ms * 1e-3) / 1e6, "unit": "Mtexels/s", "n_gpus": world,
           "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64 filter / u8 encode", "data": "synthetic",
           "config": {"workload": workload_name(a, fmt, wl, size, True) + " on the GPU (%d levels)" % len(sizes), "format": fmt,
                      "quality": a.quality, "width": size, "height": size, "mip_levels": len(sizes), "layers": world,
                      "sharding": "one texture with its chain per rank",
                      "l2": "level 0 (%d MiB) larger than L2; the tail levels are launch bound" % (img.nbytes >> 20)},
           "e2e": {"value": world * texels / (e2e_ms / steps * 1e-3) / 1e6, "unit": "Mtexels/s",
                   "h2d_bytes_per_step": int(img.nbytes) * world, "d2h_bytes_per_step": out_bytes * world,
                   "ms_per_step": e2e_ms / steps,
                   "how": "one cfx_encode_mip_chain() call per step and rank: pinned %s level 0 in, every level's blocks out" % wl["src"]},
           "gpu_launches": launches,
           "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": how, "algorithmic_bytes": alg,
                        "kernel": "resize passes + %s encode kernels of one mip chain (%d launches, step time)" %
                                  (fmt, launches // max(steps, 1) // max(world, 1))}}
    if want_cpu and world == 1:
        base, _, crop, ref_blocks = cpu_reference_mips(a, fmt, wl, size, 1, 1)
        out["cpu_baseline"] = base
        out["psnr"] = psnr_pair(a, fmt, wl, crop, ref_blocks, gpu_crop_blocks)
        out["psnr"]["sample"] += " (level 0 of the chain)"
    return out


def main():
    a = parse()
    if a.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        return reference_arm(a)

    env = Env()
    sampler = ClockSampler(env.local)
    if env.rank == 0:
        sampler.start()
    fmt = a.format or "BC7"
    want_cpu = not a.no_cpu
    if a.mips:
        out = run_mipgen(a, env, fmt, a.steps, a.warmup, want_cpu)
    else:
        out = run_encode(a, env, fmt, a.steps, a.warmup, want_cpu)
    secondary = []
    if a.format is None and not a.no_secondary and not a.size:
        for f2, mips in (("ASTC_6x6", False), ("BC6H", False), ("ETC2_R8G8B8A8", True)):
            fn = run_mipgen if mips else run_encode
            secondary.append(fn(a, env, f2, a.steps, a.warmup, want_cpu))
    if env.rank == 0:
        out["clocks"] = sampler.stop()
        if secondary:
            out["secondary"] = secondary
        print(json.dumps(out))
    env.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
