"""World-size-2 gloo test of the multi-GPU host logic (runs on CPU): block-row sharding, per-rank
slab encode, gather to rank 0, concatenation == whole-surface encode.  The per-rank encoder here is
the CPU oracle (test infrastructure) standing in for cfx_encode_device, which needs a GPU; what is
under test is shard_block_rows + the gather plumbing bench.py uses."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, fmt, w, h, result_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cuttlefish_b200 as cfx
    import oracle
    from cuttlefish_b200 import synth
    bw, bh, nbytes = cfx.block_info(fmt)
    r0, r1, y0, y1 = cfx.shard_block_rows(h, bh, rank, world)
    slab = synth.gen_image("noise+grad", w, h, seed=5, rows=(y0, y1))          # only this rank's rows
    mine = torch.from_numpy(oracle.encode(slab, fmt)) if r1 > r0 else torch.empty(0, dtype=torch.uint8)
    assert mine.numel() == (r1 - r0) * ((w + bw - 1) // bw) * nbytes
    # ragged gather: sizes first, then padded payloads
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([mine.numel()], dtype=torch.int64))
    cap = int(max(s.item() for s in sizes))
    padded = torch.zeros(cap, dtype=torch.uint8)
    padded[:mine.numel()] = mine
    parts = [torch.zeros(cap, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, parts, dst=0)
    if rank == 0:
        whole = torch.cat([p[:int(s.item())] for p, s in zip(parts, sizes)]).numpy()
        np.save(result_path, whole)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fmt,w,h", [("BC7", 64, 40), ("ASTC_6x6", 50, 34), ("BC1_RGB", 32, 6)])
def test_two_rank_sharded_encode_matches_whole(tmp_path, fmt, w, h):
    sys.path.insert(0, ROOT)
    import oracle
    from cuttlefish_b200 import synth
    if not oracle.available():
        pytest.skip("oracle library not built")
    port = 29500 + (os.getpid() + hash(fmt)) % 2000
    out = str(tmp_path / "whole.npy")
    mp.spawn(_worker, args=(2, port, fmt, w, h, out), nprocs=2, join=True)
    whole = np.load(out)
    ref = oracle.encode(synth.gen_image("noise+grad", w, h, seed=5), fmt)
    assert np.array_equal(whole, ref)
