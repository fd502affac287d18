"""world_size-2 tests of the N>1 host logic on CPU (gloo): index-blob fan-out and read sharding / merge."""
import os

import numpy as np
import torch.multiprocessing as mp

from winnowmap_b200 import multi


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    blob = rng.integers(0, 256, size=1 << 20, dtype=np.uint8) if rank == 0 else None
    got = multi.broadcast_blob(blob, 0, rank)
    q.put((rank, int(got.sum()), got.nbytes))
    dist.destroy_process_group()


def test_blob_broadcast_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29611, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res[0][1:] == res[1][1:] and res[0][2] == 1 << 20


def test_shard_and_merge_restore_reference_order():
    n, world = 11, 2
    lines = [f"read{i}\tpayload{i}".encode() for i in range(n)]
    outs = []
    for r in range(world):
        pos = multi.shard_positions(n, r, world)
        outs.append(b"".join(b"0\t%d\t" % p + lines[p] + b"\n" for p in pos))
    assert sorted(sum((multi.shard_positions(n, r, world) for r in range(world)), [])) == list(range(n))
    assert multi.merge_tagged(outs) == b"".join(ln + b"\n" for ln in lines)
