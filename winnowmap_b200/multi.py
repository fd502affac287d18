"""Multi-GPU helpers: one process per GPU, reads shard with no data-path collective; the only collective is the
one-time broadcast of the flattened index (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import numpy as np


def broadcast_blob(blob, src, rank, device=None):
    """Broadcast a numpy uint8 array from `src` to all ranks with torch.distributed; returns the array on every rank.
    With a CUDA `device` the payload travels GPU-to-GPU (NCCL), otherwise through host tensors (gloo)."""
    import torch
    import torch.distributed as dist
    n = torch.zeros(1, dtype=torch.int64, device=device or "cpu")
    if rank == src:
        n[0] = blob.nbytes
    dist.broadcast(n, src)
    size = int(n.item())
    if rank == src:
        t = torch.from_numpy(blob).to(device or "cpu")
    else:
        t = torch.empty(size, dtype=torch.uint8, device=device or "cpu")
    dist.broadcast(t, src)
    return blob if rank == src else t.cpu().numpy()


def shard_positions(n, rank, world):
    """Positions of the length-sorted mini-batch mapped by `rank` (round-robin deal, see wm_map_file)."""
    return list(range(rank, n, world))


def merge_tagged(outputs):
    """Merge per-rank outputs whose lines start with "<batch>\\t<position>\\t" back into the reference's order."""
    recs = []
    for text in outputs:
        for ln in text.splitlines():
            if not ln:
                continue
            b, p, rest = ln.split(b"\t", 2)
            recs.append((int(b), int(p), len(recs), rest))
    recs.sort(key=lambda r: (r[0], r[1], r[2]))
    return b"".join(r[3] + b"\n" for r in recs)
