"""Multi-GPU partitioning: one process per GPU (torch.distributed), cells sharded by contiguous
index range, a single gather of per-cell results at the end.  The reference has no
distributed code; every rank here ends up with the whole normalised image (stage 1 is split by channel and the
planes are broadcast, pipeline.HotPath._normalize_sharded) and the mask, runs stage 2 redundantly (1 ms),
then owns cells [lo, hi) for stages 3-5.  Collectives: all_gather of (label, confidence) and
all_reduce of the 18 per-type counts - NCCL on GPUs, gloo in the CPU tests.
A batch of images (the batch-processing CSV) is sharded by image instead: image i belongs to rank i % world
(`image_owner`), no collective touches the data path, and the owner broadcasts the compact per-cell results at the end
(`share_image_results`).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, nranks: int):
    """Contiguous, balanced split of n items: the first n % nranks ranks get one extra."""
    base, rem = divmod(n, nranks)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(local: torch.Tensor, n_total: int, lo: int, hi: int) -> torch.Tensor:
    """Concatenate every rank's rows [lo, hi) into the full (n_total, ...) tensor on every rank."""
    rank, nranks = world()
    if nranks == 1:
        return local
    assert local.shape[0] == hi - lo
    longest = (n_total + nranks - 1) // nranks
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: hi - lo] = local
    parts = [torch.empty_like(pad) for _ in range(nranks)]
    dist.all_gather(parts, pad)
    out = []
    for r, p in enumerate(parts):
        a, b = shard_range(n_total, r, nranks)
        out.append(p[: b - a])
    return torch.cat(out)


def all_reduce_sum(t: torch.Tensor) -> torch.Tensor:
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def image_owner(i: int, nranks: int) -> int:
    """Round-robin image partition of a batch (batch CSV -> images round-robin to GPUs)."""
    return i % nranks


def share_image_results(results: list, device) -> list:
    """results[i] = (label uint8 (n,), conf float32 (n,), counts int64 (18,)) on the owner rank of image i, None elsewhere.
    Every rank returns the full list: per image one broadcast of the cell count and one of each array from its owner."""
    rank, nranks = world()
    if nranks == 1:
        return results
    out = []
    for i, res in enumerate(results):
        src = image_owner(i, nranks)
        n = torch.tensor([res[0].shape[0] if rank == src else 0], dtype=torch.int64, device=device)
        dist.broadcast(n, src)
        n = int(n.item())
        if rank == src:
            label, conf, counts = (t.to(device).contiguous() for t in res)
        else:
            label = torch.empty(n, dtype=torch.uint8, device=device)
            conf = torch.empty(n, dtype=torch.float32, device=device)
            counts = torch.empty(18, dtype=torch.int64, device=device)
        for t in (label, conf, counts):
            dist.broadcast(t, src)
        out.append((label, conf, counts))
    return out


def is_writer() -> bool:
    """Filesystem side effects (log, result CSVs / PNGs, tmp wipe) belong to rank 0 when several ranks share a main_dir."""
    return world()[0] == 0


def barrier() -> None:
    if world()[1] > 1:
        dist.barrier()


def broadcast_object(obj, src: int = 0):
    """Host object computed on `src` (e.g. the unseeded KMeans region labels) -> the same object on every rank."""
    if world()[1] == 1:
        return obj
    box = [obj if world()[0] == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]
