"""Read sharding across the GPUs of one box: host-side arithmetic only.

Reads (or pairs) are independent units and `myrand` is a stateless function of the global read index
(utilities.cpp:38-48), so a rank can map any contiguous range of the input on its own replica of the index and the
records are the same as in a single-process run, provided the range keeps its GLOBAL read indices
(`ReadBatch.first_index`). There is no collective on the data path; `gather_records` (rank 0 collects the fixed-size
records in input order, as the `basal` CLI's writer does) is the only exchange and it carries results, not work.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the reads rank `rank` of `world` maps: contiguous, in input order, sizes differ by at most 1."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_batch(batch, rank: int, world: int):
    """The sub-batch of a `capi.ReadBatch` that `rank` maps; global read indices are preserved."""
    from . import capi
    b, e = shard_range(batch.n, rank, world)
    off = np.asarray(batch.offsets, dtype=np.uint64)
    bases = np.asarray(batch.bases)[int(off[b]):int(off[e])]
    sub_off = (off[b:e + 1] - off[b]).astype(np.uint64)
    index = None if batch.index is None else np.ascontiguousarray(batch.index[b:e])
    raw = None if batch.raw_len is None else np.ascontiguousarray(batch.raw_len[b:e])
    return capi.ReadBatch(np.ascontiguousarray(bases), sub_off, readset=batch.readset,
                          first_index=batch.first_index + b, index=index, raw_len=raw)


def merge_records(parts: Sequence[np.ndarray]) -> np.ndarray:
    """Concatenate per-rank record arrays (given in rank order) back into input order."""
    return np.concatenate(list(parts)) if parts else np.zeros(0)


def gather_records(local: np.ndarray, dist=None, dst: int = 0) -> Optional[np.ndarray]:
    """Rank `dst` receives every rank's records in input order; other ranks get None.

    Works with any initialised torch.distributed backend (gloo on CPU in the tests, nccl on the GPU box): records travel
    as raw bytes, lengths first because shards differ by one read.
    """
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    raw = torch.from_numpy(np.ascontiguousarray(local).view(np.uint8).copy()).to(dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([raw.numel()], dtype=torch.int64, device=dev))
    sizes = [int(s.item()) for s in sizes]
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=torch.uint8, device=dev); buf[:raw.numel()] = raw
    outs: List[torch.Tensor] = [torch.zeros(pad, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == dst else []
    dist.gather(buf, outs if rank == dst else None, dst=dst)
    if rank != dst:
        return None
    parts = [o[:s].cpu().numpy().view(local.dtype) for o, s in zip(outs, sizes)]
    return merge_records(parts)
