"""Data-parallel sharding of independent streams over the GPUs of one box.

Streams never exchange data (each chunk of ``stego_encode`` builds its own context and interval,
src/neuralstego/api.py:736-747), so ranks just take contiguous slices and the only collective is
the final gather of the cover tokens / recovered bitstreams.  Works with any ``torch.distributed``
backend: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_streams: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of rank ``rank``; the first ``n % world`` ranks take one extra stream."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(int(n_streams), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_list(items: Sequence, rank: int, world: int) -> list:
    lo, hi = shard_bounds(len(items), rank, world)
    return list(items[lo:hi])


def gather_ragged(rows: torch.Tensor, lengths: torch.Tensor, *, dst: int = 0,
                  group: Optional[dist.ProcessGroup] = None) -> Optional[List[List[int]]]:
    """Gather per-stream integer rows ``[B_local, W]`` with their valid lengths on rank ``dst``.

    Ranks may hold different numbers of streams; returns the concatenated python lists on ``dst``
    (stream order = rank order, then local order) and ``None`` elsewhere.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = rows.device
    shape = torch.tensor([rows.shape[0], rows.shape[1]], dtype=torch.int64, device=dev)
    shapes = [torch.zeros_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape, group=group)
    max_b = int(max(int(s[0]) for s in shapes))
    max_w = int(max(int(s[1]) for s in shapes))
    pad = torch.full((max_b, max_w), -1, dtype=rows.dtype, device=dev)
    pad[: rows.shape[0], : rows.shape[1]] = rows
    plen = torch.zeros(max_b, dtype=lengths.dtype, device=dev)
    plen[: lengths.shape[0]] = lengths
    out_rows = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    out_len = [torch.empty_like(plen) for _ in range(world)] if rank == dst else None
    dist.gather(pad, out_rows, dst=dst, group=group)
    dist.gather(plen, out_len, dst=dst, group=group)
    if rank != dst:
        return None
    result: List[List[int]] = []
    for r in range(world):
        nb = int(shapes[r][0])
        rr, ll = out_rows[r].cpu(), out_len[r].cpu()
        for b in range(nb):
            result.append(rr[b, : int(ll[b])].tolist())
    return result
