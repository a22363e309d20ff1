"""Multi-GPU: one process per GPU (torchrun), torch.distributed over NCCL as plumbing.

The path shards naturally: images are independent units and label rows
combine by integer sums.  No data-path collective is needed for hashing or resizing; the only
exchanges are tiny:

  * dedupe across ranks — all-gather of 32-byte digests, then the SAME deterministic device
    resolution on every rank (first occurrence = smallest global image index, matching the
    reference's sequential "first seen wins", webdav_sync.py:324-354);
  * label aggregation — all-reduce (sum, int64) of the k class totals + 7 integer partials;
    kappa is then computed from integers on every rank, bit-identical for any GPU count.

The same functions run on CPU tensors with the gloo backend (tests use world_size 2); the
collectives carry only integers, so results do not depend on the backend.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of ``n`` units for ``rank``."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_by_bytes(lengths, world_size: int):
    """Size-aware sharding of a mixed-size listing (BASELINE config 3; SURVEY.md section 8(e): "size-aware
    round-robin so bytes/GPU balance").  Greedy longest-first: images are taken by decreasing byte length (ties by
    listing position) and each goes to the rank that holds the fewest bytes so far (ties to the lowest rank), so the
    byte totals of any two ranks differ by at most the largest image.  Deterministic — every rank computes the same
    assignment from the same listing, nothing is exchanged.  Returns one int64 array of listing positions per rank,
    each in listing order (the global image index that ``global_dedupe`` orders first / last occurrences by)."""
    import heapq

    import numpy as np

    ln = np.asarray(lengths, dtype=np.int64)
    order = np.lexsort((np.arange(ln.shape[0]), -ln))
    heap = [(0, r) for r in range(world_size)]
    owner = np.empty(ln.shape[0], dtype=np.int32)
    for i in order:
        b, r = heapq.heappop(heap)
        owner[i] = r
        heapq.heappush(heap, (b + int(ln[i]), r))
    return [np.nonzero(owner == r)[0].astype(np.int64) for r in range(world_size)]


def shard_rows_by_image(image_idx_sorted, n_images: int, rank: int, world_size: int) -> Tuple[int, int, int, int]:
    """Mode M1: shard label rows by image range.  Returns (img_lo, img_hi, row_lo, row_hi) for
    rows sorted by image index (NumPy or torch 1-D)."""
    img_lo, img_hi = shard_range(n_images, rank, world_size)
    t = torch.as_tensor(image_idx_sorted)
    bounds = torch.searchsorted(t, torch.tensor([img_lo, img_hi], dtype=t.dtype, device=t.device))
    return img_lo, img_hi, int(bounds[0]), int(bounds[1])


def allreduce_partials(partials: torch.Tensor) -> torch.Tensor:
    """Sum the int64 partial vector over ranks (NCCL all-reduce over NVLink; gloo on CPU)."""
    if world()[1] > 1:
        dist.all_reduce(partials, op=dist.ReduceOp.SUM)
    return partials


def allgather_digests(digests: torch.Tensor, global_index: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Gather every rank's uint8[n_r,32] digests and int32[n_r] global image indices (equal n_r
    on all ranks: pad with an invalid index if needed).  Returns the concatenation in rank order."""
    _, ws = world()
    if ws == 1:
        return digests, global_index
    dg = [torch.empty_like(digests) for _ in range(ws)]
    ix = [torch.empty_like(global_index) for _ in range(ws)]
    dist.all_gather(dg, digests.contiguous())
    dist.all_gather(ix, global_index.contiguous())
    return torch.cat(dg, 0), torch.cat(ix, 0)


def global_dedupe(digests: torch.Tensor, global_index: torch.Tensor, existing_sorted: Optional[torch.Tensor] = None):
    """Cross-rank dedupe decision for this rank's images.  Every rank gathers all digests and
    runs the same device resolution keyed on the global index; returns this rank's slice of
    ``is_new`` plus the global (processed, created, updated) counts."""
    from . import engine

    rank, ws = world()
    n_local = digests.shape[0]
    all_d, all_i = allgather_digests(digests, global_index)
    is_new, first, last, counts = engine.dedupe_device(all_d.contiguous(), seq=all_i.contiguous(),
                                                       existing_sorted=existing_sorted)
    lo = rank * n_local
    return is_new[lo:lo + n_local], counts
