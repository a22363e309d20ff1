"""Multi-GPU: one process per GPU; the exchanges live in the C ABI (``csrc/comm.cu``, NCCL over NVLink / NVSwitch).

The path shards naturally: images are independent units and label rows combine by integer
sums.  No data-path collective is needed for hashing or resizing; the only exchanges are tiny:

  * dedupe across ranks — ``b2_dedupe_global``: all-gather of 32-byte digests + global listing positions + validity
    flags (shards may have different sizes: they are padded with invalid entries), then the SAME deterministic
    device resolution on every rank (first occurrence = smallest listing position, i.e. the reference's sequential
    "first seen wins", webdav_sync.py:324-354);
  * label aggregation — ``b2_allreduce_i64``: sum of the k class totals + 7 integer partials (+ the agreement
    histogram); kappa is then computed from integers on every rank, bit-identical for any GPU count.

This module is a thin binding: :class:`Comm` wraps ``b2_comm`` (the 128-byte NCCL unique id travels through
``torch.distributed``'s store when a process group exists, else through a file — the library itself needs neither),
plus the sharding rules, which are pure host arithmetic computed identically on every rank.  ``gather_shards`` is
the host-side statement of the padding rule on ``torch.distributed`` tensors (gloo on CPU in the tests).
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import check, lib


# --------------------------------------------------------------------------------------- sharding (host arithmetic)
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
    each in listing order (the global position that ``global_dedupe`` orders first / last occurrences by).  Shards
    generally have DIFFERENT sizes; ``global_dedupe`` pads them."""
    import heapq

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
    """Mode M1: shard label rows by image range.  Returns (img_lo, img_hi, row_lo, row_hi) for rows sorted by image
    index (NumPy array or torch tensor, host or device)."""
    img_lo, img_hi = shard_range(n_images, rank, world_size)
    if isinstance(image_idx_sorted, np.ndarray):
        lo, hi = np.searchsorted(image_idx_sorted, [img_lo, img_hi])
        return img_lo, img_hi, int(lo), int(hi)
    import torch
    t = torch.as_tensor(image_idx_sorted)
    bounds = torch.searchsorted(t, torch.tensor([img_lo, img_hi], dtype=t.dtype, device=t.device))
    return img_lo, img_hi, int(bounds[0]), int(bounds[1])


# --------------------------------------------------------------------------------------- the communicator
class Comm:
    """``b2_comm``: this rank's NCCL communicator inside libb2ingest.  Collectives are enqueued on the stream given
    (a raw ``cudaStream_t`` integer; default: PyTorch's current stream when PyTorch is loaded, else the legacy
    default stream) and never synchronise."""

    def __init__(self, rank: int, world: int, device: int, unique_id: bytes):
        assert len(unique_id) == _lib.B2_COMM_ID_BYTES
        self.rank, self.world, self.device = rank, world, device
        h = C.c_void_p()
        idbuf = (C.c_uint8 * _lib.B2_COMM_ID_BYTES).from_buffer_copy(unique_id)
        check(lib.b2_comm_init(device, rank, world, C.cast(idbuf, C.c_void_p), C.byref(h)))
        self._h = h
        self._ws = None                                       # device workspace of global_dedupe (a torch tensor)

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * _lib.B2_COMM_ID_BYTES)()
        check(lib.b2_comm_unique_id(C.cast(buf, C.c_void_p)))
        return bytes(buf)

    @classmethod
    def from_file(cls, path: str, rank: int, world: int, device: int, timeout_s: float = 120.0) -> "Comm":
        """Rendezvous without any framework: rank 0 writes the unique id to ``path`` (atomically), the others poll."""
        if rank == 0:
            uid = cls.unique_id()
            tmp = f"{path}.tmp{os.getpid()}"
            with open(tmp, "wb") as f:
                f.write(uid)
            os.replace(tmp, path)
        else:
            t0 = time.time()
            while not (os.path.exists(path) and os.path.getsize(path) == _lib.B2_COMM_ID_BYTES):
                if time.time() - t0 > timeout_s:
                    raise TimeoutError(f"no NCCL unique id at {path} after {timeout_s} s")
                time.sleep(0.01)
            with open(path, "rb") as f:
                uid = f.read()
        return cls(rank, world, device, uid)

    @classmethod
    def from_torch(cls, device: Optional[int] = None) -> "Comm":
        """Bootstrap over an initialised ``torch.distributed`` process group (any backend): rank 0's unique id is
        broadcast as a Python object.  The process group is used for nothing else."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        dev = torch.cuda.current_device() if device is None else int(device)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(rank, world, dev, box[0])

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.b2_comm_destroy(self._h)
            self._h = None

    def nccl_version(self) -> int:
        v = C.c_int()
        check(lib.b2_comm_info(self._h, None, None, C.byref(v)))
        return v.value

    # ---- collectives on device tensors (PyTorch tensors as buffers) ----
    @staticmethod
    def _stream(stream) -> int:
        if stream is not None:
            return int(stream)
        import sys
        t = sys.modules.get("torch")
        return int(t.cuda.current_stream().cuda_stream) if t is not None else 0

    def allreduce_i64(self, values, stream=None):
        """In-place sum over ranks of an int64 device tensor (partials, histogram...)."""
        assert values.is_cuda and values.is_contiguous() and values.element_size() == 8
        check(lib.b2_allreduce_i64(self._h, values.data_ptr(), values.numel(), self._stream(stream)))
        return values

    # ---- the label all-reduce over NVLink peer memory, fused into the tally kernel ----
    def enable_peer_reduce(self, max_values: int) -> None:
        """``b2_comm_enable_peer_reduce``: collective, blocking, once — mailboxes in every rank's HBM mapped by its peers
        through CUDA IPC."""
        check(lib.b2_comm_enable_peer_reduce(self._h, int(max_values)))
        self.peer_values = int(max_values)

    def peer_timed_out(self) -> bool:
        v = C.c_int()
        check(lib.b2_comm_peer_status(self._h, C.byref(v)))
        return bool(v.value)

    def peer_allreduce_i64(self, values, stream=None):
        """In-place sum over ranks through the mailboxes (one small kernel, no NCCL call)."""
        assert values.is_cuda and values.is_contiguous() and values.element_size() == 8
        check(lib.b2_peer_allreduce_i64(self._h, values.data_ptr(), values.numel(), self._stream(stream)))
        return values

    def label_tally_reduce(self, image_idx, class_idx, active, n_images: int, k: int, image_base: int, counts, vec,
                           sorted_by_image: bool = True, stream=None):
        """``b2_label_tally_reduce``: the tally of this rank's rows whose kernel also all-reduces ``vec`` = int64[k + 7 +
        B2_AGREE_BINS] (partials, then the agreement histogram) over the ranks; ``counts`` stays this rank's slab."""
        assert vec.is_cuda and vec.is_contiguous() and vec.numel() == k + _lib.B2_PARTIALS_EXTRA + _lib.B2_AGREE_BINS
        check(lib.b2_label_tally_reduce(self._h, image_idx.data_ptr(), class_idx.data_ptr(), active.data_ptr(),
                                        image_idx.numel(), image_base, n_images, k,
                                        _lib.B2_TALLY_SORTED if sorted_by_image else 0, counts.data_ptr(), vec.data_ptr(),
                                        self._stream(stream)))
        return counts, vec

    def allgather_digests(self, digests, stream=None):
        """uint8[n,32] on every rank (same n) -> uint8[world*n,32] in rank order."""
        import torch
        n = digests.numel() // 32
        out = torch.empty((self.world * n, 32), dtype=torch.uint8, device=digests.device)
        check(lib.b2_allgather_digests(self._h, digests.data_ptr(), n, out.data_ptr(), self._stream(stream)))
        return out

    def global_dedupe(self, digests, seq, n_max: int, valid=None, existing_sorted=None, stream=None):
        """``b2_dedupe_global`` for this rank's shard: ``digests`` uint8[n_local,32], ``seq`` int32/uint32[n_local]
        global listing positions, ``n_max`` = the largest shard of any rank (every rank passes the same value — the
        sharding rules above are deterministic, so nothing needs exchanging to know it).  Returns
        ``(is_new u8[n_local], first_seq i64[n_local], last_seq i64[n_local], counts i32[3])`` on the device;
        ``counts`` = (processed, created, updated) of the WHOLE listing."""
        import torch
        dev = digests.device
        n_local = digests.numel() // 32
        assert seq.numel() == n_local and seq.element_size() == 4 and n_local <= n_max
        need = int(lib.b2_dedupe_global_workspace_bytes(self.world, n_max))
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)     # cudaMalloc: 256-byte aligned
        is_new = torch.empty(n_local, dtype=torch.uint8, device=dev)
        first = torch.empty(n_local, dtype=torch.int64, device=dev)
        last = torch.empty(n_local, dtype=torch.int64, device=dev)
        counts = torch.empty(3, dtype=torch.int32, device=dev)
        m = 0 if existing_sorted is None else existing_sorted.numel() // 32
        check(lib.b2_dedupe_global(
            self._h, digests.data_ptr() if n_local else None, valid.data_ptr() if valid is not None else None,
            seq.data_ptr() if n_local else None, n_local, n_max,
            existing_sorted.data_ptr() if m else None, m, is_new.data_ptr(), first.data_ptr(), last.data_ptr(),
            counts.data_ptr(), self._ws.data_ptr(), self._ws.numel(), self._stream(stream)))
        return is_new, first, last, counts


_comm: Optional[Comm] = None


def comm() -> Optional[Comm]:
    """The process-wide communicator: created on first use from the ``torch.distributed`` process group (if one is
    initialised with more than one rank), ``None`` for a single process."""
    global _comm
    if _comm is None:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            _comm = Comm.from_torch()
    return _comm


def set_comm(c: Optional[Comm]) -> None:
    global _comm
    _comm = c


def world() -> Tuple[int, int]:
    if _comm is not None:
        return _comm.rank, _comm.world
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allreduce_partials(partials):
    """Sum the int64 partial vector over ranks: ``b2_allreduce_i64`` for device tensors; host tensors (the gloo
    tests) go through ``torch.distributed``."""
    if world()[1] > 1:
        if partials.is_cuda:
            comm().allreduce_i64(partials)
        else:
            import torch.distributed as dist
            dist.all_reduce(partials, op=dist.ReduceOp.SUM)
    return partials


def gather_shards(digests, seq):
    """Host-side statement of ``b2_dedupe_global``'s exchange on ``torch.distributed`` tensors (gloo on CPU): shards of
    different sizes are padded to the largest with invalid entries and all-gathered.  Returns ``(all_digests
    [world*n_max,32], all_seq [world*n_max] (int64), valid u8[world*n_max], (lo, hi))`` with ``[lo, hi)`` this
    rank's entries in the concatenation."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    n_local = digests.shape[0]
    if ws == 1:
        return digests, seq.to(torch.int64), torch.ones(n_local, dtype=torch.uint8, device=digests.device), (0, n_local)
    counts = torch.zeros(ws, dtype=torch.int64, device=digests.device)
    counts[rank] = n_local
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    n_max = int(counts.max())
    pad_d = torch.zeros((n_max, 32), dtype=torch.uint8, device=digests.device)
    pad_s = torch.full((n_max,), -1, dtype=torch.int64, device=digests.device)
    pad_v = torch.zeros(n_max, dtype=torch.uint8, device=digests.device)
    pad_d[:n_local], pad_s[:n_local], pad_v[:n_local] = digests, seq.to(torch.int64), 1
    out = []
    for t in (pad_d, pad_s, pad_v):
        parts = [torch.empty_like(t) for _ in range(ws)]
        dist.all_gather(parts, t.contiguous())
        out.append(torch.cat(parts, 0))
    lo = rank * n_max
    return out[0], out[1], out[2], (lo, lo + n_local)


def global_dedupe(digests, seq, n_max: Optional[int] = None, valid=None, existing_sorted=None):
    """Cross-rank dedupe decision for this rank's shard (device tensors).  Single process: ``b2_dedupe`` keyed on
    ``seq``.  ``n_max`` (largest shard over ranks) defaults to an exchange of the shard sizes.  Returns
    ``(is_new, first_seq, last_seq, counts)`` as :meth:`Comm.global_dedupe`."""
    import torch
    from . import engine
    c = comm()
    n_local = digests.numel() // 32
    if c is None:
        is_new, first, last, counts = engine.dedupe_device(digests, valid=valid, seq=seq, existing_sorted=existing_sorted)
        s64 = seq.to(torch.int64)
        pick = lambda ix: torch.where(ix >= 0, s64[ix.clamp(min=0).long()], torch.full_like(s64, -1))  # noqa: E731
        return is_new, pick(first), pick(last), counts
    if n_max is None:
        sizes = torch.zeros(c.world, dtype=torch.int64, device=digests.device)
        sizes[c.rank] = n_local
        c.allreduce_i64(sizes)
        n_max = max(1, int(sizes.max().item()))
    return c.global_dedupe(digests, seq, n_max, valid=valid, existing_sorted=existing_sorted)
