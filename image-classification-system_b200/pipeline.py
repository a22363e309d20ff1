"""Streaming ingest for fixed-shape batches held in HOST memory: the end-to-end form of the path.

    pinned host images --H2D--> [sha256 | resize+normalise] --D2H--> digests, thumbnails, previews
                                         \\--> dedupe over the whole batch --D2H--> flags + stats

The pipeline itself is native: ``b2_ingest_stream_*`` in ``csrc/host.cu`` (C ABI, host pointers only) owns
the device staging buffer, the copy stream, the 32 hash streams and the two resize streams.  SHA-256 is
serial per message: one lane hashes one image at ~48 MB/s, so a 1080p image takes ~130 ms however few
images are in flight; the batch is therefore copied in small chunks and each chunk's hash kernel runs on
one of several streams, so many hash kernels (8 warps each) overlap each other and the remaining copies;
PCIe, not the hash latency, is the limit for batches of a few thousand images.  This module only wraps the
C calls: it allocates the page-locked result buffers (PyTorch is used for pinned memory, nothing else) and
keeps them alive while a batch is in flight.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import engine
from ._lib import check, lib


@dataclass
class PipelineResult:
    digests: torch.Tensor                 # pinned uint8 [n, 32]
    is_new: torch.Tensor                  # pinned uint8 [n]
    stats: Dict[str, int]                 # {'processed','created','updated'}
    thumbs: torch.Tensor                  # pinned uint8 [n, out_h, out_w, 3]
    previews: Optional[torch.Tensor]      # pinned float32 [n, 3, out_h, out_w]
    h2d_bytes: int
    d2h_bytes: int
    first_index: Optional[torch.Tensor] = None    # pinned int32 [n]
    last_index: Optional[torch.Tensor] = None


class IngestPipeline:
    def __init__(self, in_h: int, in_w: int, max_images: int, chunk_images: int = 256, out_h: int = 256,
                 out_w: int = 256, want_preview: bool = True, device: Optional[int] = None):
        self.device = engine.init(device)
        self.in_h, self.in_w, self.out_h, self.out_w = in_h, in_w, out_h, out_w
        self.L = in_h * in_w * 3
        self.max_images = max_images
        self.chunk = max(1, min(chunk_images, max_images))
        self.want_preview = want_preview
        h = C.c_void_p()
        check(lib.b2_ingest_stream_create(self.device, in_h, in_w, out_h, out_w, max_images, self.chunk,
                                          1 if want_preview else 0, C.byref(h)))
        self._h = h
        pin = dict(pin_memory=True)
        self.h_digests = torch.empty((max_images, 32), dtype=torch.uint8, **pin)
        self.h_is_new = torch.empty(max_images, dtype=torch.uint8, **pin)
        self.h_first = torch.empty(max_images, dtype=torch.int32, **pin)
        self.h_last = torch.empty(max_images, dtype=torch.int32, **pin)
        self.h_counts = torch.empty(4, dtype=torch.int32, **pin)
        self.h_thumbs = torch.empty((max_images, out_h, out_w, 3), dtype=torch.uint8, **pin)
        self.h_prev = torch.empty((max_images, 3, out_h, out_w), dtype=torch.float32, **pin) if want_preview else None
        self.kernel_launches = 0
        self._pending = None

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.b2_ingest_stream_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def run(self, host_images: torch.Tensor, existing_sorted: Optional[torch.Tensor] = None) -> PipelineResult:
        """Blocking form: submit + result."""
        self.submit(host_images, existing_sorted)
        return self.result()

    def submit(self, host_images: torch.Tensor, existing_sorted: Optional[torch.Tensor] = None) -> None:
        """Enqueue the whole batch (copies, kernels, read-backs) without waiting for the GPU.
        host_images: uint8 [n, in_h*in_w*3] in host memory (raw RGB HWC = the synthetic "file bytes"),
        page-locked for full-speed asynchronous copies.  existing_sorted: uint8 [m, 32] digests already in the
        table, sorted in memcmp order (host).  A service keeps two pipelines and submits batch i+1 before
        asking for result i, so the hash tail of one batch hides under the copies of the next."""
        n = host_images.shape[0]
        assert n <= self.max_images and host_images.dtype == torch.uint8 and not host_images.is_cuda
        assert host_images.is_contiguous() and host_images.numel() == n * self.L
        m = 0
        ex_ptr = None
        if existing_sorted is not None and existing_sorted.numel():
            existing_sorted = existing_sorted.cpu().contiguous()
            m = existing_sorted.numel() // 32
            ex_ptr = existing_sorted.data_ptr()
        check(lib.b2_ingest_stream_submit(
            self._h, host_images.data_ptr(), n, ex_ptr, m, self.h_digests.data_ptr(), self.h_is_new.data_ptr(),
            self.h_first.data_ptr(), self.h_last.data_ptr(), self.h_counts.data_ptr(), self.h_thumbs.data_ptr(),
            self.h_prev.data_ptr() if self.h_prev is not None else None))
        self._pending = (n, host_images, existing_sorted)      # keep the inputs alive until the copies are done

    def result(self) -> PipelineResult:
        """Wait for the submitted batch and hand back the page-locked host results."""
        n = self._pending[0]
        h2d, d2h, launches = C.c_uint64(), C.c_uint64(), C.c_uint32()
        check(lib.b2_ingest_stream_wait(self._h, C.byref(h2d), C.byref(d2h), C.byref(launches)))
        self._pending = None
        self.kernel_launches = int(launches.value)
        c = self.h_counts.tolist()
        return PipelineResult(self.h_digests[:n], self.h_is_new[:n],
                              {"processed": c[0], "created": c[1], "updated": c[2]},
                              self.h_thumbs[:n], self.h_prev[:n] if self.h_prev is not None else None,
                              int(h2d.value), int(d2h.value), self.h_first[:n], self.h_last[:n])


@dataclass
class MixedResult:
    """Digests and dedupe decisions are in listing order.  Thumbnails and previews stay where the per-shape streams
    put them (page-locked, one block per shape class: scattering ~1 MB per image into listing order costs more host
    time than the GPU needs for the images); ``thumb(i)`` / ``preview(i)`` address them by listing position."""
    digests: "np.ndarray"                 # uint8 [n, 32], listing order
    is_new: "np.ndarray"                  # uint8 [n]
    first_index: "np.ndarray"             # int32 [n]: first occurrence of the same content in the listing
    last_index: "np.ndarray"
    stats: Dict[str, int]
    thumbs_by_shape: Dict                 # (h, w) -> uint8 [n_s, out_h, out_w, 3]
    previews_by_shape: Optional[Dict]     # (h, w) -> float32 [n_s, 3, out_h, out_w]
    where: "np.ndarray"                   # int32 [n, 2]: (index of the shape class in `shapes`, row inside its block)
    shapes: list
    h2d_bytes: int
    d2h_bytes: int

    def thumb(self, i: int):
        k, j = self.where[i]
        return self.thumbs_by_shape[self.shapes[k]][j]

    def preview(self, i: int):
        k, j = self.where[i]
        return None if self.previews_by_shape is None else self.previews_by_shape[self.shapes[k]][j]


class MixedShapeIngest:
    """Streaming ingest of a listing whose images have different shapes (BASELINE config 3).

    The native stream (``b2_ingest_stream_*``) is fixed-shape, so a listing is split by shape: one
    :class:`IngestPipeline` per shape class, all submitted before any is waited for (their copies and kernels
    share the GPU), and ONE dedupe over the digests of the whole listing in listing order afterwards
    (``b2_dedupe_host``) — two files of equal byte length can be byte-identical whatever their shapes, and the
    first-seen / last-seen rule of ``webdav_sync.py:324-398`` is about listing order, not shape order.

    ``run(groups)``: ``groups[(h, w)] = (images, positions)`` with ``images`` a uint8 host tensor
    ``[n_s, h*w*3]`` (page-locked for full-speed copies) and ``positions`` the listing index of each of them
    (int array, all groups together a permutation of ``0..n-1``).  As for the native stream, ``h*w*3`` must be a
    multiple of 16 (images are packed back to back and copied in 16-byte units); other shapes go through the blocking
    ``hostapi.sha256_host`` + ``hostapi.thumbnails``.
    """

    def __init__(self, capacity: Dict, chunk_bytes: int = 1 << 30, out_h: int = 256, out_w: int = 256,
                 want_preview: bool = True, device: Optional[int] = None):
        """``capacity[(h, w)]`` = most images of that shape in one listing (device staging is sized by it)."""
        self.out_h, self.out_w, self.want_preview = out_h, out_w, want_preview
        self.device = engine.init(device)
        self.pipes: Dict = {}
        for (h, w), cap in capacity.items():
            chunk = max(1, min(cap, chunk_bytes // (h * w * 3)))
            self.pipes[(h, w)] = IngestPipeline(h, w, cap, chunk_images=chunk, out_h=out_h, out_w=out_w,
                                                want_preview=want_preview, device=self.device)

    def close(self) -> None:
        for p in self.pipes.values():
            p.close()
        self.pipes = {}

    def run(self, groups: Dict, existing_sorted=None) -> MixedResult:
        """Blocking form: submit + result."""
        self.submit(groups)
        return self.result(existing_sorted)

    def submit(self, groups: Dict) -> None:
        """Enqueue every shape class of the listing without waiting for the GPU.  A caller with a long listing keeps
        two ``MixedShapeIngest`` objects and submits listing i+1 before asking for result i: the hash tail of the
        biggest images (one lane needs ~1.15 s for 50 MB) then hides under the copies of the next listing."""
        # biggest images first: their hash latency is the tail everything else hides under
        self._order = sorted(groups, key=lambda s: -s[0] * s[1])
        self._groups = groups
        for shape in self._order:
            self.pipes[shape].submit(groups[shape][0])

    def result(self, existing_sorted=None) -> MixedResult:
        import numpy as np

        from . import hostapi

        groups, order = self._groups, self._order
        n = sum(int(imgs.shape[0]) for imgs, _ in groups.values())
        digests = np.zeros((n, 32), dtype=np.uint8)
        where = np.zeros((n, 2), dtype=np.int32)
        thumbs, previews = {}, ({} if self.want_preview else None)
        seen = np.zeros(n, dtype=bool)
        h2d = d2h = 0
        for k, shape in enumerate(order):
            res = self.pipes[shape].result()
            pos = np.asarray(groups[shape][1], dtype=np.int64)
            assert pos.shape[0] == res.digests.shape[0] and not seen[pos].any(), "positions must be a permutation"
            seen[pos] = True
            digests[pos] = res.digests.numpy()
            where[pos, 0] = k
            where[pos, 1] = np.arange(pos.shape[0], dtype=np.int32)
            thumbs[shape] = res.thumbs.numpy()              # views of the pipeline's page-locked buffers: valid until
            if previews is not None:                        # the next submit() on this object
                previews[shape] = res.previews.numpy()
            h2d += res.h2d_bytes
            d2h += res.d2h_bytes
        assert seen.all(), "positions must cover the listing"
        self._groups = self._order = None
        ex = None
        if existing_sorted is not None and len(existing_sorted):
            ex = np.ascontiguousarray(np.asarray(existing_sorted, dtype=np.uint8)).reshape(-1, 32)
        is_new, first, last, c = hostapi.dedupe_host(digests, None, ex, self.device)
        return MixedResult(digests, is_new, first, last, {"processed": c[0], "created": c[1], "updated": c[2]},
                           thumbs, previews, where, list(order), h2d, d2h)
