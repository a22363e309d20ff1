"""Streaming ingest for fixed-shape batches held in HOST memory: the end-to-end form of the path.

    pinned host images --H2D--> [sha256 | resize+normalise] --D2H--> digests, thumbnails, previews
                                         \\--> dedupe over the whole batch --D2H--> flags + stats

SHA-256 is serial per message: one lane hashes one image at ~48 MB/s, so a 1080p image takes
~130 ms however few images are in flight.  The pipeline therefore copies the batch in small chunks
on ONE copy stream and launches each chunk's hash kernel on one of several compute streams, so many
hash kernels (8 warps each) overlap each other and the remaining copies; PCIe, not the hash latency,
is the limit for batches of a few thousand images.  Resize (HBM bound, ~3 us per image) follows each
chunk on a second set of streams and its outputs are copied back while later chunks still arrive.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import engine


@dataclass
class PipelineResult:
    digests: torch.Tensor                 # pinned uint8 [n, 32]
    is_new: torch.Tensor                  # pinned uint8 [n]
    stats: Dict[str, int]                 # {'processed','created','updated'}
    thumbs: torch.Tensor                  # pinned uint8 [n, out_h, out_w, 3]
    previews: Optional[torch.Tensor]      # pinned float32 [n, 3, out_h, out_w]
    h2d_bytes: int
    d2h_bytes: int


class IngestPipeline:
    def __init__(self, in_h: int, in_w: int, max_images: int, chunk_images: int = 256, out_h: int = 256,
                 out_w: int = 256, want_preview: bool = True, device: Optional[int] = None, n_streams: int = 8):
        self.dev = torch.device("cuda", engine.init(device))
        self.in_h, self.in_w, self.out_h, self.out_w = in_h, in_w, out_h, out_w
        self.L = in_h * in_w * 3
        assert self.L % 16 == 0, "fixed-shape pipeline needs a 16-byte aligned image size"
        self.chunk = max(1, min(chunk_images, max_images))
        self.max_images = max_images
        self.plan = engine.get_plan(in_h, in_w, out_h, out_w, self.dev.index)
        dev = self.dev
        self.stage = torch.empty(max_images * self.L, dtype=torch.uint8, device=dev)      # whole batch resident
        self.offsets = torch.arange(max_images, dtype=torch.int64, device=dev) * self.L
        self.lengths = torch.full((max_images,), self.L, dtype=torch.int64, device=dev)
        self.d_digests = torch.empty((max_images, 32), dtype=torch.uint8, device=dev)
        self.d_thumbs = torch.empty((max_images, out_h, out_w, 3), dtype=torch.uint8, device=dev)
        self.d_prev = torch.empty((max_images, 3, out_h, out_w), dtype=torch.float32, device=dev) if want_preview else None
        pin = dict(pin_memory=True)
        self.h_digests = torch.empty((max_images, 32), dtype=torch.uint8, **pin)
        self.h_is_new = torch.empty(max_images, dtype=torch.uint8, **pin)
        self.h_counts = torch.empty(3, dtype=torch.int32, **pin)
        self.h_thumbs = torch.empty((max_images, out_h, out_w, 3), dtype=torch.uint8, **pin)
        self.h_prev = torch.empty((max_images, 3, out_h, out_w), dtype=torch.float32, **pin) if want_preview else None
        self.copy_stream = torch.cuda.Stream(dev)
        self.hash_streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
        self.resize_streams = [torch.cuda.Stream(dev) for _ in range(2)]
        self.final_stream = torch.cuda.Stream(dev)            # joins the others; private, so pipelines do not serialise
        self.kernel_launches = 0

    def run(self, host_images: torch.Tensor, existing_sorted: Optional[torch.Tensor] = None) -> PipelineResult:
        """Blocking form: submit + result."""
        self.submit(host_images, existing_sorted)
        return self.result()

    def submit(self, host_images: torch.Tensor, existing_sorted: Optional[torch.Tensor] = None) -> None:
        """Enqueue the whole batch (copies, kernels, read-backs) without waiting for the GPU.
        host_images: pinned uint8 [n, in_h*in_w*3] (raw RGB HWC = the synthetic "file bytes").
        A service keeps two pipelines and submits batch i+1 before asking for result i, so the hash
        tail of one batch hides under the copies of the next."""
        n = host_images.shape[0]
        assert n <= self.max_images and host_images.is_pinned() and host_images.dtype == torch.uint8
        main = torch.cuda.current_stream(self.dev)
        start = torch.cuda.Event()
        start.record(main)
        self.copy_stream.wait_event(start)
        for s in self.hash_streams + self.resize_streams:
            s.wait_event(start)
        h2d = d2h = 0
        self.kernel_launches = 0
        flat = host_images.view(n, self.L)
        for c, lo in enumerate(range(0, n, self.chunk)):
            hi = min(lo + self.chunk, n)
            m = hi - lo
            dev_chunk = self.stage[lo * self.L: hi * self.L]
            with torch.cuda.stream(self.copy_stream):
                dev_chunk.copy_(flat[lo:hi].reshape(-1), non_blocking=True)
                h2d += m * self.L
                copied = torch.cuda.Event()
                copied.record(self.copy_stream)
            hs = self.hash_streams[c % len(self.hash_streams)]
            with torch.cuda.stream(hs):
                hs.wait_event(copied)
                engine.sha256_device(self.stage, self.offsets[lo:hi], self.lengths[lo:hi], None, self.d_digests[lo:hi])
                self.kernel_launches += 1
            rs = self.resize_streams[c % len(self.resize_streams)]
            with torch.cuda.stream(rs):
                rs.wait_event(copied)
                thumbs = self.d_thumbs[lo:hi]
                prev = self.d_prev[lo:hi] if self.d_prev is not None else None
                self.plan.run(self.stage, self.offsets[lo:hi], thumb=thumbs, preview=prev, want_preview=prev is not None)
                self.kernel_launches += 1
                self.h_thumbs[lo:hi].copy_(thumbs, non_blocking=True)
                d2h += thumbs.numel()
                if prev is not None:
                    self.h_prev[lo:hi].copy_(prev, non_blocking=True)
                    d2h += prev.numel() * 4
        fin = self.final_stream
        for s in [self.copy_stream] + self.hash_streams + self.resize_streams:
            fin.wait_stream(s)
        with torch.cuda.stream(fin):
            is_new, first, last, counts = engine.dedupe_device(self.d_digests[:n], existing_sorted=existing_sorted)
            self.kernel_launches += 2
            self.h_digests[:n].copy_(self.d_digests[:n], non_blocking=True)
            self.h_is_new[:n].copy_(is_new, non_blocking=True)
            self.h_counts.copy_(counts, non_blocking=True)
            d2h += n * 33 + 12
            self._done = torch.cuda.Event()
            self._done.record(fin)
        self._pending = (n, h2d, d2h)

    def result(self) -> PipelineResult:
        """Wait for the submitted batch and hand back the pinned host results."""
        n, h2d, d2h = self._pending
        self._done.synchronize()
        c = self.h_counts.tolist()
        return PipelineResult(self.h_digests[:n], self.h_is_new[:n],
                              {"processed": c[0], "created": c[1], "updated": c[2]},
                              self.h_thumbs[:n], self.h_prev[:n] if self.h_prev is not None else None, h2d, d2h)
