"""Streaming ingest of listings held in HOST memory: the end-to-end form of the path.

    host images --H2D--> [ sha256 | resize + normalise ] --D2H--> digests, thumbnails, previews (listing order)
                                   \\--> dedupe over the listing --D2H--> flags + stats

The pipeline itself is native: ``b2_ingest_ring_*`` in ``csrc/ring.cu`` (C ABI, host pointers only).  SHA-256 is
serial per message — one message moves at ~60 MB/s whatever else the GPU does — so PCIe (55 GB/s) is only kept
busy when about a thousand messages hash at once.  The ring therefore bounds its depth in BYTES: a device staging
ring of tens of GB carved into chunks of consecutive listing entries; each chunk is one H2D burst, one resize launch
per shape and one read-back per output kind straight into the listing-order slots of the caller's buffers; every ~4 GiB
of chunks is ONE hash launch; ``submit`` blocks only while the ring is full, listings complete independently
(``wait`` / ``poll``).  Any mix of image sizes (BASELINE config 3) goes through the same ring.

This module only wraps the C calls: it allocates the page-locked result buffers (``b2_host_alloc``) and keeps
inputs alive while a listing is in flight.  It imports NumPy only — the reference service has no PyTorch.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import hostapi
from ._lib import check, lib


def _addr(a) -> int:
    """Address of a host buffer: NumPy array, PyTorch CPU tensor, or anything with the buffer protocol."""
    if a is None:
        return 0
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return int(a.data_ptr())
    return np.frombuffer(a, dtype=np.uint8).ctypes.data


@dataclass
class ListingResult:
    """Everything in listing order; the arrays are views of the ring's page-locked result buffers and stay valid
    until the slot is reused (``max_listings`` submits later) — copy what must live longer."""
    digests: np.ndarray                   # uint8 [n, 32]
    is_new: np.ndarray                    # uint8 [n]
    first_index: np.ndarray               # int32 [n]: first occurrence of the same content in the listing (-1 = skipped)
    last_index: np.ndarray                # int32 [n]
    stats: Dict[str, int]                 # {'processed','created','updated'} (webdav_sync.py:308)
    thumbs: Optional[np.ndarray]          # uint8 [n, out_h, out_w, 3]
    previews: Optional[np.ndarray]        # float32 [n, 3, out_h, out_w]
    h2d_bytes: int
    d2h_bytes: int
    kernel_launches: int


class _Slot:
    def __init__(self, cap: int, out_h: int, out_w: int, want_preview: bool):
        self.cap = cap
        self.digests = hostapi.pinned_empty((cap, 32), np.uint8)
        self.is_new = hostapi.pinned_empty((cap,), np.uint8)
        self.first = hostapi.pinned_empty((cap,), np.int32)
        self.last = hostapi.pinned_empty((cap,), np.int32)
        self.counts = hostapi.pinned_empty((4,), np.uint32)
        self.thumbs = hostapi.pinned_empty((cap, out_h, out_w, 3), np.uint8)
        self.previews = hostapi.pinned_empty((cap, 3, out_h, out_w), np.float32) if want_preview else None
        self.keep = None                  # inputs of the listing in flight
        self.ticket = 0
        self.n = 0
        self.has_pixels = False


class IngestRing:
    """``b2_ingest_ring``: byte-bounded, shape-agnostic streaming ingest.

    ``submit(...)`` returns a ticket at once (it blocks only while the device ring is full); ``result(ticket)`` waits
    for that listing.  Up to ``max_listings`` listings in flight; keep at least two so the hash tail of one listing
    (a 50 MB file hashes for 0.8 s) hides under the copies of the next.
    """

    def __init__(self, ring_bytes: int = 32 << 30, chunk_bytes: int = 0, max_listings: int = 4,
                 max_images: int = 4096, out_h: int = 256, out_w: int = 256, want_preview: bool = True,
                 device: Optional[int] = None):
        self.device = hostapi.init(device)
        self.out_h, self.out_w, self.want_preview = out_h, out_w, want_preview
        self.max_listings, self.max_images = max_listings, max_images
        h = C.c_void_p()
        check(lib.b2_ingest_ring_create(self.device, int(ring_bytes), int(chunk_bytes), max_listings, out_h, out_w,
                                        1 if want_preview else 0, C.byref(h)))
        self._h = h
        self._slots: List[_Slot] = []
        self._free: List[_Slot] = []
        self._inflight: Dict[int, _Slot] = {}

    def prepare(self, n_images: Optional[int] = None, n_slots: Optional[int] = None) -> None:
        """Allocate the page-locked result buffers up front (``n_slots`` listings of up to ``n_images`` entries each;
        defaults: ``max_images`` / ``max_listings``).  Page-locking ~1 MB per image costs ~0.25 s per GB, which would
        otherwise be paid by the first ``submit`` that needs a new slot.  Frees idle slots of another size."""
        assert not self._inflight, "prepare() with listings in flight"
        n_images = n_images or self.max_images
        n_slots = min(n_slots or self.max_listings, self.max_listings)
        keep = [s for s in self._free if s.cap == n_images][:n_slots]
        self._slots = self._free = None                      # drop the others before allocating
        self._slots, self._free = list(keep), list(keep)
        while len(self._slots) < n_slots:
            s = _Slot(n_images, self.out_h, self.out_w, self.want_preview)
            self._slots.append(s)
            self._free.append(s)

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.b2_ingest_ring_destroy(self._h)
            self._h = None
            self._slots, self._free, self._inflight = [], [], {}

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def _slot(self, n: int) -> _Slot:
        for i, s in enumerate(self._free):
            if s.cap >= n:
                return self._free.pop(i)
        if len(self._slots) >= self.max_listings and self._free:
            self._slots.remove(self._free.pop(0))             # too small: replace it
        if len(self._slots) >= self.max_listings:
            raise hostapi.B2Error(-1, f"{self.max_listings} listings already in flight (call result() first)")
        s = _Slot(max(n, self.max_images), self.out_h, self.out_w, self.want_preview)
        self._slots.append(s)
        return s

    # ------------------------------------------------------------------------------------------------------
    def submit(self, pixels: Optional[Sequence] = None, shapes=None, files: Optional[Sequence] = None,
               valid=None, existing_sorted=None) -> int:
        """One listing.  ``pixels[i]``: decoded RGB HWC uint8 buffer of image i (NumPy array, CPU tensor, bytes) or
        None; ``shapes``: int array [n, 2] of (h, w) (taken from the arrays when omitted).  ``files[i]``: the
        downloaded file bytes (the message that is hashed); omitted = the pixel buffer is the message (BASELINE's
        synthetic images).  ``valid[i] == 0``: skipped before the lookup.  ``existing_sorted``: uint8 [m, 32] digests
        already stored, memcmp order.  Page-locked inputs make the copies asynchronous."""
        n = len(pixels) if pixels is not None else len(files)
        slot = self._slot(n)
        keep = [pixels, files]
        px_ptrs = hw = f_ptrs = f_lens = None
        if pixels is not None:
            px_ptrs = (C.c_void_p * n)(*[_addr(p) or None for p in pixels])
            if shapes is None:
                shapes = [(p.shape[0], p.shape[1]) if p is not None else (0, 0) for p in pixels]
            hw = np.ascontiguousarray(shapes, dtype=np.uint32).reshape(n, 2)
        if files is not None:
            f_ptrs = (C.c_void_p * n)(*[_addr(f) or None for f in files])
            f_lens = np.array([0 if f is None else (f.nbytes if hasattr(f, "nbytes") else len(f)) for f in files],
                              dtype=np.uint64)
        v = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
        ex, m = None, 0
        if existing_sorted is not None and len(existing_sorted):
            ex = np.ascontiguousarray(np.asarray(existing_sorted, dtype=np.uint8)).reshape(-1, 32)
            m = ex.shape[0]
        keep += [px_ptrs, hw, f_ptrs, f_lens, v, ex]
        ticket = C.c_uint64()
        try:
            check(lib.b2_ingest_ring_submit(
                self._h, C.cast(px_ptrs, C.c_void_p) if px_ptrs is not None else None,
                hw.ctypes.data if hw is not None else None,
                C.cast(f_ptrs, C.c_void_p) if f_ptrs is not None else None,
                f_lens.ctypes.data if f_lens is not None else None,
                v.ctypes.data if v is not None else None, n, ex.ctypes.data if ex is not None else None, m,
                slot.digests.ctypes.data, slot.is_new.ctypes.data, slot.first.ctypes.data, slot.last.ctypes.data,
                slot.counts.ctypes.data, slot.thumbs.ctypes.data if pixels is not None else None,
                slot.previews.ctypes.data if (slot.previews is not None and pixels is not None) else None,
                C.byref(ticket)))
        except Exception:
            self._free.append(slot)                           # the library drained everything before returning
            raise
        slot.keep, slot.ticket, slot.n, slot.has_pixels = keep, int(ticket.value), n, pixels is not None
        self._inflight[slot.ticket] = slot
        return slot.ticket

    def submit_packed(self, images, shape: Tuple[int, int], existing_sorted=None) -> int:
        """Fixed-shape convenience: ``images`` is ONE contiguous uint8 host buffer [n, h*w*3] (NumPy or CPU tensor)."""
        n = int(images.shape[0])
        L = shape[0] * shape[1] * 3
        base = _addr(images)
        ptrs = [base + i * L for i in range(n)]
        t = self.submit(_RawPointers(ptrs, images), np.tile(np.asarray(shape, dtype=np.uint32), (n, 1)),
                        existing_sorted=existing_sorted)
        return t

    def poll(self, ticket: int) -> Tuple[bool, int]:
        """(done, leading entries whose thumbnails / previews are already in the host buffers)."""
        done, flushed = C.c_int(), C.c_uint32()
        check(lib.b2_ingest_ring_poll(self._h, ticket, C.byref(done), C.byref(flushed)))
        return bool(done.value), int(flushed.value)

    def result(self, ticket: int) -> ListingResult:
        slot = self._inflight.pop(ticket)
        h2d, d2h, launches = C.c_uint64(), C.c_uint64(), C.c_uint32()
        try:
            check(lib.b2_ingest_ring_wait(self._h, ticket, C.byref(h2d), C.byref(d2h), C.byref(launches)))
        finally:
            slot.keep = None
            self._free.append(slot)
        n, c = slot.n, slot.counts
        return ListingResult(slot.digests[:n], slot.is_new[:n], slot.first[:n], slot.last[:n],
                             {"processed": int(c[0]), "created": int(c[1]), "updated": int(c[2])},
                             slot.thumbs[:n] if slot.has_pixels else None,
                             slot.previews[:n] if (slot.has_pixels and slot.previews is not None) else None,
                             int(h2d.value), int(d2h.value), int(launches.value))

    def stats(self) -> Dict[str, int]:
        rb, fl, ch, st = C.c_uint64(), C.c_uint64(), C.c_uint32(), C.c_uint64()
        check(lib.b2_ingest_ring_stats(self._h, C.byref(rb), C.byref(fl), C.byref(ch), C.byref(st)))
        return {"ring_bytes": rb.value, "bytes_in_flight": fl.value, "chunks_in_flight": ch.value, "stalls": st.value}


class _RawPointers:
    """A list of raw addresses into one owner buffer (kept alive with the listing)."""

    def __init__(self, ptrs: List[int], owner):
        self.ptrs, self.owner = ptrs, owner

    def __len__(self):
        return len(self.ptrs)

    def __iter__(self):
        return (_Raw(p) for p in self.ptrs)


class _Raw:
    __slots__ = ("p",)

    def __init__(self, p: int):
        self.p = p

    def data_ptr(self) -> int:
        return self.p


class IngestPipeline:
    """Fixed-shape batches through a private :class:`IngestRing` (one batch in flight per pipeline; two pipelines
    used alternately overlap the hash tail of one batch with the copies of the next)."""

    def __init__(self, in_h: int, in_w: int, max_images: int, chunk_images: int = 256, out_h: int = 256,
                 out_w: int = 256, want_preview: bool = True, device: Optional[int] = None):
        self.in_h, self.in_w = in_h, in_w
        self.L = in_h * in_w * 3
        self.max_images = max_images
        per_image = self.L + out_h * out_w * 3 * (5 if want_preview else 1) + 64
        chunk = max(1, min(chunk_images, max_images))
        self.ring = IngestRing(ring_bytes=max(64 << 20, int(max_images * per_image * 1.25) + (8 << 20)),
                               chunk_bytes=chunk * self.L, max_listings=1, max_images=max_images, out_h=out_h,
                               out_w=out_w, want_preview=want_preview, device=device)
        self.kernel_launches = 0
        self._ticket = None

    def close(self) -> None:
        self.ring.close()

    def submit(self, host_images, existing_sorted=None) -> None:
        n = int(host_images.shape[0])
        assert n <= self.max_images
        if hasattr(existing_sorted, "numpy"):
            existing_sorted = existing_sorted.cpu().numpy()
        self._ticket = self.ring.submit_packed(host_images, (self.in_h, self.in_w), existing_sorted)

    def result(self) -> ListingResult:
        res = self.ring.result(self._ticket)
        self._ticket = None
        self.kernel_launches = res.kernel_launches
        return res

    def run(self, host_images, existing_sorted=None) -> ListingResult:
        self.submit(host_images, existing_sorted)
        return self.result()
