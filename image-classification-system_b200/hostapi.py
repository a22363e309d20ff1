"""Host-pointer layer: everything the reference service needs from the path, with ctypes and NumPy only.

The reference has no tensor library in its requirements.  Every function here is one call into
``csrc/host.cu`` (``b2_sha256_host``, ``b2_dedupe_host``, ``b2_thumbnails_host``, ``b2_label_tally_host``):
host buffers in, host buffers out, the library owns staging memory and streams.  Importing this module (or the
package itself, ``services/``, ``crud/``, ``api/``) does not import PyTorch; the device-pointer layer
(``engine``, ``labels`` on device tensors, ``pipeline``, ``dist``) does, and is loaded on first use.
There is no CPU path: without the library or without a Blackwell GPU these raise ``B2Error``.
"""
from __future__ import annotations

import ctypes as C
import sys
import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import B2Error, check, lib  # noqa: F401

_tls = threading.local()


def init(device: Optional[int] = None) -> int:
    """Bind the calling thread to ``device`` (default: device 0, or PyTorch's current device when PyTorch is
    already loaded in this process) and check it is a Blackwell part.  Cheap after the first call."""
    if "torch" in sys.modules and sys.modules["torch"] is not None:
        from . import engine                       # keep PyTorch's and the library's idea of the device in step
        return engine.init(device)
    dev = 0 if device is None else int(device)
    if getattr(_tls, "device", None) != dev:
        check(lib.b2_init(dev))
        _tls.device = dev
    return dev


def pinned_empty(shape, dtype=np.uint8) -> np.ndarray:
    """NumPy array in page-locked host memory (``b2_host_alloc``): copies to and from it are asynchronous and run
    at full PCIe speed.  Freed when the array (and every view of it) is garbage-collected."""
    import weakref
    dt = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
    p = C.c_void_p()
    check(lib.b2_host_alloc(C.byref(p), max(nbytes, 1)))
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
    weakref.finalize(buf, lib.b2_host_free, C.c_void_p(p.value))
    return np.frombuffer(buf, dtype=dt, count=nbytes // dt.itemsize).reshape(shape)


def shutdown() -> None:
    """``b2_shutdown``: release the library's pool of page-locked staging buffers and its cached resize plans."""
    check(lib.b2_shutdown())


def sort_digests(digests: np.ndarray) -> np.ndarray:
    """Sort uint8[m,32] digests in memcmp order (the order ``b2_dedupe`` expects for the table of stored
    digests).  Index maintenance, not on the hot path."""
    if digests.size == 0:
        return digests.reshape(0, 32)
    d = np.ascontiguousarray(digests.reshape(-1, 32))
    keys = d.view(">u8").reshape(-1, 4)
    idx = np.lexsort((keys[:, 3], keys[:, 2], keys[:, 1], keys[:, 0]))
    return np.ascontiguousarray(d[idx])


def sha256_host(datas: Sequence[bytes], device: Optional[int] = None, want_hex: bool = True):
    """``b2_sha256_host``: byte strings in host memory -> (digests uint8[n,32], hex strings or None).  The
    library packs the messages into its own page-locked staging buffer."""
    dev = init(device)
    datas = [d if isinstance(d, bytes) else bytes(d) for d in datas]
    n = len(datas)
    digests = np.empty((n, 32), dtype=np.uint8)
    if n == 0:
        return digests, ([] if want_hex else None)
    ptrs = (C.c_char_p * n)(*datas)                     # the bytes objects' own buffers, no copy
    lens = (C.c_uint64 * n)(*[len(d) for d in datas])
    hexbuf = C.create_string_buffer(n * 64) if want_hex else None
    check(lib.b2_sha256_host(dev, C.cast(ptrs, C.c_void_p), C.cast(lens, C.c_void_p), n, digests.ctypes.data,
                             C.cast(hexbuf, C.c_void_p) if want_hex else None))
    if not want_hex:
        return digests, None
    flat = hexbuf.raw.decode("ascii")
    return digests, [flat[i:i + 64] for i in range(0, n * 64, 64)]


class _HashCoalescer:
    """Merges CONCURRENT small ``hash_batch`` calls into one device launch.

    The reference hashes one file per call from up to five service threads at once (initial sync + its WebDAV and
    Activity workers, the scheduler's two loops: app/main.py:188-226) plus request threads.  One message costs the GPU
    as long as 32 of them, so calls that arrive while a launch is in flight are queued and the next launch takes them
    all: the first caller in becomes the leader, runs ONE ``b2_sha256_host`` over everything queued and hands the
    results out; followers wait (and take over leadership if the leader has left).  No artificial delay: a lone caller
    pays nothing, the batching comes from the launch latency itself."""

    def __init__(self):
        self.lock = threading.Lock()
        self.queue: list = []                      # entries: [datas, result, error, done-event]
        self.leader_active = False
        self.calls = 0                             # hash_batch calls served
        self.launches = 0                          # device launches issued for them

    def _round(self, device) -> None:
        with self.lock:
            batch, self.queue = self.queue, []
        if not batch:
            return
        try:
            flat = [d for e in batch for d in e[0]]
            hexes = sha256_host(flat, device)[1]
            self.launches += 1
            o = 0
            for e in batch:
                e[1] = hexes[o:o + len(e[0])]
                o += len(e[0])
        except BaseException as err:  # noqa: BLE001
            if len(batch) == 1:
                batch[0][2] = err
            else:                                             # one caller's bad input must not fail the others: one by one
                for e in batch:
                    try:
                        e[1] = sha256_host(e[0], device)[1]
                        self.launches += 1
                    except BaseException as own:  # noqa: BLE001 - reaches the caller that caused it
                        e[2] = own
        for e in batch:
            e[3].set()

    def hash(self, datas: Sequence[bytes], device) -> List[str]:
        entry = [list(datas), None, None, threading.Event()]
        with self.lock:
            self.queue.append(entry)
            self.calls += 1
        while not entry[3].is_set():
            with self.lock:
                lead = not self.leader_active
                if lead:
                    self.leader_active = True
            if lead:
                try:
                    self._round(device)
                finally:
                    with self.lock:
                        self.leader_active = False
            else:
                entry[3].wait(0.002)               # woken by the leader; re-checks leadership if it has left
        if entry[2] is not None:
            raise entry[2]
        return entry[1]


_coalescers: dict = {}
_coalescers_lock = threading.Lock()
COALESCE_MAX_MESSAGES = 64                         # larger batches fill a launch on their own


def hash_batch(datas: Sequence[bytes], device: Optional[int] = None) -> List[str]:
    """Batched form of ``hashlib.sha256(data).hexdigest()`` (reference: webdav_sync.py:59,
    activity_api_sync.py:798, routes/images.py:62) for a list of host byte strings.  Small batches from concurrent
    threads share device launches (:class:`_HashCoalescer`)."""
    if len(datas) == 0 or len(datas) > COALESCE_MAX_MESSAGES:
        return sha256_host(datas, device)[1]
    dev = init(device)
    with _coalescers_lock:
        c = _coalescers.get(dev)
        if c is None:
            c = _coalescers[dev] = _HashCoalescer()
    return c.hash(datas, dev)


def dedupe_host(digests: np.ndarray, valid: Optional[np.ndarray] = None, existing_sorted: Optional[np.ndarray] = None,
                device: Optional[int] = None):
    """``b2_dedupe_host``: digests uint8[n,32] (+ validity flags, + the sorted table of stored digests) in host
    memory -> ``(is_new u8[n], first_index i32[n], last_index i32[n], (processed, created, updated))``."""
    dev = init(device)
    d = np.ascontiguousarray(digests, dtype=np.uint8).reshape(-1, 32)
    n = d.shape[0]
    v = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
    ex = None if existing_sorted is None or existing_sorted.size == 0 else np.ascontiguousarray(existing_sorted, dtype=np.uint8)
    is_new = np.zeros(n, dtype=np.uint8)
    first = np.full(n, -1, dtype=np.int32)
    last = np.full(n, -1, dtype=np.int32)
    counts = np.zeros(4, dtype=np.uint32)
    check(lib.b2_dedupe_host(dev, d.ctypes.data if n else None, v.ctypes.data if v is not None else None, n,
                             ex.ctypes.data if ex is not None else None, 0 if ex is None else ex.size // 32,
                             is_new.ctypes.data, first.ctypes.data, last.ctypes.data, counts.ctypes.data))
    return is_new, first, last, (int(counts[0]), int(counts[1]), int(counts[2]))


def thumbnails(images: Sequence[np.ndarray], out_h: int = 256, out_w: int = 256, want_preview: bool = True,
               mean: Sequence[float] = (0.0, 0.0, 0.0), inv_std: Sequence[float] = (1.0, 1.0, 1.0),
               device: Optional[int] = None) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """``b2_thumbnails_host``: decoded RGB HWC uint8 arrays (any mix of shapes) -> (thumb uint8[n,out_h,out_w,3],
    preview float32[n,3,out_h,out_w] or None): Pillow's BILINEAR resize, bit-exact, and the normalised tensor."""
    dev = init(device)
    n = len(images)
    thumb = np.empty((n, out_h, out_w, 3), dtype=np.uint8)
    prev = np.empty((n, 3, out_h, out_w), dtype=np.float32) if want_preview else None
    if n == 0:
        return thumb, prev
    arrs = []
    for im in images:
        a = np.ascontiguousarray(im, dtype=np.uint8)
        if a.ndim != 3 or a.shape[2] != 3:
            raise B2Error(_lib.B2_ERR_BAD_ARG, f"expected an HxWx3 uint8 image, got shape {a.shape}")
        arrs.append(a)
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    hw = np.array([[a.shape[0], a.shape[1]] for a in arrs], dtype=np.uint32)
    m = (C.c_float * 3)(*[float(x) for x in mean])
    s = (C.c_float * 3)(*[float(x) for x in inv_std])
    check(lib.b2_thumbnails_host(dev, C.cast(ptrs, C.c_void_p), hw.ctypes.data, n, out_h, out_w, thumb.ctypes.data,
                                 prev.ctypes.data if prev is not None else None, m, s))
    return thumb, prev


def label_tally_host(image_idx, class_idx, active, n_images: int, k: int, sorted_by_image: bool = True,
                     image_base: int = 0, device: Optional[int] = None, want_counts: bool = True,
                     agree_hist: Optional[np.ndarray] = None):
    """Rows in host memory (anything ``np.asarray`` accepts) through ``b2_label_tally_host``: host pointers in,
    ``(counts int32[n_images,k] or None, partials int64[k+7])`` out; raises ``B2Error`` for unsorted rows
    (sorted mode) or rows out of range.  ``agree_hist``: optional int64[B2_AGREE_BINS] array that receives the
    agreement histogram (general-n kappa from integers)."""
    dev = init(device)
    img = np.ascontiguousarray(image_idx, dtype=np.int32)
    cls = np.ascontiguousarray(class_idx, dtype=np.uint8)
    act = np.ascontiguousarray(active, dtype=np.uint8)
    assert img.ndim == 1 and cls.shape == img.shape and act.shape == img.shape
    counts = np.empty((n_images, k), dtype=np.int32) if want_counts else None
    partials = np.empty(k + _lib.B2_PARTIALS_EXTRA, dtype=np.int64)
    check(lib.b2_label_tally_host(dev, img.ctypes.data, cls.ctypes.data, act.ctypes.data, img.size, image_base,
                                  n_images, k, _lib.B2_TALLY_SORTED if sorted_by_image else 0,
                                  counts.ctypes.data if counts is not None else None, partials.ctypes.data,
                                  agree_hist.ctypes.data if agree_hist is not None else None))
    return counts, partials


def distinct_images_host(annotator_idx, image_idx, active, n_annotators: int, device: Optional[int] = None) -> np.ndarray:
    """``b2_distinct_images_host``: rows sorted by (annotator, image) in host memory -> distinct active images
    per annotator, uint32[n_annotators] (bulk form of routes/classificacoes.py:224-230)."""
    dev = init(device)
    a = np.ascontiguousarray(annotator_idx, dtype=np.int32)
    i = np.ascontiguousarray(image_idx, dtype=np.int32)
    act = np.ascontiguousarray(active, dtype=np.uint8)
    assert a.ndim == 1 and i.shape == a.shape and act.shape == a.shape
    out = np.zeros(n_annotators, dtype=np.uint32)
    check(lib.b2_distinct_images_host(dev, a.ctypes.data, i.ctypes.data, act.ctypes.data, a.size, n_annotators,
                                      out.ctypes.data))
    return out


PARTIAL_NAMES = ("S2", "R", "n_rated", "n_pairs_images", "pairs", "rows_seen", "unsorted_pairs")


def check_tally(partials_host: np.ndarray, k: int, rows: int) -> None:
    """Raise B2Error if the tally met unsorted rows (sorted mode) or out-of-range rows."""
    p = np.ascontiguousarray(partials_host, dtype=np.int64)
    check(lib.b2_label_tally_status(p.ctypes.data, k, rows))


def partials_dict(partials_host: np.ndarray, k: int) -> dict:
    p = np.asarray(partials_host, dtype=np.int64)
    out: dict = {"class_totals": p[:k].copy()}
    for i, name in enumerate(PARTIAL_NAMES):
        out[name] = int(p[k + i])
    return out
