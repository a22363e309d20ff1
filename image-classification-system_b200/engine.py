"""Device-level entry points: thin, typed wrappers over the C ABI (include/b2ingest.h).

PyTorch is used for device/pinned memory, streams and (in dist.py) the process group — plumbing
only.  Every function here launches hand-written sm_100a kernels from libb2ingest.so on the
current torch CUDA stream and returns without synchronising unless it hands back Python values.
There is no CPU path: without the library or without a GPU these raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import B2Error, check, lib

ALIGN = 16            # message starts are 16-byte aligned so the kernels use 128-bit loads
_tls = threading.local()


# --------------------------------------------------------------------------- plumbing
def init(device: Optional[int] = None) -> int:
    """Bind the calling thread to ``device`` (default: torch's current CUDA device) and check it
    is a Blackwell part.  Cheap after the first call per (thread, device)."""
    if not torch.cuda.is_available():
        # let the library produce the canonical message (and keep the no-fallback contract)
        check(lib.b2_init(0 if device is None else int(device)))
        raise B2Error(_lib.B2_ERR_NO_DEVICE, "no CUDA device")
    dev = torch.cuda.current_device() if device is None else int(device)
    # the cache only skips the capability check: the CURRENT device is compared on every call, because the caller
    # may have switched it in between (`with torch.cuda.device(1): ...`) and the library launches on the current one
    if torch.cuda.current_device() != dev:
        torch.cuda.set_device(dev)
    checked = getattr(_tls, "checked", None)
    if checked is None:
        checked = _tls.checked = set()
    if dev not in checked:
        torch.cuda.current_stream(dev)       # make sure torch created the primary context
        check(lib.b2_init(dev))
        checked.add(dev)
    return dev


def bind_host_thread_to_gpu_node(device: Optional[int] = None) -> Optional[int]:
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off, so that page-locked staging
    buffers allocated afterwards (first touch) live in the host memory closest to it.  With one process per
    GPU and every rank streaming ~50 GB/s of host memory, buffers on the wrong socket halve the end-to-end
    rate.  Returns the node, or None when sysfs does not say (single-socket box, container without sysfs)."""
    try:
        dev = torch.cuda.current_device() if device is None else int(device)
        pr = torch.cuda.get_device_properties(dev)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError):
        return None


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    assert t.is_contiguous(), "libb2ingest needs contiguous buffers"
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream(torch.cuda.current_device()).cuda_stream


def _need_cuda(*tensors: Optional[torch.Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise B2Error(_lib.B2_ERR_BAD_ARG, "expected a CUDA tensor (this path has no CPU fallback)")


def sm_count(device: Optional[int] = None) -> int:
    dev = init(device)
    out = C.c_int(0)
    check(lib.b2_device_sm_count(dev, C.byref(out)))
    return out.value


# --------------------------------------------------------------------------- packing
class PackedMessages:
    """A batch of byte strings laid out in ONE pinned host buffer, each start 16-byte aligned.

    ``order`` lists message indices by decreasing length: a warp hashes 32 consecutive slots of
    it, so its lanes finish together.
    """

    def __init__(self, datas: Sequence[bytes], pin: bool = True):
        n = len(datas)
        lengths = np.fromiter((len(d) for d in datas), dtype=np.int64, count=n)
        padded = (lengths + (ALIGN - 1)) & ~np.int64(ALIGN - 1)
        offsets = np.zeros(n, dtype=np.int64)
        if n > 1:
            np.cumsum(padded[:-1], out=offsets[1:])
        total = int(padded.sum()) if n else 0
        self.n = n
        self.total_bytes = max(total, ALIGN)
        use_pin = pin and torch.cuda.is_available()
        self.data = torch.zeros(self.total_bytes, dtype=torch.uint8, pin_memory=use_pin)
        host = self.data.numpy()
        for i, d in enumerate(datas):
            if lengths[i]:
                o = int(offsets[i])
                host[o:o + int(lengths[i])] = np.frombuffer(d, dtype=np.uint8)
        self.offsets = torch.from_numpy(offsets)
        self.lengths = torch.from_numpy(lengths)
        self.order = torch.from_numpy(np.argsort(-lengths, kind="stable").astype(np.int32))

    def to_device(self, device: Optional[int] = None):
        dev = torch.device("cuda", init(device))
        nb = True
        return (self.data.to(dev, non_blocking=nb), self.offsets.to(dev, non_blocking=nb),
                self.lengths.to(dev, non_blocking=nb), self.order.to(dev, non_blocking=nb))


# --------------------------------------------------------------------------- a1: hash
def sha256_device(data: torch.Tensor, offsets: torch.Tensor, lengths: torch.Tensor,
                  order: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """SHA-256 of n messages resident in HBM -> uint8[n, 32] digests (FIPS 180-4 byte order).
    ``offsets``/``lengths`` are int64[n] (byte offsets into ``data``), ``order`` int32[n] or None."""
    _need_cuda(data, offsets, lengths, order, out)
    init(data.device.index)
    n = offsets.numel()
    assert offsets.dtype == torch.int64 and lengths.dtype == torch.int64 and lengths.numel() == n
    assert data.dtype == torch.uint8
    if order is not None:
        assert order.dtype == torch.int32 and order.numel() == n
    if out is None:
        out = torch.empty((n, 32), dtype=torch.uint8, device=data.device)
    assert out.dtype == torch.uint8 and out.numel() == n * 32
    check(lib.b2_sha256_batch(_ptr(data), _ptr(offsets), _ptr(lengths), _ptr(order), n, _ptr(out), _stream()))
    return out


def digest_hex_device(digests: torch.Tensor) -> torch.Tensor:
    """uint8[n,32] -> uint8[n,64] lowercase ASCII hex (the String(64) primary key)."""
    _need_cuda(digests)
    init(digests.device.index)
    n = digests.numel() // 32
    out = torch.empty((n, 64), dtype=torch.uint8, device=digests.device)
    check(lib.b2_digest_hex(_ptr(digests), n, _ptr(out), _stream()))
    return out


def hex_strings(hex_dev: torch.Tensor) -> List[str]:
    """Device hex chars -> Python strings (one D2H copy, synchronises)."""
    host = hex_dev.cpu().numpy()
    if host.size == 0:
        return []
    flat = host.tobytes().decode("ascii")
    return [flat[i:i + 64] for i in range(0, len(flat), 64)]


from .hostapi import dedupe_host, hash_batch, sha256_host, sort_digests  # noqa: E402,F401  (host-pointer layer)


# --------------------------------------------------------------------------- a4: dedupe
def dedupe_device(digests: torch.Tensor, valid: Optional[torch.Tensor] = None,
                  seq: Optional[torch.Tensor] = None, existing_sorted: Optional[torch.Tensor] = None):
    """First-occurrence-wins resolution of a batch of digests (webdav_sync.py:311-400).
    Returns ``(is_new u8[n], first_index i32[n], last_index i32[n], counts u32[3])`` on device;
    counts = (processed, created, updated)."""
    _need_cuda(digests, valid, seq, existing_sorted)
    dev = digests.device
    init(dev.index)
    n = digests.numel() // 32
    is_new = torch.empty(n, dtype=torch.uint8, device=dev)
    first = torch.empty(n, dtype=torch.int32, device=dev)
    last = torch.empty(n, dtype=torch.int32, device=dev)
    counts = torch.empty(3, dtype=torch.int32, device=dev)
    ws_bytes = int(lib.b2_dedupe_workspace_bytes(n))
    ws = torch.empty(max(ws_bytes, 8) // 8, dtype=torch.int64, device=dev)
    m = 0 if existing_sorted is None else existing_sorted.numel() // 32
    check(lib.b2_dedupe(_ptr(digests), _ptr(valid), _ptr(seq), n,
                        _ptr(existing_sorted) if m else None, m,
                        _ptr(is_new), _ptr(first), _ptr(last), _ptr(counts), _ptr(ws), ws.numel() * 8, _stream()))
    return is_new, first, last, counts


def lookup_sorted_device(digests: torch.Tensor, existing_sorted: Optional[torch.Tensor]) -> torch.Tensor:
    """Position of each digest in the sorted table, or -1 (routes/images.py:65)."""
    _need_cuda(digests, existing_sorted)
    init(digests.device.index)
    n = digests.numel() // 32
    m = 0 if existing_sorted is None else existing_sorted.numel() // 32
    out = torch.empty(n, dtype=torch.int64, device=digests.device)
    check(lib.b2_lookup_sorted(_ptr(digests), n, _ptr(existing_sorted) if m else None, m, _ptr(out), _stream()))
    return out


# --------------------------------------------------------------------------- a12: resize
class ResizePlan:
    """Tap tables + launch geometry for one (in_h, in_w) -> (out_h, out_w) shape."""

    def __init__(self, in_h: int, in_w: int, out_h: int = 256, out_w: int = 256, device: Optional[int] = None):
        self.device = init(device)
        self.in_h, self.in_w, self.out_h, self.out_w = in_h, in_w, out_h, out_w
        handle = C.c_void_p()
        check(lib.b2_resize_plan_create(in_h, in_w, out_h, out_w, C.byref(handle)))
        self._h = handle

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.b2_resize_plan_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def taps(self, axis: int) -> Tuple[np.ndarray, np.ndarray, int]:
        out = self.out_w if axis == 0 else self.out_h
        ks = C.c_int(0)
        check(lib.b2_resize_plan_taps(self._h, axis, C.byref(ks), None, None, 0))
        bounds = np.zeros((out, 2), dtype=np.int32)
        coeffs = np.zeros((out, ks.value), dtype=np.int32)
        check(lib.b2_resize_plan_taps(self._h, axis, C.byref(ks), bounds.ctypes.data, coeffs.ctypes.data, coeffs.size))
        return bounds, coeffs, ks.value

    def run(self, rgb: torch.Tensor, offsets: torch.Tensor, thumb: Optional[torch.Tensor] = None,
            preview: Optional[torch.Tensor] = None, want_preview: bool = True,
            mean: Sequence[float] = (0.0, 0.0, 0.0), inv_std: Sequence[float] = (1.0, 1.0, 1.0),
            out_slot: Optional[torch.Tensor] = None):
        """rgb: uint8 device buffer holding n HWC images at byte ``offsets`` (int64[n]).
        Returns (thumb uint8[n,out_h,out_w,3], preview float32[n,3,out_h,out_w] or None)."""
        _need_cuda(rgb, offsets, thumb, preview, out_slot)
        init(rgb.device.index)
        n = offsets.numel()
        n_slots = n if out_slot is None else None
        if thumb is None:
            assert n_slots is not None, "pass `thumb` when using out_slot"
            thumb = torch.empty((n, self.out_h, self.out_w, 3), dtype=torch.uint8, device=rgb.device)
        if preview is None and want_preview:
            assert n_slots is not None, "pass `preview` when using out_slot"
            preview = torch.empty((n, 3, self.out_h, self.out_w), dtype=torch.float32, device=rgb.device)
        m = (C.c_float * 3)(*[float(x) for x in mean])
        s = (C.c_float * 3)(*[float(x) for x in inv_std])
        check(lib.b2_resize_normalize_batch(self._h, _ptr(rgb), _ptr(offsets), _ptr(out_slot), n,
                                            _ptr(thumb), _ptr(preview), m, s, _stream()))
        return thumb, preview


_plans: Dict[Tuple[int, int, int, int, int], ResizePlan] = {}
_plans_lock = threading.Lock()


def get_plan(in_h: int, in_w: int, out_h: int = 256, out_w: int = 256, device: Optional[int] = None) -> ResizePlan:
    dev = init(device)
    key = (dev, in_h, in_w, out_h, out_w)
    with _plans_lock:
        p = _plans.get(key)
        if p is None:
            p = _plans[key] = ResizePlan(in_h, in_w, out_h, out_w, dev)
        return p


def thumbnails(images: Sequence[np.ndarray], out_h: int = 256, out_w: int = 256, want_preview: bool = True,
               mean: Sequence[float] = (0.0, 0.0, 0.0), inv_std: Sequence[float] = (1.0, 1.0, 1.0),
               device: Optional[int] = None):
    """Host convenience: decoded RGB HWC uint8 arrays (any mix of shapes) -> (thumb uint8
    [n,out_h,out_w,3], preview float32 [n,3,out_h,out_w] or None) as NumPy arrays.  Images are
    grouped by shape; each group is one device call."""
    dev = torch.device("cuda", init(device))
    n = len(images)
    thumb = torch.empty((n, out_h, out_w, 3), dtype=torch.uint8, device=dev)
    preview = torch.empty((n, 3, out_h, out_w), dtype=torch.float32, device=dev) if want_preview else None
    groups: Dict[Tuple[int, int], List[int]] = {}
    for i, im in enumerate(images):
        assert im.dtype == np.uint8 and im.ndim == 3 and im.shape[2] == 3, "expected HxWx3 uint8"
        groups.setdefault((im.shape[0], im.shape[1]), []).append(i)
    for (h, w), idxs in groups.items():
        packed = PackedMessages([np.ascontiguousarray(images[i]).tobytes() for i in idxs])
        d_data, d_off, _, _ = packed.to_device(dev.index)
        slots = torch.tensor(idxs, dtype=torch.int32).to(dev)
        get_plan(h, w, out_h, out_w, dev.index).run(d_data, d_off, thumb=thumb, preview=preview,
                                                   want_preview=want_preview, mean=mean, inv_std=inv_std,
                                                   out_slot=slots)
    torch.cuda.current_stream().synchronize()
    return thumb.cpu().numpy(), (preview.cpu().numpy() if preview is not None else None)


# --------------------------------------------------------------------------- a13: tally
def label_tally_device(image_idx: torch.Tensor, class_idx: torch.Tensor, active: torch.Tensor,
                       n_images: int, k: int, image_base: int = 0, sorted_by_image: bool = True,
                       counts: Optional[torch.Tensor] = None, partials: Optional[torch.Tensor] = None,
                       agree_hist: Optional[torch.Tensor] = None):
    """Per-image class tally of ACTIVE rows + integer Fleiss partials.  Rows are SoA device
    tensors (int32, uint8, uint8).  Returns ``(counts int32[n_images,k], partials int64[k+7])``;
    never synchronises — call :func:`check_tally` on the host copy of ``partials``.
    ``agree_hist``: optional int64[B2_AGREE_BINS] device tensor receiving the agreement histogram (the
    general-n kappa from integers, :func:`labels.fleiss_kappa_from_hist`).  ``partials`` may be a view of a
    larger int64 buffer whose tail is ``agree_hist`` so that ONE all-reduce covers both."""
    _need_cuda(image_idx, class_idx, active, counts, partials, agree_hist)
    dev = image_idx.device
    init(dev.index)
    rows = image_idx.numel()
    assert image_idx.dtype == torch.int32 and class_idx.dtype == torch.uint8 and active.dtype == torch.uint8
    assert class_idx.numel() == rows and active.numel() == rows
    if counts is None:
        counts = torch.empty((n_images, k), dtype=torch.int32, device=dev)
    if partials is None:
        partials = torch.empty(k + _lib.B2_PARTIALS_EXTRA, dtype=torch.int64, device=dev)
    flags = _lib.B2_TALLY_SORTED if sorted_by_image else 0
    check(lib.b2_label_tally(_ptr(image_idx), _ptr(class_idx), _ptr(active), rows, image_base, n_images, k, flags,
                             _ptr(counts), _ptr(partials), _ptr(agree_hist), _stream()))
    return counts, partials


from .hostapi import PARTIAL_NAMES, check_tally, partials_dict  # noqa: E402,F401


def encode_label_rows_device(img_hex: torch.Tensor, opc_uuid: torch.Tensor, ativo: torch.Tensor,
                             image_keys_sorted: torch.Tensor, option_keys_sorted: torch.Tensor):
    """``b2_encode_label_rows``: rows of table ``classificacoes`` as raw keys on the device — ``img_hex`` uint8[R,64]
    (the char64 id_img), ``opc_uuid`` uint8[R,16], ``ativo`` uint8[R] — against the sorted image-digest table
    uint8[N,32] and the sorted option UUIDs uint8[k,16].  Returns ``(image_idx int32[R], class_idx uint8[R],
    active uint8[R], unknown int64[2])``; image -1 / class 255 mark keys missing from a dictionary."""
    _need_cuda(img_hex, opc_uuid, ativo, image_keys_sorted, option_keys_sorted)
    dev = img_hex.device
    init(dev.index)
    rows = ativo.numel()
    assert img_hex.dtype == torch.uint8 and img_hex.numel() == rows * 64
    assert opc_uuid.dtype == torch.uint8 and opc_uuid.numel() == rows * 16 and ativo.dtype == torch.uint8
    n_images = image_keys_sorted.numel() // 32
    k = option_keys_sorted.numel() // 16
    image_idx = torch.empty(rows, dtype=torch.int32, device=dev)
    class_idx = torch.empty(rows, dtype=torch.uint8, device=dev)
    active = torch.empty(rows, dtype=torch.uint8, device=dev)
    unknown = torch.empty(2, dtype=torch.int64, device=dev)
    check(lib.b2_encode_label_rows(_ptr(img_hex), _ptr(opc_uuid), _ptr(ativo), rows,
                                   _ptr(image_keys_sorted) if n_images else None, n_images,
                                   _ptr(option_keys_sorted) if k else None, k,
                                   _ptr(image_idx), _ptr(class_idx), _ptr(active), _ptr(unknown), _stream()))
    return image_idx, class_idx, active, unknown


def fleiss_partials_device(counts: torch.Tensor, want_sum_pi: bool = False, agree_hist: Optional[torch.Tensor] = None):
    """Partials (and optionally the float64 sum of per-image agreements P_i, and the integer agreement histogram)
    from a count matrix."""
    _need_cuda(counts)
    dev = counts.device
    init(dev.index)
    assert counts.dtype == torch.int32 and counts.dim() == 2
    n_images, k = counts.shape
    partials = torch.empty(k + _lib.B2_PARTIALS_EXTRA, dtype=torch.int64, device=dev)
    sum_pi = ws = None
    ws_bytes = 0
    if want_sum_pi:
        sum_pi = torch.zeros(1, dtype=torch.float64, device=dev)
        ws_bytes = int(lib.b2_fleiss_workspace_bytes(n_images))
        ws = torch.empty(ws_bytes // 8 + 1, dtype=torch.int64, device=dev)
    check(lib.b2_fleiss_partials(_ptr(counts), n_images, k, _ptr(partials), _ptr(sum_pi), _ptr(agree_hist), _ptr(ws),
                                 ws_bytes, _stream()))
    return partials, sum_pi


def distinct_images_per_annotator_device(annotator_idx: torch.Tensor, image_idx: torch.Tensor,
                                         active: torch.Tensor, n_annotators: int) -> torch.Tensor:
    """Bulk COUNT(DISTINCT id_img) WHERE id_con=? AND ativo (routes/classificacoes.py:224-230);
    rows sorted by (annotator_idx, image_idx)."""
    _need_cuda(annotator_idx, image_idx, active)
    init(image_idx.device.index)
    out = torch.empty(n_annotators, dtype=torch.int32, device=image_idx.device)
    check(lib.b2_distinct_images_per_annotator(_ptr(annotator_idx), _ptr(image_idx), _ptr(active),
                                               image_idx.numel(), n_annotators, _ptr(out), _stream()))
    return out
