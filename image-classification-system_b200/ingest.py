"""Batched ingest: content hash -> dedupe decision (-> thumbnail / preview tensors).

This is the batching seam SURVEY.md section 8(f) names: download N, ONE device call, bulk
upsert.  The single-image reference functions (services/*.py in this package) are 1..50-element
batches of these.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from . import hostapi


@dataclass
class DedupeDecision:
    hashes: List[Optional[str]]           # lowercase 64-char hex, None for skipped entries
    is_new: List[bool]                    # the insert branch runs (webdav_sync.py:326-354)
    first_index: List[int]                # first valid occurrence of the same content in the batch (-1 = skipped)
    last_index: List[int]                 # last valid occurrence (nome_img / caminho_img come from it)
    stats: Dict[str, int] = field(default_factory=dict)   # {'processed','created','updated'} (:308)


def _hex_to_digests(hexes: Iterable[str]) -> np.ndarray:
    hs = list(hexes)
    if not hs:
        return np.zeros((0, 32), dtype=np.uint8)
    return np.frombuffer(bytes.fromhex("".join(hs)), dtype=np.uint8).reshape(-1, 32).copy()


def hash_and_dedupe(datas: Sequence[Optional[bytes]], existing_hashes=None,
                    device: Optional[int] = None) -> DedupeDecision:
    """``datas[i] is None`` marks an image skipped before the lookup (invalid extension/MIME,
    webdav_sync.py:314, or failed download, :320): hashed nowhere, counted nowhere.

    ``existing_hashes``: either an iterable of hex strings already in table ``imagens`` or a
    callable ``f(list_of_hex) -> iterable of those present`` (one ``IN`` query).
    """
    n = len(datas)
    hostapi.init(device)
    if n == 0:
        return DedupeDecision([], [], [], [], {"processed": 0, "created": 0, "updated": 0})
    valid_np = np.fromiter((d is not None for d in datas), dtype=np.uint8, count=n)
    digests, hex_all = hostapi.sha256_host([d if d is not None else b"" for d in datas], device)
    hashes: List[Optional[str]] = [h if v else None for h, v in zip(hex_all, valid_np)]

    present = [h for h in hashes if h is not None]
    if callable(existing_hashes):
        existing = set(existing_hashes(sorted(set(present))))
    else:
        existing = set(existing_hashes) if existing_hashes is not None else set()
    existing &= set(present)              # only the keys this batch can hit matter
    table = hostapi.sort_digests(_hex_to_digests(sorted(existing))) if existing else None
    is_new, first, last, c = hostapi.dedupe_host(digests, valid_np, table, device)
    return DedupeDecision(
        hashes=hashes,
        is_new=[bool(x) for x in is_new],
        first_index=[int(x) for x in first],
        last_index=[int(x) for x in last],
        stats={"processed": c[0], "created": c[1], "updated": c[2]},
    )


@dataclass
class IngestResult:
    decision: DedupeDecision
    thumbs: Optional[np.ndarray]          # uint8 [n, out_h, out_w, 3]
    previews: Optional[np.ndarray]        # float32 [n, 3, out_h, out_w]


def ingest_batch(datas: Sequence[Optional[bytes]], decoded_rgb: Optional[Sequence[Optional[np.ndarray]]] = None,
                 existing_hashes=None, out_h: int = 256, out_w: int = 256, want_preview: bool = True,
                 device: Optional[int] = None) -> IngestResult:
    """Hash + dedupe the file bytes and, when the caller supplies the decoded RGB pixels (decode
    stays in the reference's own host library, Pillow), resize them to thumbnails/previews."""
    decision = hash_and_dedupe(datas, existing_hashes, device)
    thumbs = previews = None
    if decoded_rgb is not None:
        idx = [i for i, im in enumerate(decoded_rgb) if im is not None]
        if idx:
            t, p = hostapi.thumbnails([decoded_rgb[i] for i in idx], out_h, out_w, want_preview, device=device)
            thumbs = np.zeros((len(datas), out_h, out_w, 3), dtype=np.uint8)
            thumbs[idx] = t
            if p is not None:
                previews = np.zeros((len(datas), 3, out_h, out_w), dtype=np.float32)
                previews[idx] = p
    return IngestResult(decision, thumbs, previews)
