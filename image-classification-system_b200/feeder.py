"""Download / decode feeder: the step BEFORE the hot path.

The reference downloads one file, hashes it, touches the table, then downloads the next
(``app/services/webdav_sync.py:311-321`` inside the 50-image batch loop ``:273-283``; every
``NextCloudClient.get_file`` is a blocking HTTP GET with a 60 s timeout,
``app/services/nextcloud_service.py:384-422``).  The device calls of this package want whole
batches, so the feeder turns the listing into a stream of ready batches:

  * downloads run on ``download_workers`` threads (network bound, the GIL is released in the
    socket reads);
  * each downloaded file is header-parsed (the reference's ``_get_image_metadata``) and, when
    thumbnails are wanted, decoded to RGB HWC uint8 by Pillow on ``decode_workers`` threads
    (Pillow releases the GIL while decoding) — decode stays in the reference's own host library
    (north star), nothing here touches the GPU;
  * at most ``prefetch_batches`` batches are in flight beyond the one being consumed, so memory
    is bounded by ``(prefetch_batches + 1) * batch_size`` files however long the listing is;
  * batches come out in listing order with every image in its listing position: the sequential
    first-seen / last-seen semantics of the reference's dedupe (``webdav_sync.py:324-398``) only
    survive batching if order does.

Error convention = the reference's: an invalid extension / MIME type (``:314``) or a failed
download (``:320``, any exception) yields ``data=None`` for that image and the batch goes on;
a file that does not decode yields ``rgb=None`` and ``metadata={}`` (``:101-103``).
Pure host code: no tensor library, no device call; the tests run without a GPU.
"""
from __future__ import annotations

import io
import logging
import threading
from collections import deque
from concurrent.futures import Future, ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import Callable, Deque, Dict, Iterator, List, Optional, Sequence

import numpy as np

logger = logging.getLogger(__name__)


@dataclass
class FeederBatch:
    infos: List[Dict]                                   # the listing entries, in listing order
    datas: List[Optional[bytes]]                        # file bytes; None = skipped (invalid or failed download)
    metadata: List[Dict] = field(default_factory=list)  # {'width','height','format','mode'} or {} (webdav_sync.py:83-103)
    rgb: List[Optional[np.ndarray]] = field(default_factory=list)   # decoded HWC uint8 RGB, None when absent


def image_metadata(data: bytes) -> Dict:
    """The reference's ``_get_image_metadata`` (header parse only, ``{}`` on any error)."""
    try:
        from PIL import Image as PILImage

        img = PILImage.open(io.BytesIO(data))
        return {"width": img.width, "height": img.height, "format": img.format, "mode": img.mode}
    except Exception as e:  # noqa: BLE001 - the reference swallows everything here
        logger.warning("metadata extraction failed: %s", e)
        return {}


def decode_rgb(data: bytes) -> Optional[np.ndarray]:
    """File bytes -> contiguous HWC uint8 RGB (what ``b2_thumbnails_host`` takes), None if undecodable."""
    try:
        from PIL import Image as PILImage

        with PILImage.open(io.BytesIO(data)) as img:
            return np.ascontiguousarray(np.asarray(img.convert("RGB"), dtype=np.uint8))
    except Exception as e:  # noqa: BLE001
        logger.warning("decode failed: %s", e)
        return None


class DownloadDecodeFeeder:
    """``for batch in feeder.batches(listing): ...`` — see the module docstring.

    ``fetch(info) -> bytes | None`` downloads one file and must not raise (``WebDAVSync._fetch``);
    ``validate(info) -> bool`` is the extension / MIME filter (``WebDAVSync._validate_image``).
    """

    def __init__(self, fetch: Callable[[Dict], Optional[bytes]], validate: Callable[[Dict], bool] = lambda info: True,
                 batch_size: int = 50, download_workers: int = 8, decode_workers: int = 4,
                 prefetch_batches: int = 1, decode: bool = False, want_metadata: bool = True):
        if batch_size < 1 or download_workers < 1 or decode_workers < 1 or prefetch_batches < 0:
            raise ValueError("batch_size, download_workers, decode_workers >= 1 and prefetch_batches >= 0")
        self.fetch, self.validate = fetch, validate
        self.batch_size, self.prefetch_batches = batch_size, prefetch_batches
        self.download_workers, self.decode_workers = download_workers, decode_workers
        self.decode, self.want_metadata = decode, want_metadata
        self._lock = threading.Lock()
        self.peak_inflight_files = 0                    # high-water mark of files held by the feeder (for the tests)
        self._inflight = 0

    # one image: download, then header parse (+ decode) on the decode pool
    def _one(self, info: Dict, decode_pool: ThreadPoolExecutor):
        with self._lock:
            self._inflight += 1
            self.peak_inflight_files = max(self.peak_inflight_files, self._inflight)
        if not self.validate(info):
            return None, {}, None
        try:
            data = self.fetch(info)
        except Exception as e:  # noqa: BLE001 - a fetch that raises is a failed download, not a failed batch
            logger.warning("download failed for %s: %s", info.get("name", "unknown"), e)
            data = None
        if data is None:
            return None, {}, None
        meta_f: Optional[Future] = decode_pool.submit(image_metadata, data) if self.want_metadata else None
        rgb_f: Optional[Future] = decode_pool.submit(decode_rgb, data) if self.decode else None
        return data, (meta_f.result() if meta_f else {}), (rgb_f.result() if rgb_f else None)

    def batches(self, images: Sequence[Dict]) -> Iterator[FeederBatch]:
        groups = [list(images[i:i + self.batch_size]) for i in range(0, len(images), self.batch_size)]
        if not groups:
            return
        with ThreadPoolExecutor(self.download_workers, thread_name_prefix="b2-feeder-dl") as dl, \
                ThreadPoolExecutor(self.decode_workers, thread_name_prefix="b2-feeder-dec") as dec:
            pending: Deque[List[Future]] = deque()
            nxt = 0

            def top_up(limit: int):
                nonlocal nxt
                while nxt < len(groups) and len(pending) < limit:
                    pending.append([dl.submit(self._one, info, dec) for info in groups[nxt]])
                    nxt += 1

            top_up(self.prefetch_batches + 1)                    # the first batch plus the prefetch window
            done = 0
            while pending:
                futures = pending.popleft()
                results = [f.result() for f in futures]          # listing order, whatever order they finished in
                top_up(self.prefetch_batches)                    # later batches download while this one is consumed
                yield FeederBatch(infos=groups[done], datas=[r[0] for r in results],
                                  metadata=[r[1] for r in results], rgb=[r[2] for r in results])
                with self._lock:
                    self._inflight -= len(futures)
                if not pending:
                    top_up(1)                                    # prefetch_batches == 0: strictly one batch at a time
                done += 1
