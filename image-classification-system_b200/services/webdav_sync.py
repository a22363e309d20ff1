"""Drop-in for the ingest hot path of ``app/services/webdav_sync.py`` (reference).

Same class name, method names, argument meaning, return types and error behaviour as the
reference's ``WebDAVSync`` for the methods ON the path:

  _calculate_hash_from_bytes(data) -> str                  webdav_sync.py:49-59
  _validate_image(file_info) -> bool                       webdav_sync.py:61-81
  _get_image_metadata(image_data) -> Dict                  webdav_sync.py:83-103  (host, Pillow header)
  _download_and_process_image(image_info) -> (hash, meta)  webdav_sync.py:428-465
  _process_image_batch(images, folder_path, conjunto_id)   webdav_sync.py:296-426
  sync_images_in_folder(folder_path, conjunto_id)          webdav_sync.py:247-294 (batch loop only)

What changes is the shape of the work: the reference hashes, looks up and inserts one image at
a time; here a batch is downloaded first, hashed in ONE device call (libb2ingest
``b2_sha256_batch``), resolved against the table with ONE ``IN`` lookup plus the device dedupe
kernel (``b2_dedupe``), and then applied in arrival order.  Results — hashes, created/updated
decisions, the stats dict and the final rows — are identical to the sequential reference
(tests/test_services_gpu.py replays tests/golden/reference_ingest.json).

``db`` is an :class:`..store.ImageStore` (the storage engine is out of scope; INTEGRATION.md
shows the SQLAlchemy adapter).  Listing folders, marking removed images and folder bookkeeping
(:105-245, :467-532) stay in the reference.
"""
from __future__ import annotations

import logging
from datetime import datetime, timezone
from typing import Callable, Dict, List, Optional, Tuple

from .. import hostapi
from ..feeder import DownloadDecodeFeeder, FeederBatch, image_metadata
from ..ingest import hash_and_dedupe
from ..store import DuplicateKeyError

logger = logging.getLogger(__name__)

NEXTCLOUD_SYNC_BATCH_SIZE = 50        # app/core/config.py:56


def _utc_now() -> datetime:
    return datetime.now(timezone.utc)


class WebDAVSync:
    """Batch-on-device version of the reference's full-scan ingest."""

    ALLOWED_MIME_TYPES = [
        "image/jpeg", "image/jpg", "image/png", "image/gif", "image/bmp", "image/tiff", "image/webp",
    ]
    ALLOWED_EXTENSIONS = [".jpg", ".jpeg", ".png", ".gif", ".bmp", ".tiff", ".webp"]
    SYNC_METHOD = "webdav"

    def __init__(self, nextcloud_client, db, now: Callable[[], datetime] = _utc_now,
                 batch_size: int = NEXTCLOUD_SYNC_BATCH_SIZE, device: Optional[int] = None,
                 download_workers: int = 1, prefetch_batches: int = 1, store_thumbnails: bool = False):
        """``download_workers`` > 1: ``sync_images_in_folder`` downloads through the feeder (``feeder.py``) —
        that many GETs in flight and the next ``prefetch_batches`` batches downloading while the current one is
        hashed and written; 1 = one GET at a time, as the reference does.  Same results either way."""
        self.client = nextcloud_client
        self.db = db
        self._now = now
        self.batch_size = batch_size
        self.device = device
        self.download_workers = download_workers
        self.prefetch_batches = prefetch_batches
        # thumbnails of newly inserted images written back through the store (needs the feeder's decoded pixels)
        self.store_thumbnails = store_thumbnails

    # ------------------------------------------------------------------ single-item API
    def _calculate_hash_from_bytes(self, data: bytes) -> str:
        """SHA-256 of the file bytes, lowercase hex (64 chars): a 1-element device batch."""
        return hostapi.hash_batch([data], self.device)[0]

    def _validate_image(self, file_info: Dict) -> bool:
        name = file_info.get("name", "").lower()
        if not any(name.endswith(ext) for ext in self.ALLOWED_EXTENSIONS):
            return False
        content_type = file_info.get("content_type", "").lower()
        return any(mime in content_type for mime in self.ALLOWED_MIME_TYPES)

    def _get_image_metadata(self, image_data: bytes) -> Dict:
        """Header parse only, in the reference's own host library (Pillow); ``{}`` on any error."""
        return image_metadata(image_data)

    def _fetch(self, image_info: Dict) -> Optional[bytes]:
        """Download into memory; ``None`` on any failure (never aborts the batch)."""
        try:
            return self.client.get_file(image_info.get("path", "")).content
        except Exception as e:  # noqa: BLE001 - connection error, timeout, anything: skip the image
            logger.warning("download failed for %s: %s", image_info.get("name", "unknown"), e)
            return None

    def _download_and_process_image(self, image_info: Dict) -> Tuple[Optional[str], Dict]:
        data = self._fetch(image_info)
        if data is None:
            return None, {}
        try:
            return self._calculate_hash_from_bytes(data), self._get_image_metadata(data)
        except Exception as e:  # noqa: BLE001
            logger.debug("processing failed for %s: %s", image_info.get("name", "unknown"), e)
            return None, {}

    # ------------------------------------------------------------------ batch API
    @staticmethod
    def _nextcloud_meta(info: Dict, full: bool) -> Dict:
        lm = info.get("last_modified")
        meta = {
            "file_id": info.get("file_id", ""),
            "etag": info.get("etag", ""),
            "last_modified": lm.isoformat() if lm else None,
        }
        if full:
            meta["content_type"] = info.get("content_type", "")
            meta["size"] = info.get("content_length", 0)
        return meta

    def _process_image_batch(self, images: List[Dict], folder_path: str, conjunto_id,
                             prefetched: Optional[FeederBatch] = None) -> Dict[str, int]:
        """``prefetched``: the same images already downloaded (and header-parsed) by the feeder."""
        now = self._now()
        if prefetched is not None:
            datas: List[Optional[bytes]] = prefetched.datas
        else:
            datas = [self._fetch(info) if self._validate_image(info) else None for info in images]
        try:
            decision = hash_and_dedupe(datas, existing_hashes=lambda hs: self.db.get_many(hs).keys(),
                                       device=self.device)
        except Exception as e:  # noqa: BLE001 - device failure: nothing of the batch is processed
            logger.error("device ingest failed for batch in %s: %s", folder_path, e)
            return {"processed": 0, "created": 0, "updated": 0}

        thumbs = self._thumbnails(prefetched, decision) if self.store_thumbnails else None
        metadata = prefetched.metadata if prefetched is not None else None
        return self._apply_batch(images, decision.hashes, decision.is_new, datas, metadata, thumbs, now, conjunto_id)

    def _apply_batch(self, images: List[Dict], hashes, is_new, datas, metadata, thumbs, now, conjunto_id) -> Dict[str, int]:
        """The apply step of ``_process_image_batch`` (webdav_sync.py:324-424), shared by the batch form and the
        streaming form.  ``metadata``: header dicts parsed by the feeder (else parsed here from ``datas``)."""
        # Apply in arrival order.  The device decided created / updated for a table nobody else writes; a concurrent
        # session (the Activity-API sync runs beside this one) can still insert the same hash first or
        # remove a row, so every image keeps the reference's own safety net: duplicate key on insert -> rollback,
        # re-read, minimal merge, counted as updated (:355-369); any other error -> log, rollback, next image
        # (:421-424).  The stats are therefore counted here; without interference they equal ``decision.stats``.
        stats = {"processed": 0, "created": 0, "updated": 0}
        for i, info in enumerate(images):
            content_hash = hashes[i]
            if not content_hash:
                continue
            try:
                row = None if is_new[i] else self.db.get(content_hash)
                if row is None:                               # the insert branch (also: the row vanished meanwhile)
                    nc = self._nextcloud_meta(info, full=True)
                    image_md = metadata[i] if metadata is not None else self._get_image_metadata(datas[i])
                    if thumbs is not None and thumbs[i] is not None:
                        image_md = dict(image_md, thumb=self.db.put_thumbnail(content_hash, thumbs[i]))
                    try:
                        self.db.insert({
                            "content_hash": content_hash,
                            "nome_img": info.get("name", ""),
                            "caminho_img": info.get("path", ""),
                            "metadados": {
                                "nextcloud": {"file_id": nc["file_id"], "etag": nc["etag"],
                                              "content_type": nc["content_type"], "size": nc["size"],
                                              "last_modified": nc["last_modified"]},
                                "image": image_md,
                                "sync": {"sync_method": self.SYNC_METHOD, "sync_timestamp": now.isoformat()},
                            },
                            "existe_no_nextcloud": True,
                            "data_proc": now,
                            "data_sinc": now,
                            "id_cnj": conjunto_id,
                        })
                        stats["created"] += 1
                    except DuplicateKeyError:
                        self.db.rollback()
                        if self.db.get(content_hash) is None:
                            logger.debug("hash %s not found after duplicate-key error", content_hash[:16])
                            continue
                        self.db.update(content_hash, {"nome_img": info.get("name", ""), "caminho_img": info.get("path", ""),
                                                      "existe_no_nextcloud": True, "data_sinc": now})
                        stats["updated"] += 1
                else:
                    md = row.get("metadados")
                    if md:
                        if "nextcloud" in md:
                            md["nextcloud"].update(self._nextcloud_meta(info, full=False))
                        else:
                            md["nextcloud"] = self._nextcloud_meta(info, full=True)
                        md["sync"] = {"sync_method": self.SYNC_METHOD, "sync_timestamp": now.isoformat()}
                    self.db.update(content_hash, {
                        "nome_img": info.get("name", ""),
                        "caminho_img": info.get("path", ""),
                        "existe_no_nextcloud": True,
                        "data_sinc": now,
                        "metadados": md,
                    })
                    stats["updated"] += 1
                stats["processed"] += 1
            except Exception as e:  # noqa: BLE001 - one image never aborts the batch
                logger.debug("failed to apply %s: %s", info.get("name", "unknown"), e)
                self.db.rollback()
                continue
        return stats

    def _thumbnails(self, prefetched: Optional[FeederBatch], decision):
        """256x256 thumbnails of the images about to be INSERTED (decoded pixels come from
        the feeder), one device call for the batch; written back through ``ImageStore.put_thumbnail`` and referenced
        from ``metadados['image']['thumb']`` in arrival order.  ``None`` entries: no decoded pixels."""
        n = len(decision.hashes)
        out = [None] * n
        if prefetched is None or not hasattr(self.db, "put_thumbnail"):
            return out
        idx = [i for i in range(n) if decision.hashes[i] and decision.is_new[i] and i < len(prefetched.rgb) and prefetched.rgb[i] is not None]
        if idx:
            t, _ = hostapi.thumbnails([prefetched.rgb[i] for i in idx], 256, 256, want_preview=False, device=self.device)
            for j, i in enumerate(idx):
                out[i] = t[j]
        return out

    def sync_images_in_folder_streaming(self, folder_path: str, conjunto_id, ring, listings_in_flight: int = 3) -> Dict[str, int]:
        """The same loop for a bulk (re-)sync, as a stream (SURVEY 8(f) rank 1 + 4): the feeder downloads and decodes
        batch i+2, the ingest ring (``pipeline.IngestRing``) hashes and resizes batch i+1 — file bytes are the message,
        decoded pixels make the thumbnail, failed downloads are skipped — while batch i is applied to the table.  A
        batch is a LISTING of the ring: its digests, the first-occurrence flags inside the batch and the thumbnails come
        back in listing order; created / updated is then decided against the table as it stands when the batch is
        applied (ONE ``IN`` lookup), batches strictly in listing order — the results are those of the sequential
        reference loop.  Thumbnails are written back when ``store_thumbnails`` is set."""
        import numpy as np
        stats = {"images_processed": 0, "images_created": 0, "images_updated": 0, "images_marked_removed": 0}
        pending: List = []

        def finish(entry):
            ticket, fb = entry
            res = ring.result(ticket)
            valid = [d is not None for d in fb.datas]
            hashes = [bytes(res.digests[i]).hex() if valid[i] else None for i in range(len(valid))]
            stored = self.db.get_many([h for h in hashes if h])
            is_new = [bool(res.is_new[i]) and hashes[i] not in stored for i in range(len(valid))]
            thumbs = None
            if self.store_thumbnails and res.thumbs is not None and hasattr(self.db, "put_thumbnail"):
                thumbs = [np.array(res.thumbs[i], copy=True) if (is_new[i] and fb.rgb[i] is not None) else None
                          for i in range(len(valid))]
            b = self._apply_batch(fb.infos, hashes, is_new, fb.datas, fb.metadata, thumbs, self._now(), conjunto_id)
            stats["images_processed"] += b["processed"]
            stats["images_created"] += b["created"]
            stats["images_updated"] += b["updated"]
            self.db.commit()

        try:
            items = self.client.list_folder(folder_path, depth=1)
            images = self.client.filter_images(items)
            feeder = DownloadDecodeFeeder(self._fetch, self._validate_image, self.batch_size, max(self.download_workers, 1),
                                          prefetch_batches=max(self.prefetch_batches, 1), decode=True)
            for fb in feeder.batches(images):
                if len(pending) >= listings_in_flight:
                    finish(pending.pop(0))
                files = [np.frombuffer(d, dtype=np.uint8) if d else None for d in fb.datas]
                valid = np.array([d is not None for d in fb.datas], dtype=np.uint8)
                if not valid.any():
                    continue                                  # nothing downloadable in this batch
                pending.append((ring.submit(fb.rgb, files=files, valid=valid), fb))
            while pending:
                finish(pending.pop(0))
        except Exception:
            self.db.rollback()
            raise
        return stats

    def sync_images_in_folder(self, folder_path: str, conjunto_id) -> Dict[str, int]:
        """The batch loop of the reference (:273-283).  Marking removed images (:286) is
        bookkeeping outside the hot path: ``images_marked_removed`` stays 0 here."""
        stats = {"images_processed": 0, "images_created": 0, "images_updated": 0, "images_marked_removed": 0}
        try:
            items = self.client.list_folder(folder_path, depth=1)
            images = self.client.filter_images(items)
            if self.download_workers > 1 or self.store_thumbnails:
                feeder = DownloadDecodeFeeder(self._fetch, self._validate_image, self.batch_size,
                                              self.download_workers, prefetch_batches=self.prefetch_batches,
                                              decode=self.store_thumbnails)
                batches = ((fb.infos, fb) for fb in feeder.batches(images))
            else:
                batches = ((images[i:i + self.batch_size], None) for i in range(0, len(images), self.batch_size))
            for infos, fb in batches:
                b = self._process_image_batch(infos, folder_path, conjunto_id, prefetched=fb)
                stats["images_processed"] += b["processed"]
                stats["images_created"] += b["created"]
                stats["images_updated"] += b["updated"]
                self.db.commit()
        except Exception:
            self.db.rollback()
            raise
        return stats
