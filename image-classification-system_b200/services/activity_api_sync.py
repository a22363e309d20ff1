"""Drop-in for ``ActivityAPISync._process_new_image`` (app/services/activity_api_sync.py:779-900):
the single-image variant of the ingest path.  Same name, argument and ``bool`` result; the
hash is a 1-element device batch and the dedupe decision is the primary-key lookup.  Event
fetching / folder events (:54-778) stay in the reference.
"""
from __future__ import annotations

import logging
from datetime import datetime
from typing import Callable, Dict, Optional

from .. import hostapi
from ..store import DuplicateKeyError
from .webdav_sync import WebDAVSync, _utc_now

logger = logging.getLogger(__name__)


class ActivityAPISync:
    ALLOWED_MIME_TYPES = WebDAVSync.ALLOWED_MIME_TYPES
    ALLOWED_EXTENSIONS = WebDAVSync.ALLOWED_EXTENSIONS

    def __init__(self, nextcloud_client, db, now: Callable[[], datetime] = _utc_now, device: Optional[int] = None):
        self.client = nextcloud_client
        self.db = db
        self._now = now
        self.device = device

    _validate_image = WebDAVSync._validate_image
    _get_image_metadata = WebDAVSync._get_image_metadata

    def _process_new_image(self, image_info: Dict) -> bool:
        try:
            if not self._validate_image(image_info):
                return False
            image_data = self.client.get_file(image_info.get("path", "")).content
            content_hash = hostapi.hash_batch([image_data], self.device)[0]
            row = self.db.get(content_hash)
            now = self._now()

            image_path = image_info.get("path", "")
            folder_path = image_path.rsplit("/", 1)[0] if "/" in image_path else ""
            folder_items = self.client.list_folder(folder_path, depth=0)
            folder_info = next((it for it in folder_items if it.get("is_collection", False)), None)
            if not folder_info:
                logger.warning("folder not found for %s", image_path)
                return False
            conjunto = self.db.folder_for(folder_info.get("file_id", ""), folder_info.get("name", ""), folder_path, now)
            if not conjunto:
                return False

            minimal = {"nome_img": image_info.get("name", ""), "caminho_img": image_info.get("path", ""),
                       "existe_no_nextcloud": True, "data_sinc": now}
            if row is None:
                lm = image_info.get("last_modified")
                new_row = {
                    "content_hash": content_hash,
                    "nome_img": image_info.get("name", ""),
                    "caminho_img": image_info.get("path", ""),
                    "metadados": {
                        "nextcloud": {
                            "file_id": image_info.get("file_id", ""),
                            "etag": image_info.get("etag", ""),
                            "content_type": image_info.get("content_type", ""),
                            "size": image_info.get("content_length", 0),
                            "last_modified": lm.isoformat() if lm else None,
                        },
                        "image": self._get_image_metadata(image_data),     # only when new (:843-845)
                        "sync": {"sync_method": "activity_api", "sync_timestamp": now.isoformat()},
                    },
                    "existe_no_nextcloud": True,
                    "data_proc": now,
                    "data_sinc": now,
                    "id_cnj": conjunto["id_cnj"],
                }
                try:
                    self.db.insert(new_row)
                    self.db.commit()
                    return True
                except DuplicateKeyError:
                    # another session inserted the same content first: merge into its row (:875-887)
                    self.db.rollback()
                    if self.db.get(content_hash) is None:
                        return False
                    self.db.update(content_hash, minimal)
                    self.db.commit()
                    return True
            self.db.update(content_hash, minimal)
            self.db.commit()
            return True
        except Exception as e:  # noqa: BLE001 - same contract as the reference: log, rollback, False
            logger.error("failed to process new image: %s", e)
            self.db.rollback()
            return False
