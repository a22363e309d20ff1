"""The persistence seam either side of the hot path.

The reference talks to PostgreSQL through SQLAlchemy (``db.query(Imagem).filter_by(
content_hash=...).first()``, ``db.add``, ``db.flush``, ``db.commit`` — webdav_sync.py:324-398).
The storage engine is out of scope; the drop-in only needs the four
operations below on table ``imagens`` (primary key ``content_hash``, app/db/models.py:202-222)
and one on ``conjuntos_imagens``.  ``DictImageStore`` is the in-memory implementation used by
the tests and the benchmark; INTEGRATION.md shows the SQLAlchemy adapter a maintainer would add.
Rows are plain dicts with the ``Imagem`` column names.
"""
from __future__ import annotations

import threading
import uuid
from typing import Dict, Iterable, List, Optional, Protocol


class DuplicateKeyError(KeyError):
    """``insert`` met a row with the same primary key: a concurrent writer got there first — what SQLAlchemy
    reports as ``IntegrityError`` on flush (webdav_sync.py:355, activity_api_sync.py:875).  The caller rolls back
    and takes the merge branch."""


class ImageStore(Protocol):
    def get(self, content_hash: str) -> Optional[Dict]: ...
    def get_many(self, content_hashes: Iterable[str]) -> Dict[str, Dict]: ...
    def insert(self, row: Dict) -> None: ...
    def update(self, content_hash: str, fields: Dict) -> None: ...
    def commit(self) -> None: ...
    def rollback(self) -> None: ...
    def folder_for(self, file_id: str, name: str, path: str, now) -> Optional[Dict]: ...
    # optional (SURVEY 8(f) rank 2 iii): thumbnails keyed by content hash, written in arrival order
    # def put_thumbnail(self, content_hash: str, thumb_u8_hwc) -> str: ...   returns the reference stored in metadados


class DictImageStore:
    """``imagens`` as a dict keyed by content_hash; ``conjuntos_imagens`` keyed by file_id."""

    def __init__(self, rows: Optional[Dict[str, Dict]] = None):
        self.rows: Dict[str, Dict] = rows if rows is not None else {}
        self.folders: Dict[str, Dict] = {}
        self.thumbs: Dict[str, "object"] = {}                 # side table keyed by content_hash
        self.commits = 0
        self._lock = threading.Lock()

    def get(self, content_hash: str) -> Optional[Dict]:
        return self.rows.get(content_hash)

    def get_many(self, content_hashes: Iterable[str]) -> Dict[str, Dict]:
        """One ``WHERE content_hash IN (...)`` instead of one lookup per image."""
        return {h: self.rows[h] for h in set(content_hashes) if h in self.rows}

    def insert(self, row: Dict) -> None:
        with self._lock:
            if row["content_hash"] in self.rows:
                raise DuplicateKeyError(f"duplicate primary key {row['content_hash']}")
            self.rows[row["content_hash"]] = row

    def update(self, content_hash: str, fields: Dict) -> None:
        self.rows[content_hash].update(fields)

    def commit(self) -> None:
        self.commits += 1

    def rollback(self) -> None:
        pass

    def folder_for(self, file_id: str, name: str, path: str, now) -> Optional[Dict]:
        """Lookup-or-create of the ConjuntoImagens row by NextCloud file_id
        (activity_api_sync.py:818-841)."""
        with self._lock:
            f = self.folders.get(file_id)
            if f is None:
                f = self.folders[file_id] = {
                    "id_cnj": uuid.uuid4(), "nome_conj": name, "caminho_conj": path, "file_id": file_id,
                    "imagens_sincronizadas": False, "existe_no_nextcloud": True, "data_proc": now, "data_sinc": now,
                }
            return f

    def put_thumbnail(self, content_hash: str, thumb) -> str:
        """Side table ``thumbnails(content_hash PRIMARY KEY, h, w, pixels)``: last write wins, like nome_img."""
        import numpy as np
        self.thumbs[content_hash] = np.array(thumb, copy=True)
        return f"thumbnails/{content_hash}"

    def sorted_hashes(self) -> List[str]:
        return sorted(self.rows)
