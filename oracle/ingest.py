"""Oracle: ingest batch (validate -> hash -> metadata -> dedupe on content_hash).
TEST INFRASTRUCTURE — see oracle/__init__.py.

Restates, with a ``dict`` standing in for the ``imagens`` table (primary key
``content_hash``, app/db/models.py:212):
  app/services/webdav_sync.py:61-81    _validate_image
  app/services/webdav_sync.py:83-103   _get_image_metadata
  app/services/webdav_sync.py:296-426  _process_image_batch
  app/api/routes/images.py:47-94       buscar_imagens_por_hash
Pinned on tests/golden/reference_ingest.json (the reference's own functions run against
a stub session; tests/golden/make_reference_golden.py).
"""
from __future__ import annotations

import io
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

from .hashing import sha256_hex

# webdav_sync.py:29-36
ALLOWED_MIME_TYPES = [
    "image/jpeg", "image/jpg", "image/png", "image/gif",
    "image/bmp", "image/tiff", "image/webp",
]
ALLOWED_EXTENSIONS = [".jpg", ".jpeg", ".png", ".gif", ".bmp", ".tiff", ".webp"]


def validate_image(file_info: Dict) -> bool:
    """webdav_sync.py:61-81 — extension AND content-type substring must both match."""
    name = file_info.get("name", "").lower()
    if not any(name.endswith(ext) for ext in ALLOWED_EXTENSIONS):
        return False
    content_type = file_info.get("content_type", "").lower()
    if not any(mime in content_type for mime in ALLOWED_MIME_TYPES):
        return False
    return True


def image_metadata(image_data: bytes) -> Dict:
    """webdav_sync.py:83-103 — PIL header parse only (lazy open, no pixel decode);
    ``{}`` on any exception (e.g. headerless raw RGB -> UnidentifiedImageError)."""
    try:
        from PIL import Image as PILImage

        img = PILImage.open(io.BytesIO(image_data))
        return {"width": img.width, "height": img.height, "format": img.format, "mode": img.mode}
    except Exception:
        return {}


def dedupe_batch(
    hashes: Sequence[Optional[str]],
    existing: Optional[Iterable[str]] = None,
) -> Tuple[List[bool], List[int], Dict[str, int]]:
    """Sequential first-occurrence-wins resolution (webdav_sync.py:311-400).

    ``hashes[i] is None`` models an image that was skipped before the lookup (invalid
    extension/MIME at :314, or failed download at :320): not counted anywhere.
    Returns ``(is_new, first_index, stats)``: ``is_new[i]`` — the insert branch (:326-354)
    ran; ``first_index[i]`` — index in this batch of the first occurrence of the same hash
    (``-1`` for skipped entries; for a hash already in ``existing`` it is still the first
    in-batch index); ``stats`` — ``{'processed','created','updated'}`` (:308, :354, :398, :400).
    """
    table = set(existing) if existing is not None else set()
    first: Dict[str, int] = {}
    is_new: List[bool] = []
    first_index: List[int] = []
    stats = {"processed": 0, "created": 0, "updated": 0}
    for i, h in enumerate(hashes):
        if not h:                      # :320  `if not content_hash: continue`
            is_new.append(False)
            first_index.append(-1)
            continue
        first.setdefault(h, i)
        first_index.append(first[h])
        if h not in table:             # :324 lookup misses -> insert + flush (:352-354)
            table.add(h)
            is_new.append(True)
            stats["created"] += 1
        else:                          # :371-398 update branch
            is_new.append(False)
            stats["updated"] += 1
        stats["processed"] += 1        # :400
    return is_new, first_index, stats


def process_image_batch(
    images: List[Dict],
    fetch: Callable[[str], bytes],
    table: Dict[str, Dict],
    conjunto_id=None,
    now_iso: str = "1970-01-01T00:00:00+00:00",
    sync_method: str = "webdav",
) -> Dict[str, int]:
    """Full restatement of ``_process_image_batch`` (webdav_sync.py:296-426) for a single
    session (the IntegrityError branches :355-369 / :402-420 need a concurrent writer and
    never fire here).  ``table`` maps content_hash -> row dict with the ``Imagem`` columns
    (models.py:202-222); it is mutated in place exactly as the ORM objects would be:
    identity/``data_proc``/``id_cnj``/``metadados.image`` come from the FIRST occurrence,
    ``nome_img``/``caminho_img``/``data_sinc`` from the LAST.
    ``fetch(path)`` stands in for ``client.get_file(path).content`` (:441-442); raising
    means a failed download -> ``(None, {})`` -> skipped (:455-465, :320).
    """
    stats = {"processed": 0, "created": 0, "updated": 0}
    for info in images:
        if not validate_image(info):
            continue
        try:
            data = fetch(info.get("path", ""))
            content_hash, metadata = sha256_hex(data), image_metadata(data)
        except Exception:
            content_hash, metadata = None, {}
        if not content_hash:
            continue
        lm = info.get("last_modified")
        row = table.get(content_hash)
        if row is None:
            table[content_hash] = {
                "content_hash": content_hash,
                "nome_img": info.get("name", ""),
                "caminho_img": info.get("path", ""),
                "metadados": {
                    "nextcloud": {
                        "file_id": info.get("file_id", ""),
                        "etag": info.get("etag", ""),
                        "content_type": info.get("content_type", ""),
                        "size": info.get("content_length", 0),
                        "last_modified": lm.isoformat() if lm else None,
                    },
                    "image": metadata,
                    "sync": {"sync_method": sync_method, "sync_timestamp": now_iso},
                },
                "existe_no_nextcloud": True,
                "data_proc": now_iso,
                "data_sinc": now_iso,
                "id_cnj": conjunto_id,
            }
            stats["created"] += 1
        else:
            row["nome_img"] = info.get("name", "")
            row["caminho_img"] = info.get("path", "")
            row["existe_no_nextcloud"] = True
            row["data_sinc"] = now_iso
            md = row.get("metadados")
            if md:
                if "nextcloud" in md:
                    md["nextcloud"].update({
                        "file_id": info.get("file_id", ""),
                        "etag": info.get("etag", ""),
                        "last_modified": lm.isoformat() if lm else None,
                    })
                else:
                    md["nextcloud"] = {
                        "file_id": info.get("file_id", ""),
                        "etag": info.get("etag", ""),
                        "content_type": info.get("content_type", ""),
                        "size": info.get("content_length", 0),
                        "last_modified": lm.isoformat() if lm else None,
                    }
                md["sync"] = {"sync_method": sync_method, "sync_timestamp": now_iso}
            stats["updated"] += 1
        stats["processed"] += 1
    return stats


def buscar_por_hash(
    uploads: Sequence[Tuple[Optional[str], bytes]],
    table: Dict[str, Dict],
) -> Dict:
    """routes/images.py:47-94.  ``uploads`` = ``(content_type, data)`` per file.
    Non-``image/*`` uploads yield ``hash=""`` and are never hashed (:49-56)."""
    resultados = []
    found = 0
    for content_type, data in uploads:
        if not content_type or not content_type.startswith("image/"):
            resultados.append({"hash": "", "encontrada": False, "imagem": None})
            continue
        h = sha256_hex(data)
        row = table.get(h)
        if row is not None:
            found += 1
            resultados.append({
                "hash": h,
                "encontrada": True,
                "imagem": {
                    "content_hash": row["content_hash"],
                    "nome_img": row["nome_img"],
                    "caminho_img": row["caminho_img"],
                },
            })
        else:
            resultados.append({"hash": h, "encontrada": False, "imagem": None})
    return {"total_enviadas": len(uploads), "total_encontradas": found, "resultados": resultados}
