"""Drop-in for ``buscar_imagens_por_hash`` (app/api/routes/images.py:18-101): hash every
uploaded file and look it up by primary key.  The reference hashes and queries per file; here
all ``image/*`` uploads are hashed in ONE device call and resolved with ONE ``IN`` lookup.  The
response has the shape of ``RespostaBuscaImagens`` (app/schemas/image_schema.py:8-29) as plain
dicts; FastAPI/Pydantic wiring stays in the reference.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

from ... import hostapi


class NoFilesError(ValueError):
    """The reference answers HTTP 400 when no file is sent (images.py:37-41)."""


def buscar_imagens_por_hash(files: Sequence[Tuple[Optional[str], bytes]], db, device: Optional[int] = None) -> Dict:
    """``files``: ``(content_type, data)`` per upload (what ``UploadFile.content_type`` and
    ``await file.read()`` give).  Non-``image/*`` uploads yield ``hash=""`` and are not hashed."""
    if not files:
        raise NoFilesError("Nenhuma imagem foi enviada. Envie pelo menos uma imagem.")
    is_image = [bool(ct) and ct.startswith("image/") for ct, _ in files]
    hashes = hostapi.hash_batch([data for (ct, data), ok in zip(files, is_image) if ok], device)
    found = db.get_many(hashes)
    resultados: List[Dict] = []
    total = 0
    it = iter(hashes)
    for ok in is_image:
        if not ok:
            resultados.append({"hash": "", "encontrada": False, "imagem": None})
            continue
        h = next(it)
        row = found.get(h)
        if row is not None:
            total += 1
            resultados.append({"hash": h, "encontrada": True, "imagem": {
                "content_hash": row["content_hash"], "nome_img": row["nome_img"], "caminho_img": row["caminho_img"]}})
        else:
            resultados.append({"hash": h, "encontrada": False, "imagem": None})
    return {"total_enviadas": len(files), "total_encontradas": total, "resultados": resultados}
