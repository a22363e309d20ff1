"""Label-aggregation functions of ``app/crud/classificacao_crud.py`` and
``app/api/routes/classificacoes.py`` that sit on the hot path, with the reference's names.

  obter_classificacoes_imagens(db, id_con, imagens) -> dict      classificacao_crud.py:284-324
  obter_contagem_classificacoes(db, id_con) -> {"total": int}    routes/classificacoes.py:204-236

Per request these touch <= 20 images / <= 100 rows, so the per-user grouping is host
bookkeeping of row objects (no arithmetic to accelerate); what runs on the device is the BULK
form over the whole table — ``contagem_classificacoes_todos`` (every annotator's distinct-image
count in one pass, ``b2_distinct_images_per_annotator``) and the cross-annotator tally + Fleiss
kappa in ``labels.py`` — which is what BASELINE.json's label configs measure.

``db`` exposes ``classificacoes``: an iterable of row dicts with keys id_cla, id_con, id_img,
id_opc, ativo in storage order (the storage engine is out of scope, see store.py).
"""
from __future__ import annotations

import uuid
from typing import Dict, List, Optional, Sequence

import numpy as np

from .. import labels


def _as_uuid(id_con):
    try:
        return uuid.UUID(id_con) if isinstance(id_con, str) else id_con
    except (ValueError, TypeError, AttributeError):
        return None


def _content_hash(img) -> str:
    return img["content_hash"] if isinstance(img, dict) else img.content_hash


def obter_classificacoes_imagens(db, id_con, imagens: Sequence) -> dict:
    """{content_hash: [row, ...]} of the user's ACTIVE rows for the given images, row order =
    storage order; ``{}`` for a malformed user id or an empty image list."""
    key = _as_uuid(id_con)
    if key is None or not imagens:
        return {}
    wanted = {_content_hash(i) for i in imagens}
    out: Dict[str, List] = {}
    for c in db.classificacoes:
        if _as_uuid(c["id_con"]) == key and c["ativo"] is True and c["id_img"] in wanted:
            out.setdefault(c["id_img"], []).append(c)
    return out


def obter_contagem_classificacoes(db, id_con) -> Dict[str, int]:
    """COUNT(DISTINCT id_img) WHERE id_con = ? AND ativo; ``{"total": 0}`` on a bad id."""
    key = _as_uuid(id_con)
    if key is None:
        return {"total": 0}
    return {"total": len({c["id_img"] for c in db.classificacoes
                          if _as_uuid(c["id_con"]) == key and c["ativo"] is True})}


def contagem_classificacoes_todos(db, device: Optional[int] = None) -> Dict[str, int]:
    """Device bulk form of :func:`obter_contagem_classificacoes` for EVERY annotator at once
    (audits the incremental counter of classificacao_crud.py:471-475)."""
    rows = list(db.classificacoes)
    users = sorted({str(_as_uuid(c["id_con"])) for c in rows})
    images = sorted({c["id_img"] for c in rows})
    if not rows:
        return {}
    u_idx = {u: i for i, u in enumerate(users)}
    i_idx = {h: i for i, h in enumerate(images)}
    a = np.fromiter((u_idx[str(_as_uuid(c["id_con"]))] for c in rows), dtype=np.int32, count=len(rows))
    im = np.fromiter((i_idx[c["id_img"]] for c in rows), dtype=np.int32, count=len(rows))
    act = np.fromiter((1 if c["ativo"] is True else 0 for c in rows), dtype=np.uint8, count=len(rows))
    d = labels.distinct_images_per_annotator(a, im, act, len(users), device)
    return {u: int(d[i]) for u, i in u_idx.items()}
