"""Label-aggregation functions of ``app/crud/classificacao_crud.py`` and
``app/api/routes/classificacoes.py`` that sit on the hot path, with the reference's names.

  obter_classificacoes_imagens(db, id_con, imagens) -> dict      classificacao_crud.py:284-324
  obter_contagem_classificacoes(db, id_con) -> {"total": int}    routes/classificacoes.py:204-236
  agrupar_historico(resultados, id_amb) -> items                 routes/classificacoes.py:543-576 (the grouping
                                                                 loop of listar_historico_usuario)
  calcular_delta_classificacao(ativas, inativas, novas)          classificacao_crud.py:411-420, 471-475 (the set
                                                                 deltas and the progress-counter rule of
                                                                 criar_ou_atualizar_classificacao)

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
from typing import Dict, Iterable, List, Optional, Sequence, Set, Tuple
from urllib.parse import quote

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


def _get(row, name):
    return row[name] if isinstance(row, dict) else getattr(row, name)


def agrupar_historico(resultados: Sequence[Tuple], id_amb: Optional[str] = None) -> List[Dict]:
    """The grouping loop of ``listar_historico_usuario`` (routes/classificacoes.py:543-576) over one page
    (<= 100 joined rows, in query order) of ``(classificacao, imagem, opcao, conjunto, ambiente)`` — row objects
    or dicts with the reference's column names.  Groups by ``content_hash`` in first-seen order; an option is
    appended only when its TEXT is new for the image (:554-556); ``opcao_escolhida`` joins the texts with ", ";
    the image URL is the percent-quoted path under ``/nextcloud/images/`` (:560-561)."""
    grouped: Dict[str, Dict] = {}
    for classificacao, imagem, opcao, _conjunto, ambiente in resultados:
        final_id_amb = id_amb if id_amb else str(_get(ambiente, "id_amb"))
        h = _get(imagem, "content_hash")
        texto = _get(opcao, "texto")
        if h in grouped:
            item = grouped[h]
            if texto not in item["opcoes_lista"]:
                item["opcoes_lista"].append(texto)
                item["ids_opcoes"].append(str(_get(opcao, "id_opc")))
        else:
            path_limpo = _get(imagem, "caminho_img").lstrip("/")
            grouped[h] = {
                "content_hash": h,
                "nome_img": _get(imagem, "nome_img"),
                "url_img": f"/nextcloud/images/{quote(path_limpo, safe='/')}",
                "opcoes_lista": [texto],
                "ids_opcoes": [str(_get(opcao, "id_opc"))],
                "data_classificacao": _get(classificacao, "data_criado"),
                "nome_ambiente": _get(ambiente, "titulo_amb"),
                "id_amb": final_id_amb,
            }
    items = []
    for item in grouped.values():
        item["opcao_escolhida"] = ", ".join(item["opcoes_lista"])
        del item["opcoes_lista"]
        items.append(item)
    return items


def calcular_delta_classificacao(ativas: Iterable, inativas: Iterable, novas: Iterable) -> Tuple[Set, Set, Set, int, int]:
    """The arithmetic of ``criar_ou_atualizar_classificacao`` (classificacao_crud.py:411-420, :471-475) for one
    (user, image): option ids with an active row, with an inactive row, and wanted now ->
    ``(inativar, criar, reativar, total_novas, counter_delta)``.  ``total_classificadas`` grows by one iff the image
    had no active row before and something was created or reactivated."""
    ativas, inativas, manter = set(ativas), set(inativas), set(novas)
    inativar = ativas - manter
    criar = manter - ativas - inativas
    reativar = manter & inativas
    total_novas = len(criar)
    tinha_classificacao = len(ativas) > 0
    delta = 0
    if total_novas > 0 or (reativar and not tinha_classificacao):
        if not tinha_classificacao:
            delta = 1
    return inativar, criar, reativar, total_novas, delta
