"""Oracle: label aggregation.  TEST INFRASTRUCTURE — see oracle/__init__.py.

Two groups:

(1) What the reference actually implements (per *user*), restated and pinned on
    tests/golden/reference_labels.json (reference functions run against stub sessions):
      app/crud/classificacao_crud.py:284-324   obter_classificacoes_imagens (group by image)
      app/api/routes/classificacoes.py:224-230 COUNT(DISTINCT id_img) of active rows
      app/api/routes/classificacoes.py:543-576 history grouping
      app/crud/classificacao_crud.py:411-420, 471-475  classification delta + counter rule
    Only rows with ``ativo == True`` count (classificacao_crud.py:115,314;
    classificacoes.py:227).

(2) What BASELINE.json requires but the reference lacks — cross-annotator per-image tally
    and Fleiss' kappa.  **Parity unpinned by the reference.**  NumPy ``bincount`` over
    active rows + the textbook formula evaluated in float64 from *integer* partials
    (SURVEY.md section 8(c)), pinned on the Fleiss (1971) worked example.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


# ------------------------------------------------------------------ (2) tally + kappa
def label_tally(
    image_idx: np.ndarray,
    class_idx: np.ndarray,
    active: Optional[np.ndarray],
    n_images: int,
    k: int,
) -> np.ndarray:
    """counts[N,k] int32: number of ACTIVE rows per (image, class).  Each active row is
    one rating (multi-choice environments allow several per annotator, models.py:79)."""
    image_idx = np.asarray(image_idx).astype(np.int64)
    class_idx = np.asarray(class_idx).astype(np.int64)
    if active is not None:
        m = np.asarray(active) != 0
        image_idx, class_idx = image_idx[m], class_idx[m]
    if image_idx.size:
        assert image_idx.min() >= 0 and image_idx.max() < n_images
        assert class_idx.min() >= 0 and class_idx.max() < k
    flat = np.bincount(image_idx * k + class_idx, minlength=n_images * k)
    assert flat.max(initial=0) < 2**31
    return flat.reshape(n_images, k).astype(np.int32)


def fleiss_partials(counts: np.ndarray) -> Dict[str, object]:
    """Integer partials, all exact in int64 (identical for any sharding of the images):
    ``class_totals[k]`` T_j; ``S2 = sum n_ij^2``; ``R = sum n_i``; ``n_rated`` images with
    n_i >= 1; ``n_pairs_images`` images with n_i >= 2; ``pairs = sum n_i (n_i - 1)``."""
    c = counts.astype(np.int64)
    n_i = c.sum(axis=1)
    return {
        "class_totals": c.sum(axis=0),
        "S2": int((c * c).sum()),
        "R": int(n_i.sum()),
        "n_rated": int((n_i >= 1).sum()),
        "n_pairs_images": int((n_i >= 2).sum()),
        "pairs": int((n_i * (n_i - 1)).sum()),
    }


def fleiss_kappa(class_totals: np.ndarray, S2: int, R: int, n_images: int, n_raters: int) -> float:
    """Classical Fleiss kappa for a constant number ``n`` of ratings per image:
    P_bar = (S2 - R) / (N n (n-1)),  P_e = sum_j (T_j / R)^2,  kappa = (P_bar - P_e)/(1 - P_e).
    float64 from integer inputs -> bit-identical on any GPU count."""
    p_bar = float(S2 - R) / float(n_images * n_raters * (n_raters - 1))
    pj = np.asarray(class_totals, dtype=np.float64) / float(R)
    p_e = float(np.sum(pj * pj))
    return (p_bar - p_e) / (1.0 - p_e)


def fleiss_kappa_general(counts: np.ndarray) -> float:
    """Variable ratings per image: P_i = (sum_j n_ij^2 - n_i)/(n_i (n_i - 1)) for
    n_i >= 2, mean over those images; P_e from the class totals of all ratings."""
    c = counts.astype(np.int64)
    n_i = c.sum(axis=1)
    m = n_i >= 2
    s2_i = (c * c).sum(axis=1)
    p_i = (s2_i[m] - n_i[m]).astype(np.float64) / (n_i[m] * (n_i[m] - 1)).astype(np.float64)
    p_bar = float(p_i.sum()) / float(m.sum())
    pj = c.sum(axis=0).astype(np.float64) / float(n_i.sum())
    p_e = float(np.sum(pj * pj))
    return (p_bar - p_e) / (1.0 - p_e)


def agreement_hist(counts: np.ndarray, bins: int = 1024) -> np.ndarray:
    """The integer form of sum_i P_i the product accumulates in its tally pass: bin n (2 <= n < bins) =
    sum over images with n_i = n of (sum_j n_ij^2 - n_i); bin 0 = images with n_i >= bins; bin 1 = 0.
    Graft-defined (general-n kappa without a second pass, exact under sharding)."""
    c = counts.astype(np.int64)
    n_i = c.sum(axis=1)
    num = (c * c).sum(axis=1) - n_i
    h = np.zeros(bins, dtype=np.int64)
    m = (n_i >= 2) & (n_i < bins)
    np.add.at(h, n_i[m], num[m])
    h[0] = int((n_i >= bins).sum())
    return h


# ------------------------------------------------------------------ (1) per-user paths
def group_by_image(rows: Sequence[Dict], id_con, content_hashes: Sequence[str]) -> Dict[str, List[Dict]]:
    """classificacao_crud.py:305-324.  ``rows`` = the ``classificacoes`` table in storage
    order, dicts with keys id_cla,id_con,id_img,id_opc,ativo.  Returns
    ``{content_hash: [row,...]}`` for rows of user ``id_con`` that are active and whose
    image is in ``content_hashes``; row order within a key = table order; ``{}`` when the
    image list is empty."""
    if not content_hashes:
        return {}
    wanted = set(content_hashes)
    out: Dict[str, List[Dict]] = {}
    for c in rows:
        if c["id_con"] == id_con and c["id_img"] in wanted and c["ativo"] is True:
            out.setdefault(c["id_img"], []).append(c)
    return out


def distinct_image_count(rows: Iterable[Dict], id_con) -> int:
    """classificacoes.py:224-230 — COUNT(DISTINCT id_img) WHERE id_con=? AND ativo."""
    return len({c["id_img"] for c in rows if c["id_con"] == id_con and c["ativo"] is True})


def history_grouping(joined: Sequence[Tuple[str, str, str]]) -> List[Dict]:
    """classificacoes.py:543-576.  ``joined`` = page of ``(content_hash, opcao_texto,
    id_opc)`` in query order.  Groups by image in first-seen order, de-duplicates option
    *texts* (the id is appended only when the text is new, :554-556), joins with ', '."""
    grouped: Dict[str, Dict] = {}
    for content_hash, texto, id_opc in joined:
        if content_hash in grouped:
            item = grouped[content_hash]
            if texto not in item["opcoes_lista"]:
                item["opcoes_lista"].append(texto)
                item["ids_opcoes"].append(str(id_opc))
        else:
            grouped[content_hash] = {
                "content_hash": content_hash,
                "opcoes_lista": [texto],
                "ids_opcoes": [str(id_opc)],
            }
    out = []
    for item in grouped.values():
        out.append({
            "content_hash": item["content_hash"],
            "ids_opcoes": item["ids_opcoes"],
            "opcao_escolhida": ", ".join(item["opcoes_lista"]),
        })
    return out


def classification_delta(
    active: Iterable, inactive: Iterable, wanted: Iterable
) -> Tuple[set, set, set, int, bool]:
    """classificacao_crud.py:411-420 and :471-475.  Returns ``(inativar, criar,
    reativar, total_novas, counter_incremented)``: the progress counter grows by one iff
    the image had no active row before and something was created or reactivated."""
    ativas, inativas, manter = set(active), set(inactive), set(wanted)
    inativar = ativas - manter
    criar = manter - ativas - inativas
    reativar = manter & inativas
    total_novas = len(criar)
    tinha = len(ativas) > 0
    inc = (total_novas > 0 or (bool(reativar) and not tinha)) and not tinha
    return inativar, criar, reativar, total_novas, inc


def encode_label_rows(rows: Sequence[Dict], image_hashes: Sequence[str], option_ids: Sequence[str]):
    """Dictionary-encode rows of table ``classificacoes`` (app/db/models.py:224-241) for the tally:
    ``id_img`` (char64 foreign key, :229) -> position of the hash in ``sorted(image_hashes)`` (-1 = not a stored
    image), ``id_opc`` (UUID, :231) -> position in the option UUIDs sorted by their 16 bytes (255 = not an option
    of the environment), ``ativo`` (:233, compared ``== True`` as classificacao_crud.py:314 does) -> 0/1.
    Graft-defined (the reference never builds these arrays): parity unpinned by the reference."""
    import uuid
    img_pos = {h: i for i, h in enumerate(sorted(image_hashes))}
    opt_sorted = sorted(uuid.UUID(str(o)).bytes for o in option_ids)
    opt_pos = {b: i for i, b in enumerate(opt_sorted)}
    n = len(rows)
    img = np.full(n, -1, dtype=np.int32)
    cls = np.full(n, 255, dtype=np.uint8)
    act = np.zeros(n, dtype=np.uint8)
    for r, row in enumerate(rows):
        img[r] = img_pos.get(row["id_img"], -1)
        cls[r] = opt_pos.get(uuid.UUID(str(row["id_opc"])).bytes, 255)
        act[r] = 1 if row["ativo"] is True else 0
    return img, cls, act
