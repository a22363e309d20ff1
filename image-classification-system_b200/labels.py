"""Label aggregation: cross-annotator per-image tally + Fleiss' kappa, and the bulk form of the
reference's per-user counts.

The reference aggregates only per user (app/crud/classificacao_crud.py:284-324,
app/api/routes/classificacoes.py:224-230); BASELINE.json configs 1,4,5 require the
cross-annotator tally and kappa.  Rows are the dictionary-encoded columns of table
``classificacoes`` (app/db/models.py:224-241): ``image_idx`` int32 (id_img), ``class_idx``
uint8 (id_opc), ``active`` uint8 (ativo).  Only active rows count (classificacao_crud.py:314).

kappa is computed on the host in float64 from INTEGER partials produced on the device, so it is
bit-identical for any number of GPUs once the partials are all-reduced (dist.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import numpy as np

from . import hostapi

import sys


def _engine():
    from . import engine                  # device-pointer layer (imports PyTorch)
    return engine


def _is_tensor(a) -> bool:
    """True for a PyTorch tensor — without importing PyTorch (if it is not loaded, nobody can hold one)."""
    t = sys.modules.get("torch")
    return t is not None and isinstance(a, t.Tensor)


@dataclass
class TallyResult:
    counts: np.ndarray            # int32 [n_images, k]
    class_totals: np.ndarray      # int64 [k]
    S2: int                       # sum n_ij^2
    R: int                        # sum n_i  (active rows tallied)
    n_rated: int                  # images with n_i >= 1
    n_pairs_images: int           # images with n_i >= 2
    pairs: int                    # sum n_i (n_i - 1)
    agree_hist: Optional[np.ndarray] = None   # int64 [B2_AGREE_BINS]: see fleiss_kappa_from_hist

    def kappa(self, n_images: int, n_raters: int) -> float:
        """Constant-n form: only meaningful when every image has exactly ``n_raters`` active ratings."""
        return fleiss_kappa(self.class_totals, self.S2, self.R, n_images, n_raters)

    def kappa_general(self) -> float:
        """Variable ratings per image, from the integer agreement histogram of the same pass."""
        return fleiss_kappa_from_hist(self.class_totals, self.R, self.n_pairs_images, self.agree_hist)


def fleiss_kappa(class_totals: Sequence[int], S2: int, R: int, n_images: int, n_raters: int) -> float:
    """Classical Fleiss kappa, constant ``n_raters`` ratings per image:
    P_bar = (S2 - R) / (N n (n - 1));  P_e = sum_j (T_j / R)^2;  kappa = (P_bar - P_e) / (1 - P_e)."""
    p_bar = float(int(S2) - int(R)) / float(int(n_images) * int(n_raters) * (int(n_raters) - 1))
    pj = np.asarray(class_totals, dtype=np.float64) / float(int(R))
    p_e = float(np.sum(pj * pj))
    return (p_bar - p_e) / (1.0 - p_e)


def fleiss_kappa_general(class_totals: Sequence[int], R: int, sum_pi: float, n_pairs_images: int) -> float:
    """Variable ratings per image: P_bar = mean of P_i over images with n_i >= 2 (``sum_pi`` from
    :func:`engine.fleiss_partials_device`), P_e from the class totals of all ratings."""
    p_bar = float(sum_pi) / float(n_pairs_images)
    pj = np.asarray(class_totals, dtype=np.float64) / float(int(R))
    p_e = float(np.sum(pj * pj))
    return (p_bar - p_e) / (1.0 - p_e)


def sum_pi_from_hist(agree_hist) -> float:
    """sum_i P_i over images with n_i >= 2 from the agreement histogram of ``b2_label_tally``: bin n holds
    sum (sum_j n_ij^2 - n_i) over the images with n_i = n, so sum_i P_i = sum_n bin[n] / (n (n - 1)) — float64 from
    integers in a fixed order (ascending n): bit-identical for any sharding once the bins are all-reduced.  Raises
    when images with >= B2_AGREE_BINS ratings exist (bin 0): use ``engine.fleiss_partials_device(want_sum_pi=True)``."""
    h = np.asarray(agree_hist, dtype=np.int64)
    if int(h[0]) != 0:
        raise ValueError(f"{int(h[0])} images have >= {h.shape[0]} ratings: the histogram does not cover them")
    total = 0.0
    for n in np.nonzero(h[2:])[0] + 2:
        total += float(int(h[n])) / float(int(n) * (int(n) - 1))
    return total


def fleiss_kappa_from_hist(class_totals: Sequence[int], R: int, n_pairs_images: int, agree_hist) -> float:
    """General-n Fleiss kappa (P_bar = mean of P_i over images with n_i >= 2) from integers only."""
    return fleiss_kappa_general(class_totals, R, sum_pi_from_hist(agree_hist), n_pairs_images)


def _rows_to_device(image_idx, class_idx, active, device):
    import torch
    dev = torch.device("cuda", _engine().init(device))

    def up(a, dt):
        if isinstance(a, torch.Tensor):
            return a.to(dev, dtype=dt, non_blocking=True).contiguous()
        arr = np.ascontiguousarray(a, dtype={torch.int32: np.int32, torch.uint8: np.uint8}[dt])
        return torch.from_numpy(arr).to(dev, non_blocking=True)

    return up(image_idx, torch.int32), up(class_idx, torch.uint8), up(active, torch.uint8)


def label_tally(image_idx, class_idx, active, n_images: int, k: int, sorted_by_image: Optional[bool] = None,
                image_base: int = 0, device: Optional[int] = None) -> TallyResult:
    """Tally host (NumPy) or device rows.  ``sorted_by_image=None`` checks the order on the host
    for NumPy input (cheap) and assumes clustered input for device tensors."""
    if active is None:
        active = np.ones(len(image_idx), dtype=np.uint8)
    if sorted_by_image is None:
        if isinstance(image_idx, np.ndarray):
            sorted_by_image = bool(np.all(image_idx[1:] >= image_idx[:-1])) if image_idx.size else True
        else:
            sorted_by_image = True
    from ._lib import B2_AGREE_BINS
    if not any(_is_tensor(a) for a in (image_idx, class_idx, active)):
        # host rows: one native call with host pointers (b2_label_tally_host stages, tallies, reads back, checks)
        hist = np.zeros(B2_AGREE_BINS, dtype=np.int64)
        counts, p = label_tally_host(image_idx, class_idx, active, n_images, k, sorted_by_image, image_base, device,
                                     agree_hist=hist)
    else:
        import torch
        d_img, d_cls, d_act = _rows_to_device(image_idx, class_idx, active, device)
        d_hist = torch.empty(B2_AGREE_BINS, dtype=torch.int64, device=d_img.device)
        d_counts, partials = _engine().label_tally_device(d_img, d_cls, d_act, n_images, k, image_base, sorted_by_image,
                                                          agree_hist=d_hist)
        p = partials.cpu().numpy()
        hostapi.check_tally(p, k, d_img.numel())
        counts = d_counts.cpu().numpy()
        hist = d_hist.cpu().numpy()
    d = hostapi.partials_dict(p, k)
    return TallyResult(counts=counts, class_totals=d["class_totals"], S2=d["S2"], R=d["R"],
                       n_rated=d["n_rated"], n_pairs_images=d["n_pairs_images"], pairs=d["pairs"], agree_hist=hist)


label_tally_host = hostapi.label_tally_host


def distinct_images_per_annotator(annotator_idx, image_idx, active, n_annotators: int,
                                  device: Optional[int] = None) -> np.ndarray:
    """Bulk ``COUNT(DISTINCT id_img) WHERE id_con = ? AND ativo`` (routes/classificacoes.py:
    224-230) for every annotator; rows in any order (sorted here by (annotator, image))."""
    a = np.ascontiguousarray(annotator_idx, dtype=np.int32)
    i = np.ascontiguousarray(image_idx, dtype=np.int32)
    act = np.ascontiguousarray(active, dtype=np.uint8)
    order = np.lexsort((i, a))
    # host rows: one native call with host pointers (no PyTorch needed — the crud mirror runs where it is absent)
    return hostapi.distinct_images_host(a[order], i[order], act[order], n_annotators, device).astype(np.int32)


class LabelEncoder:
    """Dictionary encoder for ``classificacoes`` rows: id_img
    (char64) -> dense image index, id_opc (UUID) -> class index, ativo -> uint8."""

    def __init__(self, image_hashes: Sequence[str], option_ids: Sequence[str]):
        self.image_index: Dict[str, int] = {h: i for i, h in enumerate(image_hashes)}
        self.class_index: Dict[str, int] = {str(o): j for j, o in enumerate(option_ids)}
        if len(self.class_index) > 256:
            raise ValueError("class_idx is uint8: at most 256 options per environment")

    def encode(self, rows: Sequence[Dict]):
        """rows: dicts with id_img, id_opc, ativo.  Returns SoA arrays sorted by image index."""
        n = len(rows)
        img = np.fromiter((self.image_index[r["id_img"]] for r in rows), dtype=np.int32, count=n)
        cls = np.fromiter((self.class_index[str(r["id_opc"])] for r in rows), dtype=np.uint8, count=n)
        act = np.fromiter((1 if r["ativo"] is True else 0 for r in rows), dtype=np.uint8, count=n)
        order = np.argsort(img, kind="stable")
        return img[order], cls[order], act[order]


class DeviceLabelEncoder:
    """Bulk form of :class:`LabelEncoder`: the dictionaries live on the device
    as sorted key tables — stored image digests (the table ``b2_dedupe`` already uses) and the environment's option
    UUIDs — and whole row sets are encoded by ``b2_encode_label_rows``.  Image index = position of the hash in
    ``sorted(image_hashes)``, class index = position of the option among the UUIDs sorted by their bytes."""

    def __init__(self, image_hashes: Sequence[str], option_ids: Sequence[str], device: Optional[int] = None):
        import uuid
        import torch
        self.dev = torch.device("cuda", _engine().init(device))
        self.image_hashes = sorted(image_hashes)
        keys = np.frombuffer(bytes.fromhex("".join(self.image_hashes)), dtype=np.uint8).reshape(-1, 32) \
            if self.image_hashes else np.zeros((0, 32), np.uint8)
        self.option_ids = sorted((uuid.UUID(str(o)) for o in option_ids), key=lambda u: u.bytes)
        if len(self.option_ids) > 255:
            raise ValueError("class_idx is uint8 and 255 marks an unknown option: at most 255 options per environment")
        opts = np.frombuffer(b"".join(u.bytes for u in self.option_ids), dtype=np.uint8).reshape(-1, 16) \
            if self.option_ids else np.zeros((0, 16), np.uint8)
        self.d_image_keys = torch.from_numpy(np.array(keys, copy=True)).to(self.dev)       # hex order = byte order
        self.d_option_keys = torch.from_numpy(np.array(opts, copy=True)).to(self.dev)

    def encode_columns(self, img_hex, opc_uuid, ativo, sort: bool = True):
        """Raw key columns (device or host tensors: uint8[R,64], uint8[R,16], uint8[R]) -> device SoA arrays
        ``(image_idx, class_idx, active, unknown)``; with ``sort`` the rows come back ordered by image index
        (stable), ready for the sorted-mode tally."""
        import torch
        cols = [c.to(self.dev, non_blocking=True).contiguous() for c in (img_hex, opc_uuid, ativo)]
        img, cls, act, unknown = _engine().encode_label_rows_device(cols[0], cols[1], cols[2], self.d_image_keys,
                                                                    self.d_option_keys)
        if sort:
            order = torch.sort(img, stable=True).indices
            img, cls, act = img[order].contiguous(), cls[order].contiguous(), act[order].contiguous()
        return img, cls, act, unknown

    def encode(self, rows: Sequence[Dict], sort: bool = True):
        """Row dicts with id_img (64-char hex), id_opc (UUID or its string), ativo -> the same as
        :meth:`encode_columns`."""
        import uuid
        import torch
        n = len(rows)
        hexs = np.frombuffer("".join(str(r["id_img"]).ljust(64)[:64] for r in rows).encode("latin-1", "replace"),
                             dtype=np.uint8).reshape(n, 64) if n else np.zeros((0, 64), np.uint8)
        opcs = np.frombuffer(b"".join(uuid.UUID(str(r["id_opc"])).bytes for r in rows), dtype=np.uint8).reshape(n, 16) \
            if n else np.zeros((0, 16), np.uint8)
        act = np.fromiter((1 if r["ativo"] is True else 0 for r in rows), dtype=np.uint8, count=n)
        return self.encode_columns(torch.from_numpy(hexs.copy()), torch.from_numpy(opcs.copy()), torch.from_numpy(act), sort)
