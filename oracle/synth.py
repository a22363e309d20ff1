"""Seeded synthetic inputs of the BASELINE.json shapes.
TEST INFRASTRUCTURE — see oracle/__init__.py.  Host-side (NumPy) generators; the bench
generates its large on-device batches with torch and copies a sample back for the
oracle, tests use these directly.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

IMAGE_SEED = 0xB200
LABEL_SEED = 0xF1E155


def synth_image(g: int, h: int, w: int, seed: int = IMAGE_SEED) -> np.ndarray:
    """Image ``g``: HxWx3 uint8 HWC, i.i.d. uniform bytes keyed (seed, g).  The same
    buffer is both the "file bytes" that get hashed and the decoded RGB that gets resized
    (BASELINE configs say synthetic RGB images; decode is outside the path)."""
    rng = np.random.Generator(np.random.Philox(key=[seed, g]))
    return rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)


def synth_images(n: int, h: int, w: int, seed: int = IMAGE_SEED, start: int = 0) -> List[np.ndarray]:
    return [synth_image(start + g, h, w, seed) for g in range(n)]


def synth_duplicate_map(n: int, n_unique: int) -> np.ndarray:
    """C5 duplicate rule: image g >= n_unique is a byte copy of image
    (g * 2654435761 mod n_unique)  ->  exactly n_unique distinct contents."""
    src = np.arange(n, dtype=np.int64)
    dup = src >= n_unique
    src[dup] = (src[dup] * 2654435761) % n_unique
    return src


def synth_label_rows(
    n_images: int,
    k: int,
    n_raters: int,
    seed: int = LABEL_SEED,
    p_true: float = 0.7,
    p_active: float = 0.95,
    shuffled: bool = False,
) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """SoA label rows: ``image_idx int32[R]`` (clustered by image: row // n_raters),
    ``class_idx uint8[R]`` (the image's "true" class w.p. ``p_true`` else uniform, so kappa
    is non-trivial), ``active uint8[R]`` (1 w.p. ``p_active``); R = n_images * n_raters."""
    rng = np.random.Generator(np.random.Philox(key=[seed, n_images]))
    rows = n_images * n_raters
    image_idx = (np.arange(rows, dtype=np.int64) // n_raters).astype(np.int32)
    true_cls = rng.integers(0, k, size=n_images, dtype=np.int64)
    pick_true = rng.random(rows) < p_true
    uniform = rng.integers(0, k, size=rows, dtype=np.int64)
    class_idx = np.where(pick_true, true_cls[image_idx], uniform).astype(np.uint8)
    active = (rng.random(rows) < p_active).astype(np.uint8)
    if shuffled:
        perm = rng.permutation(rows)
        image_idx, class_idx, active = image_idx[perm], class_idx[perm], active[perm]
    return image_idx, class_idx, active
