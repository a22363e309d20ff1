"""Oracle: thumbnail / preview tensor.  TEST INFRASTRUCTURE — see oracle/__init__.py.

**Parity unpinned by the reference**: the reference performs no resize anywhere
(SURVEY.md section 0); BASELINE.json's configs require "hash + 256x256 thumbnail", so the
oracle is the reference's pinned image library called directly — Pillow
(requirements.txt:9 pins pillow==10.1.0; this image has 12.2.0, whose 8-bit resampler
is the same algorithm) — as SURVEY.md section 8(c) defines:

    thumb   = Image.fromarray(rgb, 'RGB').resize((out_w, out_h), Image.BILINEAR)
    preview = ((thumb / 255) - mean) * inv_std, HWC -> CHW, float32

``resample_restated`` restates Pillow's published two-pass algorithm
(libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
ImagingResampleHorizontal_8bpc / Vertical_8bpc) in NumPy so that the CUDA kernel has an
exact integer specification; tests check it against Pillow itself (0 differing pixels).
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # 22-bit fixed-point coefficients


def thumbnail_u8(rgb_hwc: np.ndarray, out_h: int = 256, out_w: int = 256) -> np.ndarray:
    """Pillow BILINEAR (antialiased when down-scaling), non-aspect-preserving."""
    from PIL import Image

    assert rgb_hwc.dtype == np.uint8 and rgb_hwc.ndim == 3 and rgb_hwc.shape[2] == 3
    img = Image.fromarray(np.ascontiguousarray(rgb_hwc), "RGB")
    return np.asarray(img.resize((out_w, out_h), Image.BILINEAR))


def preview_f32(
    thumb_hwc_u8: np.ndarray,
    mean: Sequence[float] = (0.0, 0.0, 0.0),
    inv_std: Sequence[float] = (1.0, 1.0, 1.0),
) -> np.ndarray:
    """float32 CHW: ``(u8 * (1/255) - mean) * inv_std`` evaluated in float32, in that order
    (the CUDA kernel performs the same three float32 operations, so results are
    bit-identical in practice; the test tolerance is 1e-5 relative per north_star)."""
    x = thumb_hwc_u8.astype(np.float32) * np.float32(1.0 / 255.0)
    m = np.asarray(mean, dtype=np.float32).reshape(1, 1, 3)
    s = np.asarray(inv_std, dtype=np.float32).reshape(1, 1, 3)
    return np.ascontiguousarray(((x - m) * s).transpose(2, 0, 1))


def precompute_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Triangle-filter taps for one axis, Pillow style.

    Returns ``(bounds[out,2] int32 = (first input index, tap count), kk[out,ksize] int32
    22-bit fixed point, ksize)``.  All arithmetic in float64 in the same order as
    Pillow's precompute_coeffs so the integer coefficients are identical.
    """
    scale = float(in_size) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale            # bilinear filter support = 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        n = xmax - xmin
        w = np.empty(n, dtype=np.float64)
        ww = 0.0
        for x in range(n):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            v = 1.0 - a if a < 1.0 else 0.0
            w[x] = v
            ww += v                         # sequential double sum, like Pillow
        if ww != 0.0:
            w = w / ww
        for x in range(n):
            v = w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx, 0] = xmin
        bounds[xx, 1] = n
    return bounds, kk, ksize


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resample_restated(rgb_hwc: np.ndarray, out_h: int = 256, out_w: int = 256) -> np.ndarray:
    """Horizontal pass -> uint8 intermediate -> vertical pass, 22-bit fixed point,
    rounding constant 2^21 added before the shift — bit-identical to Pillow's 8-bit
    path (a pass whose in == out size is an identity and Pillow skips it)."""
    assert rgb_hwc.dtype == np.uint8 and rgb_hwc.ndim == 3
    in_h, in_w, ch = rgb_hwc.shape
    src = rgb_hwc
    if in_w != out_w:
        hb, hk, _ = precompute_coeffs(in_w, out_w)
        tmp = np.empty((in_h, out_w, ch), dtype=np.uint8)
        s64 = src.astype(np.int64)
        for x in range(out_w):
            x0, n = int(hb[x, 0]), int(hb[x, 1])
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(
                s64[:, x0:x0 + n, :], hk[x, :n].astype(np.int64), axes=([1], [0]))
            tmp[:, x, :] = _clip8(acc)
        src = tmp
    if in_h != out_h:
        vb, vk, _ = precompute_coeffs(in_h, out_h)
        out = np.empty((out_h, src.shape[1], ch), dtype=np.uint8)
        s64 = src.astype(np.int64)
        for y in range(out_h):
            y0, n = int(vb[y, 0]), int(vb[y, 1])
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(
                vk[y, :n].astype(np.int64), s64[y0:y0 + n, :, :], axes=([0], [0]))
            out[y] = _clip8(acc)
        src = out
    return np.ascontiguousarray(src)


# ---------------------------------------------------------------------------------------------
# Model of the CUDA band kernel's arithmetic (csrc/resize.cu), so that the two reformulations it
# relies on are checked on the CPU too, over many more shapes than the GPU tests visit:
#   * horizontal pass: a 22-bit tap as three 8-bit limbs, one unsigned dot product per limb
#     (IDP.4A), recombined as s0 + (s1 << 8) + (s2 << 16) in uint32, NO clip;
#   * vertical pass in scatter form: per INPUT row the taps of the <= 3 output rows whose
#     windows cover it; output row oy accumulates in slot oy % 3; the uint32 accumulators are
#     never reset — they run on modulo 2^32 and a second value remembers where the row started.
# ---------------------------------------------------------------------------------------------
def scatter_table(in_size: int, out_size: int):
    """``(table int64[in_size, 4], eligible)`` — the table ``b2_resize_plan_create`` builds for
    ``kVMode == 2``: columns 0..2 = tap of the output row living in accumulator 0, 1, 2 for this
    input row (0 when none), column 3 = ``(oy << 2) | (oy % 3 + 1)`` of the output row that ENDS
    with this input row, else 0.  ``eligible`` is False when the plan must use the gather form:
    windows not monotone, more than three output rows over one input row, or a negative tap."""
    bounds, kk, _ = precompute_coeffs(in_size, out_size)
    first = bounds[:, 0].astype(np.int64)
    last = first + bounds[:, 1].astype(np.int64) - 1
    table = np.zeros((in_size, 4), dtype=np.int64)
    ok = bool(np.all(first[:-1] <= first[1:]) and np.all(last[:-1] < last[1:])) if out_size > 1 else True
    lo = 0
    for r in range(in_size):
        while lo < out_size and last[lo] < r:
            lo += 1
        for j in range(3):
            oy = lo + j
            if oy < out_size and first[oy] <= r <= last[oy]:
                c = int(kk[oy, r - first[oy]])
                ok = ok and c >= 0
                table[r, oy % 3] = c
        if lo + 3 < out_size and first[lo + 3] <= r:
            ok = False
        if lo < out_size and last[lo] == r:
            table[r, 3] = (lo << 2) | (lo % 3 + 1)
    return table, ok


def kernel_model(rgb_hwc: np.ndarray, out_h: int, out_w: int):
    """What ``resize_bands_kernel<KQ, 2>`` computes, in its own arithmetic; ``None`` when the
    scatter form does not apply to this vertical geometry (the kernel then gathers)."""
    in_h, in_w, _ = rgb_hwc.shape
    table, ok = scatter_table(in_h, out_h)
    if not ok:
        return None
    hb, hk, _ = precompute_coeffs(in_w, out_w)
    src = rgb_hwc.astype(np.uint64)
    mask32 = np.uint64(0xFFFFFFFF)
    rnd = np.uint64(1 << (PRECISION_BITS - 1))
    inter = np.empty((in_h, out_w, 3), dtype=np.uint64)
    for x in range(out_w):
        x0, n = int(hb[x, 0]), int(hb[x, 1])
        k = hk[x, :n].astype(np.uint64)
        assert (hk[x, :n] >= 0).all()
        limbs = [(k >> np.uint64(8 * i)) & np.uint64(0xFF) for i in range(3)]
        s = [np.tensordot(src[:, x0:x0 + n, :], l, axes=([1], [0])) for l in limbs]     # three dot products
        acc = (rnd + s[0] + (s[1] << np.uint64(8)) + (s[2] << np.uint64(16))) & mask32
        inter[:, x, :] = acc >> np.uint64(PRECISION_BITS)                                # no clip
    assert int(inter.max(initial=0)) <= 255
    out = np.zeros((out_h, out_w, 3), dtype=np.uint8)
    va = np.full((3, out_w, 3), int(rnd), dtype=np.uint64)       # three output rows in flight
    vb = np.zeros((3, out_w, 3), dtype=np.uint64)
    written = np.zeros(out_h, dtype=bool)
    for r in range(in_h):
        for j in range(3):
            va[j] = (va[j] + inter[r] * np.uint64(table[r, j])) & mask32
        w = int(table[r, 3])
        if w:
            oy, j = w >> 2, (w & 3) - 1
            res = ((va[j] - vb[j]) & mask32) >> np.uint64(PRECISION_BITS)
            assert int(res.max(initial=0)) <= 255 and not written[oy]
            out[oy] = res.astype(np.uint8)
            written[oy] = True
            vb[j] = (va[j] - rnd) & mask32
    assert written.all()
    return out
