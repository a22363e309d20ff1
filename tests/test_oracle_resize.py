"""Oracle pinning, resize (parity unpinned by the reference: it has no resize).  The NumPy
restatement of Pillow's 8-bit two-pass resampler must reproduce Pillow bit for bit: on the
committed fixtures (Pillow outputs recorded by tests/golden/make_static_golden.py) and on fresh
random images against the Pillow installed here."""
import numpy as np
import pytest

from oracle import precompute_coeffs, preview_f32, resample_restated, thumbnail_u8


def test_restated_vs_committed_pillow_outputs(pillow_cases):
    for img, want in pillow_cases:
        got = resample_restated(img, want.shape[0], want.shape[1])
        assert np.array_equal(got, want), (img.shape, want.shape)


@pytest.mark.parametrize("shape,out", [((270, 480), (64, 64)), ((33, 57), (256, 256)), ((128, 128), (256, 256)),
                                       ((540, 960), (256, 256))])
def test_restated_vs_live_pillow(shape, out):
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, size=(*shape, 3), dtype=np.uint8)
    assert np.array_equal(resample_restated(img, *out), thumbnail_u8(img, *out))


def test_coeffs_sum_to_one():
    for a, b in [(1920, 256), (1080, 256), (512, 256), (100, 256), (3840, 256)]:
        bounds, kk, ks = precompute_coeffs(a, b)
        sums = kk.sum(axis=1)
        assert np.all(np.abs(sums - (1 << 22)) <= ks)      # each tap rounds by <= 0.5
        assert np.all(bounds[:, 0] >= 0) and np.all(bounds[:, 0] + bounds[:, 1] <= a)


def test_preview_layout():
    t = np.arange(2 * 3 * 3, dtype=np.uint8).reshape(2, 3, 3)
    p = preview_f32(t, mean=(0.5, 0.0, 0.0), inv_std=(2.0, 1.0, 1.0))
    assert p.shape == (3, 2, 3) and p.dtype == np.float32
    assert p[0, 0, 0] == np.float32((np.float32(0) * np.float32(1 / 255) - np.float32(0.5)) * np.float32(2.0))
    assert p[1, 0, 0] == np.float32(1) * np.float32(1 / 255)
