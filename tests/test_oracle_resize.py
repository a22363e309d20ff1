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


# ---- the two reformulations the CUDA band kernel relies on (csrc/resize.cu), checked on the CPU over more
# ---- geometries than the GPU tests visit: limb dot products without a clip; scatter-form vertical pass
@pytest.mark.parametrize("in_hw,out_hw", [((1080, 1920), (256, 256)), ((97, 131), (32, 48)), ((300, 257), (64, 96)),
                                          ((512, 512), (256, 256)), ((257, 256), (256, 256)), ((65, 64), (64, 64)),
                                          ((1000, 37), (7, 5)), ((531, 257), (256, 129)), ((2160, 384), (256, 32))])
def test_kernel_arithmetic_model_equals_pillow(in_hw, out_hw):
    from oracle import kernel_model
    rng = np.random.default_rng(in_hw[0] * 7 + out_hw[1])
    img = rng.integers(0, 256, (*in_hw, 3), dtype=np.uint8)
    img[: in_hw[0] // 3] = 255                                   # saturated rows: the no-clip claim is about these
    got = kernel_model(img, *out_hw)
    assert got is not None, "a downscale must be eligible for the scatter form"
    assert np.array_equal(got, thumbnail_u8(img, *out_hw))


def test_scatter_form_eligibility_rule():
    """Every downscale (scale > 1) qualifies: windows advance monotonically, one output row ends per input row, never
    more than three over an input row, and every input row's taps are accounted for exactly once.  Upscales with more
    than one output row per input row do not qualify (the kernel gathers instead)."""
    from oracle import scatter_table
    for in_size in list(range(1, 40)) + [97, 255, 256, 257, 511, 1080, 2160, 4096]:
        for out_size in (1, 2, 3, 5, 16, 31, 64, 100, 256):
            if out_size >= in_size:
                continue
            table, ok = scatter_table(in_size, out_size)
            assert ok, (in_size, out_size)
            bounds, kk, _ = precompute_coeffs(in_size, out_size)
            assert int(table[:, :3].sum()) == int(kk.sum()), (in_size, out_size)
            ends = table[:, 3][table[:, 3] != 0] >> 2
            assert np.array_equal(ends, np.arange(out_size)), (in_size, out_size)
    for in_size in (20, 64, 200, 255):                     # upscales: several output rows end with the same input row
        assert scatter_table(in_size, 256)[1] is False
    assert scatter_table(64, 64)[1] is False               # same size: the clipped last window ends with the one before it
