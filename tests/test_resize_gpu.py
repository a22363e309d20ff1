"""Parity, thumbnail / preview (parity unpinned by the reference: it has no resize; oracle =
Pillow, its pinned image library).  north_star tolerance: +-1 LSB for uint8, 1e-5 relative for
float32 — the kernels reproduce Pillow's integer arithmetic, so the tests assert the stronger
bit-exact result and state the tolerance they would fall back to."""
import numpy as np
import pytest
import torch

from ics_b200 import engine
from oracle import precompute_coeffs, preview_f32, synth_image, thumbnail_u8

pytestmark = pytest.mark.gpu

U8_TOL = 0          # contract: <= 1 LSB; achieved: identical
F32_RTOL = 1e-5


def _check(imgs, out_h, out_w, mean=(0, 0, 0), inv_std=(1, 1, 1)):
    thumb, prev = engine.thumbnails(imgs, out_h, out_w, True, mean, inv_std)
    for i, im in enumerate(imgs):
        want = thumbnail_u8(im, out_h, out_w)
        diff = np.abs(thumb[i].astype(np.int16) - want.astype(np.int16))
        assert diff.max(initial=0) <= U8_TOL, (im.shape, int(diff.max()), float((diff > 0).mean()))
        np.testing.assert_allclose(prev[i], preview_f32(want, mean, inv_std), rtol=F32_RTOL, atol=1e-7)


def _drop_plans():
    with engine._plans_lock:
        plans = list(engine._plans.values())
        engine._plans.clear()
    for p in plans:
        p.close()


@pytest.fixture(params=["auto", "gather", "generic"])
def resize_path(request, monkeypatch):
    """auto = what a call gets: the band kernel with the scatter-form vertical pass for downscales and the
    gather-form one for upscales; gather = the band kernel with the gather-form vertical pass forced everywhere;
    generic = the thread-per-output-pixel fallback.  Plans are cached per shape and read some switches when
    they are created, so the cache is emptied around every case."""
    monkeypatch.setenv("B2_RESIZE_PATH", "1" if request.param == "generic" else "0")
    monkeypatch.setenv("B2_RESIZE_VSCAT", "0" if request.param == "gather" else "1")
    _drop_plans()
    yield request.param
    _drop_plans()


def test_committed_pillow_fixtures(pillow_cases, resize_path):
    for img, want in pillow_cases:
        thumb, _ = engine.thumbnails([img], want.shape[0], want.shape[1], want_preview=False)
        assert np.array_equal(thumb[0], want), (img.shape, want.shape)


def test_tap_tables_match_pillow_restatement():
    for (ih, iw) in [(1080, 1920), (512, 512), (2160, 3840), (300, 257), (100, 100)]:
        plan = engine.ResizePlan(ih, iw, 256, 256)
        for axis, size in ((0, iw), (1, ih)):
            b, k, ks = plan.taps(axis)
            ob, ok, oks = precompute_coeffs(size, 256)
            assert ks == oks and np.array_equal(b, ob) and np.array_equal(k, ok)
        plan.close()


@pytest.mark.parametrize("shape", [(512, 512), (1080, 1920), (256, 256), (300, 257), (2160, 3840), (97, 1031),
                                   (128, 96), (4096, 4096)])
def test_baseline_shapes_to_256(shape, resize_path):
    if resize_path == "generic" and shape[0] * shape[1] > 1080 * 1920:
        pytest.skip("generic kernel is the slow fallback; large shapes covered by auto")
    n = 1 if shape[0] >= 2160 else 2
    imgs = [synth_image(g, *shape) for g in range(n)]
    _check(imgs, 256, 256)


def test_mixed_shapes_one_call_and_normalisation():
    imgs = [synth_image(0, 512, 512), synth_image(1, 270, 480), synth_image(2, 512, 512), synth_image(3, 31, 17)]
    _check(imgs, 256, 256, mean=(0.485, 0.456, 0.406), inv_std=(1 / 0.229, 1 / 0.224, 1 / 0.225))


def test_non_square_outputs_and_upscale(resize_path):
    imgs = [synth_image(4, 200, 300)]
    _check(imgs, 64, 96)
    _check(imgs, 256, 128)
    _check([synth_image(5, 20, 30)], 256, 256)


def test_whole_image_ctas_and_bands_agree(resize_path):
    """A batch of >= 4 x SM-count images runs one CTA per image (every stage of the ring, every output row through
    the same three accumulators); a small batch of the same images is cut into bands.  Same pixels either way, and
    both equal Pillow."""
    n, H, W = 640, 203, 310
    g = torch.Generator(device="cuda").manual_seed(77)
    data = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda", generator=g)
    off = torch.arange(n, dtype=torch.int64, device="cuda") * (H * W * 3)
    plan = engine.get_plan(H, W, 48, 64)
    thumb, prev = plan.run(data.view(-1), off)
    few, fprev = plan.run(data.view(-1), off[:3].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(thumb[:3], few) and torch.equal(prev[:3], fprev)
    for i in (0, 1, 317, 639):
        assert np.array_equal(thumb[i].cpu().numpy(), thumbnail_u8(data[i].cpu().numpy(), 48, 64))


@pytest.mark.parametrize("in_shape,out_shape", [((1080, 1919), (256, 256)), ((1081, 1922), (255, 250)),
                                                ((4320, 7680), (256, 256)), ((700, 4500), (100, 256)),
                                                ((257, 256), (256, 256)), ((511, 300), (256, 129))])
def test_odd_widths_and_scales(in_shape, out_shape):
    """Row pitches that are not multiples of 4 or 16 (every funnel-shift phase, the ragged image tail), scale factors
    from 1.004 to 30 (tap capacities 4 .. 36 and the generic kernel beyond), non-square outputs."""
    _drop_plans()
    _check([synth_image(11, *in_shape)], *out_shape)
    _drop_plans()


def test_structured_images_extremes():
    black = np.zeros((1080, 1920, 3), np.uint8)
    white = np.full((1080, 1920, 3), 255, np.uint8)
    yy, xx = np.mgrid[0:1080, 0:1920]
    checker = (((xx // 3 + yy // 5) % 2) * 255).astype(np.uint8)[..., None].repeat(3, 2)
    _check([black, white, np.ascontiguousarray(checker)], 256, 256)


def test_device_batch_full_size_properties():
    """64 x 1080p on device: duplicates give identical thumbnails; a constant image stays
    constant; sampled images match Pillow."""
    n, H, W = 64, 1080, 1920
    g = torch.Generator(device="cuda").manual_seed(0xB200)
    data = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda", generator=g)
    data[7] = data[3]
    data[9] = 200
    off = torch.arange(n, dtype=torch.int64, device="cuda") * (H * W * 3)
    plan = engine.get_plan(H, W, 256, 256)
    thumb, prev = plan.run(data.view(-1), off)
    torch.cuda.synchronize()
    assert torch.equal(thumb[7], thumb[3])
    assert bool((thumb[9] == 200).all())
    assert torch.equal(prev[9], torch.full_like(prev[9], float(np.float32(200) * np.float32(1 / 255))))
    for i in (0, 3, 63):
        want = thumbnail_u8(data[i].cpu().numpy(), 256, 256)
        assert np.array_equal(thumb[i].cpu().numpy(), want)


@pytest.mark.parametrize("h,w,oh,ow,n,pad", [(90, 160, 32, 32, 3, 0), (45, 67, 16, 24, 600, 0), (45, 67, 16, 24, 5, 7),
                                            (45, 67, 16, 24, 600, 13), (20, 30, 64, 48, 4, 3), (90, 2000, 16, 40, 2, 1)])
def test_image_offsets_any_alignment(h, w, oh, ow, n, pad, resize_path):
    """Images packed back to back at arbitrary byte offsets (pad != 0: no image starts 16-byte aligned, so the band
    kernel takes its cooperative-copy path instead of the bulk copy); bands and whole-image CTAs; first and last
    image against Pillow."""
    from PIL import Image
    g = torch.Generator(device="cuda").manual_seed(9)
    L = h * w * 3
    stride = L + pad
    data = torch.randint(0, 256, (n * stride + 64,), dtype=torch.uint8, device="cuda", generator=g)
    off = torch.arange(n, dtype=torch.int64, device="cuda") * stride + (1 if pad else 0)
    plan = engine.get_plan(h, w, oh, ow)
    thumb, prev = plan.run(data, off)
    torch.cuda.synchronize()
    for i in (0, n // 2, n - 1):
        o = int(off[i])
        img = data[o:o + L].cpu().numpy().reshape(h, w, 3)
        want = np.asarray(Image.fromarray(img, "RGB").resize((ow, oh), Image.BILINEAR))
        assert np.array_equal(thumb[i].cpu().numpy(), want), i
        np.testing.assert_allclose(prev[i].cpu().numpy(), preview_f32(want, (0, 0, 0), (1, 1, 1)), rtol=F32_RTOL, atol=1e-7)
