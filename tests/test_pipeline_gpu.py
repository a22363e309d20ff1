"""End-to-end parity of the host-buffer pipeline and of the remaining BASELINE config shapes
(config 3: mixed sizes 256^2..4096^2; config 5: 4K images with the 20 % duplicate rule + thumbnails +
tally), each against the oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

import ics_b200
from ics_b200.pipeline import IngestPipeline
from oracle import (dedupe_batch, label_tally, preview_f32, sha256_digest, sha256_hex, synth_duplicate_map, synth_image,
                    synth_label_rows, thumbnail_u8)

pytestmark = pytest.mark.gpu


def _pinned(images):
    flat = np.stack([im.reshape(-1) for im in images])
    t = torch.empty(flat.shape, dtype=torch.uint8, pin_memory=True)
    t.copy_(torch.from_numpy(flat))
    return t


@pytest.mark.parametrize("shape,n,chunk", [((96, 128), 37, 8), ((270, 480), 20, 7), ((1080, 1920), 12, 5)])
def test_pipeline_matches_oracle(shape, n, chunk):
    src = synth_duplicate_map(n, n - n // 5)
    base = [synth_image(g, *shape) for g in range(n - n // 5)]
    images = [base[int(s)] for s in src]
    host = _pinned(images)
    pipe = IngestPipeline(shape[0], shape[1], n, chunk_images=chunk, n_streams=3)
    for _ in range(2):                                    # a pipeline object is reusable
        res = pipe.run(host)
        hashes = [sha256_hex(im.tobytes()) for im in images]
        is_new, _, stats = dedupe_batch(hashes)
        assert [bytes(d).hex() for d in res.digests.numpy()] == hashes
        assert [bool(x) for x in res.is_new.numpy()] == is_new
        assert res.stats == stats
        for i in (0, n // 2, n - 1):
            want = thumbnail_u8(images[i], 256, 256)
            assert np.array_equal(res.thumbs[i].numpy(), want)
            np.testing.assert_allclose(res.previews[i].numpy(), preview_f32(want), rtol=1e-5, atol=1e-7)
        assert res.h2d_bytes == n * shape[0] * shape[1] * 3


def test_two_pipelines_in_flight():
    shape, n = (128, 160), 24
    a = [synth_image(g, *shape) for g in range(n)]
    b = [synth_image(1000 + g, *shape) for g in range(n)]
    ha, hb = _pinned(a), _pinned(b)
    pa, pb = IngestPipeline(*shape, n, chunk_images=5), IngestPipeline(*shape, n, chunk_images=5)
    pa.submit(ha)
    pb.submit(hb)                                         # second batch enqueued before the first is read
    ra, rb = pa.result(), pb.result()
    assert [bytes(d) for d in ra.digests.numpy()] == [sha256_digest(im.tobytes()) for im in a]
    assert [bytes(d) for d in rb.digests.numpy()] == [sha256_digest(im.tobytes()) for im in b]
    assert np.array_equal(rb.thumbs[3].numpy(), thumbnail_u8(b[3], 256, 256))


def test_config3_mixed_sizes():
    """Config 3 shape: square images of side 256 * 2^j, j = 0..4 (one 4096^2 image keeps the oracle fast)."""
    sides = [256, 512, 1024, 2048, 4096, 512, 256, 1024]
    images = [synth_image(g, s, s) for g, s in enumerate(sides)]
    res = ics_b200.ingest_batch([im.tobytes() for im in images], decoded_rgb=images)
    assert res.decision.hashes == [sha256_hex(im.tobytes()) for im in images]
    assert res.decision.stats == {"processed": 8, "created": 8, "updated": 0}
    for i, im in enumerate(images):
        assert np.array_equal(res.thumbs[i], thumbnail_u8(im, 256, 256)), sides[i]


def test_config5_4k_duplicates_thumbnails_tally():
    """Config 5 shape at reduced count: 3840x2160 images, 20 % byte-copies, thumbnails, and the label tally of
    the rows that reference them — one sync batch end to end."""
    n, nu = 10, 8
    src = synth_duplicate_map(n, nu)
    base = [synth_image(g, 2160, 3840) for g in range(nu)]
    images = [base[int(s)] for s in src]
    res = ics_b200.ingest_batch([im.tobytes() for im in images], decoded_rgb=images, want_preview=False)
    hashes = [sha256_hex(im.tobytes()) for im in images]
    is_new, first, stats = dedupe_batch(hashes)
    assert res.decision.hashes == hashes and res.decision.is_new == is_new and res.decision.first_index == first
    assert stats == res.decision.stats == {"processed": 10, "created": 8, "updated": 2}
    for i in (0, 5, 9):
        assert np.array_equal(res.thumbs[i], thumbnail_u8(images[i], 256, 256))
    # label rows keyed by the dense index of the UNIQUE images (what the dedupe table yields)
    img, cls, act = synth_label_rows(nu, 50, 100)
    t = ics_b200.label_tally(img, cls, act, nu, 50)
    assert np.array_equal(t.counts, label_tally(img, cls, act, nu, 50))
