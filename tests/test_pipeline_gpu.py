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
    pipe = IngestPipeline(shape[0], shape[1], n, chunk_images=chunk)
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


def test_existing_table_and_occurrence_indices():
    """The sorted table of digests already stored decides created/updated exactly as the sequential reference
    loop does; first/last occurrence indices come back too."""
    from ics_b200 import engine
    shape, n = (64, 80), 30
    src = synth_duplicate_map(n, 20)
    base = [synth_image(500 + g, *shape) for g in range(20)]
    images = [base[int(s)] for s in src]
    hashes = [sha256_hex(im.tobytes()) for im in images]
    stored = {hashes[1], hashes[7], sha256_hex(b"not in this batch")}
    table = engine.sort_digests(np.frombuffer(bytes.fromhex("".join(sorted(stored))), dtype=np.uint8).reshape(-1, 32))
    res = IngestPipeline(*shape, n, chunk_images=4).run(_pinned(images), torch.from_numpy(table))
    is_new, first, stats = dedupe_batch(hashes, stored)
    assert [bool(x) for x in res.is_new.numpy()] == is_new
    assert res.first_index.tolist() == first and res.stats == stats
    last = [max(j for j in range(n) if hashes[j] == h) for h in hashes]
    assert res.last_index.tolist() == last


def test_native_stream_from_plain_ctypes_and_numpy():
    """What a maintainer of the reference would write: no tensor library, the C ABI with host
    pointers only — page-locked memory from b2_host_alloc viewed as NumPy arrays."""
    import ctypes as C
    lib = C.CDLL(ics_b200.LIB_PATH)
    lib.b2_last_error.restype = C.c_char_p
    shape, n = (96, 112), 9
    L = shape[0] * shape[1] * 3
    images = [synth_image(900 + g, *shape) for g in range(n)]

    def pinned(nbytes, dtype):
        p = C.c_void_p()
        assert lib.b2_host_alloc(C.byref(p), C.c_uint64(nbytes)) == 0, lib.b2_last_error()
        return p, np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (nbytes,)).view(dtype)

    p_in, h_in = pinned(n * L, np.uint8)
    h_in[:] = np.concatenate([im.reshape(-1) for im in images])
    p_dig, h_dig = pinned(n * 32, np.uint8)
    p_new, h_new = pinned(n, np.uint8)
    p_cnt, h_cnt = pinned(16, np.uint32)
    p_th, h_th = pinned(n * 256 * 256 * 3, np.uint8)
    st = C.c_void_p()
    assert lib.b2_ingest_stream_create(0, shape[0], shape[1], 256, 256, n, 4, 0, C.byref(st)) == 0, lib.b2_last_error()
    assert lib.b2_ingest_stream_submit(st, p_in, n, None, C.c_uint64(0), p_dig, p_new, None, None, p_cnt, p_th, None) == 0, \
        lib.b2_last_error()
    assert lib.b2_ingest_stream_submit(st, p_in, n, None, C.c_uint64(0), p_dig, p_new, None, None, p_cnt, p_th, None) == -1  # busy
    h2d, d2h, launches = C.c_uint64(), C.c_uint64(), C.c_uint32()
    assert lib.b2_ingest_stream_wait(st, C.byref(h2d), C.byref(d2h), C.byref(launches)) == 0
    assert h2d.value == n * L and launches.value == 2 * 3 + 2
    assert [bytes(d).hex() for d in h_dig.reshape(n, 32)] == [sha256_hex(im.tobytes()) for im in images]
    assert h_new.tolist() == [1] * n and h_cnt[:3].tolist() == [n, n, 0]
    assert np.array_equal(h_th.reshape(n, 256, 256, 3)[5], thumbnail_u8(images[5], 256, 256))
    assert lib.b2_ingest_stream_destroy(st) == 0
    for p in (p_in, p_dig, p_new, p_cnt, p_th):
        assert lib.b2_host_free(p) == 0


def test_label_tally_host_entry_point():
    from ics_b200 import labels
    img, cls, act = synth_label_rows(3000, 50, 12)
    counts, partials = labels.label_tally_host(img, cls, act, 3000, 50)
    assert np.array_equal(counts, label_tally(img, cls, act, 3000, 50))
    assert int(partials[50 + 1]) == int(act.sum()) and int(partials[50 + 5]) == img.size
    _, p2 = labels.label_tally_host(img, cls, act, 3000, 50, want_counts=False)
    assert np.array_equal(partials, p2)
    with pytest.raises(ics_b200.B2Error) as e:
        labels.label_tally_host(img[::-1].copy(), cls, act, 3000, 50)
    assert e.value.code == -3


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


def test_mixed_shape_listing_streams_per_shape_and_dedupes_across_shapes():
    """BASELINE config 3 in miniature: a listing of three shapes, two of equal byte length (so byte-identical files of
    different shapes exist), duplicates inside and across shape classes, one stored digest.  Digests == hashlib,
    thumbnails == Pillow, dedupe == the sequential loop of webdav_sync.py:311-400 over the listing."""
    import hashlib

    from PIL import Image

    from ics_b200.hostapi import sort_digests
    from ics_b200.pipeline import MixedShapeIngest

    rng = np.random.default_rng(33)
    shapes = [(64, 256), (128, 128), (96, 72)]                        # 64x256 and 128x128: same byte length
    listing = []
    for i in range(17):
        h, w = shapes[i % 3]
        listing.append(((h, w), rng.integers(0, 256, h * w * 3, dtype=np.uint8)))
    listing[7] = ((128, 128), listing[0][1].copy())                   # same bytes as listing[0], other shape
    listing[9] = (listing[3][0], listing[3][1].copy())                # same bytes, same shape
    listing[16] = ((64, 256), listing[1][1].copy())                   # same bytes as listing[1], other shape
    stored = hashlib.sha256(listing[5][1].tobytes()).digest()
    groups = {}
    for s in shapes:
        pos = [i for i, (sh, _) in enumerate(listing) if sh == s]
        groups[s] = (_pinned(np.stack([listing[i][1] for i in pos])), pos)
    mixed = MixedShapeIngest({s: 8 for s in shapes}, chunk_bytes=3 * 64 * 256 * 3, out_h=32, out_w=48)
    table = sort_digests(np.frombuffer(stored, dtype=np.uint8).reshape(1, 32))
    for _ in range(2):                                                # reusable
        res = mixed.run(groups, table)
    mixed.close()
    seen, created = {stored}, 0
    for i, ((h, w), buf) in enumerate(listing):
        d = hashlib.sha256(buf.tobytes()).digest()
        assert bytes(res.digests[i]) == d
        assert bool(res.is_new[i]) == (d not in seen)
        created += d not in seen
        seen.add(d)
        want = np.asarray(Image.fromarray(buf.reshape(h, w, 3), "RGB").resize((48, 32), Image.BILINEAR))
        assert np.array_equal(res.thumb(i), want)
        np.testing.assert_allclose(res.preview(i), preview_f32(want, (0, 0, 0), (1, 1, 1)), rtol=1e-5, atol=1e-7)
    assert res.stats == {"processed": 17, "created": created, "updated": 17 - created} and created == 13
    assert res.first_index[7] == 0 and res.last_index[0] == 7 and res.first_index[16] == 1
