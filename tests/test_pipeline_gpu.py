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
        assert [bytes(d).hex() for d in res.digests] == hashes
        assert [bool(x) for x in res.is_new] == is_new
        assert res.stats == stats
        for i in (0, n // 2, n - 1):
            want = thumbnail_u8(images[i], 256, 256)
            assert np.array_equal(res.thumbs[i], want)
            np.testing.assert_allclose(res.previews[i], preview_f32(want), rtol=1e-5, atol=1e-7)
        n_chunks = -(-n // chunk)
        assert n * shape[0] * shape[1] * 3 <= res.h2d_bytes <= n * (shape[0] * shape[1] * 3 + 32)   # images + chunk metadata
        assert n_chunks + 1 + 2 <= res.kernel_launches <= 2 * n_chunks + 2   # resize per chunk, hash per group of chunks, dedupe insert + resolve
    pipe.close()


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
    assert [bool(x) for x in res.is_new] == is_new
    assert res.first_index.tolist() == first and res.stats == stats
    last = [max(j for j in range(n) if hashes[j] == h) for h in hashes]
    assert res.last_index.tolist() == last


def test_native_ring_from_plain_ctypes_and_numpy():
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
    ptrs = (C.c_void_p * n)(*[p_in.value + i * L for i in range(n)])
    hw = (C.c_uint32 * (2 * n))(*([shape[0], shape[1]] * n))
    ring, ticket = C.c_void_p(), C.c_uint64()
    u64 = C.c_uint64
    assert lib.b2_ingest_ring_create(0, u64(64 << 20), u64(4 * L), 2, 256, 256, 0, C.byref(ring)) == 0, lib.b2_last_error()
    args = (ring, ptrs, hw, None, None, None, n, None, u64(0), p_dig, p_new, None, None, p_cnt, p_th, None)
    assert lib.b2_ingest_ring_submit(*args, C.byref(ticket)) == 0, lib.b2_last_error()
    done = C.c_int(-1)
    assert lib.b2_ingest_ring_poll(ring, ticket, C.byref(done), None) == 0 and done.value in (0, 1)
    h2d, d2h, launches = u64(), u64(), C.c_uint32()
    assert lib.b2_ingest_ring_wait(ring, u64(999), None, None, None) == -1            # unknown ticket
    assert lib.b2_ingest_ring_wait(ring, ticket, C.byref(h2d), C.byref(d2h), C.byref(launches)) == 0
    assert lib.b2_ingest_ring_wait(ring, ticket, None, None, None) == -1              # already waited for
    assert h2d.value >= n * L and 3 + 1 + 2 <= launches.value <= 2 * 3 + 2
    assert [bytes(d).hex() for d in h_dig.reshape(n, 32)] == [sha256_hex(im.tobytes()) for im in images]
    assert h_new.tolist() == [1] * n and h_cnt[:3].tolist() == [n, n, 0]
    assert np.array_equal(h_th.reshape(n, 256, 256, 3)[5], thumbnail_u8(images[5], 256, 256))
    assert lib.b2_ingest_ring_submit(ring, ptrs, hw, None, None, None, n, None, u64(0), p_dig, p_new, None, None, p_cnt, p_th,
                                     p_th, C.byref(ticket)) == -1                      # previews from a ring created without them
    assert lib.b2_ingest_ring_destroy(ring) == 0
    assert lib.b2_ingest_ring_destroy(None) == 0
    for p in (p_in, p_dig, p_new, p_cnt, p_th):
        assert lib.b2_host_free(p) == 0
    assert lib.b2_shutdown() == 0


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


def test_listings_in_flight_waited_out_of_order():
    from ics_b200.pipeline import IngestRing
    shape, n = (128, 160), 24
    a = [synth_image(g, *shape) for g in range(n)]
    b = [synth_image(1000 + g, *shape) for g in range(n)]
    c = [synth_image(2000 + g, 96, 64) for g in range(5)]
    ring = IngestRing(ring_bytes=64 << 20, chunk_bytes=5 * 128 * 160 * 3, max_listings=3, max_images=n)
    ta, tb, tc = ring.submit(a), ring.submit(b), ring.submit(c)      # three listings enqueued before any is read
    with pytest.raises(ics_b200.B2Error):
        ring.submit(c)                                               # a fourth: no slot
    rc, rb, ra = ring.result(tc), ring.result(tb), ring.result(ta)
    assert [bytes(d) for d in ra.digests] == [sha256_digest(im.tobytes()) for im in a]
    assert [bytes(d) for d in rb.digests] == [sha256_digest(im.tobytes()) for im in b]
    assert [bytes(d) for d in rc.digests] == [sha256_digest(im.tobytes()) for im in c]
    assert np.array_equal(rb.thumbs[3], thumbnail_u8(b[3], 256, 256))
    assert np.array_equal(rc.thumbs[4], thumbnail_u8(c[4], 256, 256))
    assert ring.stats()["chunks_in_flight"] >= 0
    ring.close()


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


def test_mixed_shape_listing_through_the_ring():
    """BASELINE config 3 in miniature: a listing of three shapes, two of equal byte length (so byte-identical files of
    different shapes exist), duplicates inside and across shapes, one stored digest, chunks that mix shapes.  Digests ==
    hashlib, thumbnails == Pillow, dedupe == the sequential loop of webdav_sync.py:311-400 over the listing."""
    import hashlib

    from PIL import Image

    from ics_b200.hostapi import sort_digests
    from ics_b200.pipeline import IngestRing

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
    ring = IngestRing(ring_bytes=64 << 20, chunk_bytes=3 * 64 * 256 * 3, max_listings=2, max_images=32, out_h=32, out_w=48)
    table = sort_digests(np.frombuffer(stored, dtype=np.uint8).reshape(1, 32))
    pixels = [buf.reshape(h, w, 3) for (h, w), buf in listing]
    for _ in range(2):                                                # reusable
        res = ring.result(ring.submit(pixels, existing_sorted=table))
    seen, created = {stored}, 0
    for i, ((h, w), buf) in enumerate(listing):
        d = hashlib.sha256(buf.tobytes()).digest()
        assert bytes(res.digests[i]) == d
        assert bool(res.is_new[i]) == (d not in seen)
        created += d not in seen
        seen.add(d)
        want = np.asarray(Image.fromarray(buf.reshape(h, w, 3), "RGB").resize((48, 32), Image.BILINEAR))
        assert np.array_equal(res.thumbs[i], want)
        np.testing.assert_allclose(res.previews[i], preview_f32(want, (0, 0, 0), (1, 1, 1)), rtol=1e-5, atol=1e-7)
    assert res.stats == {"processed": 17, "created": created, "updated": 17 - created} and created == 13
    assert res.first_index[7] == 0 and res.last_index[0] == 7 and res.first_index[16] == 1
    ring.close()


def test_ring_file_bytes_pixels_and_skipped_entries():
    """The real service's listing (feeder output): the message hashed is the downloaded FILE, the thumbnail comes from
    the decoded pixels; a failed download is skipped (webdav_sync.py:320), a file Pillow cannot decode has no pixels."""
    import hashlib

    from ics_b200.pipeline import IngestRing

    rng = np.random.default_rng(5)
    n = 11
    files = [rng.integers(0, 256, int(rng.integers(1, 70000)), dtype=np.uint8) for _ in range(n)]
    files[4] = files[2].copy()                                        # duplicate content
    pixels = [synth_image(300 + i, 40 + 3 * i, 57 + i) for i in range(n)]
    pixels[6] = None                                                  # undecodable: hashed and counted, no thumbnail
    valid = np.ones(n, dtype=np.uint8)
    valid[3] = 0                                                      # failed download
    files[3] = None
    pixels[3] = None
    ring = IngestRing(ring_bytes=64 << 20, chunk_bytes=150_000, max_listings=2, max_images=16, out_h=24, out_w=24, want_preview=False)
    res = ring.result(ring.submit(pixels, files=files, valid=valid))
    seen = set()
    for i in range(n):
        if not valid[i]:
            assert not res.is_new[i] and res.first_index[i] == -1
            continue
        d = hashlib.sha256(files[i].tobytes()).digest()
        assert bytes(res.digests[i]) == d, i
        assert bool(res.is_new[i]) == (d not in seen)
        seen.add(d)
        if pixels[i] is not None:
            assert np.array_equal(res.thumbs[i], thumbnail_u8(pixels[i], 24, 24)), i
    assert res.stats == {"processed": 10, "created": 9, "updated": 1}
    # hash + dedupe only (no pixels at all)
    res = ring.result(ring.submit(None, files=[f for f in files if f is not None]))
    assert res.thumbs is None and res.stats == {"processed": 10, "created": 9, "updated": 1}
    ring.close()


def test_ring_back_pressure_and_errors():
    """A listing several times larger than the device ring streams through it (submit blocks while the ring is full);
    an image that cannot fit is refused, everything in flight is drained, and the ring stays usable."""
    from ics_b200.pipeline import IngestRing
    shape, n = (512, 512), 120                                        # 94 MB of images through a 64 MiB ring
    base = [synth_image(7000 + g, *shape) for g in range(6)]
    images = [base[g % 6] for g in range(n)]
    host = _pinned(images)
    ring = IngestRing(ring_bytes=64 << 20, chunk_bytes=8 << 20, max_listings=2, max_images=n, want_preview=False)
    t = ring.submit_packed(host, shape)
    done, flushed = ring.poll(t)
    assert 0 <= flushed <= n
    res = ring.result(t)
    want = [sha256_digest(im.tobytes()) for im in base]
    assert [bytes(d) for d in res.digests] == [want[g % 6] for g in range(n)]
    assert res.stats == {"processed": n, "created": 6, "updated": n - 6}
    assert np.array_equal(res.thumbs[n - 1], thumbnail_u8(images[n - 1], 256, 256))
    assert ring.stats()["stalls"] > 0
    big = np.zeros((5000, 5000, 3), dtype=np.uint8)                   # 75 MB: larger than the ring
    with pytest.raises(ics_b200.B2Error):
        ring.submit([images[0], big, images[1]])
    res = ring.result(ring.submit(images[:7]))
    assert [bytes(d) for d in res.digests] == [want[g % 6] for g in range(7)]
    ring.close()


def test_ring_stress_random_listings_wraparound():
    """Many listings of random sizes (empty files, one-byte files, odd image shapes, more entries than one chunk holds)
    through a small ring that wraps many times, three in flight, waited for in rotation: every digest == hashlib, every
    thumbnail == Pillow, every dedupe decision == the sequential loop over that listing."""
    import hashlib

    from ics_b200.pipeline import IngestRing
    rng = np.random.default_rng(2024)
    ring = IngestRing(ring_bytes=64 << 20, chunk_bytes=3 << 20, max_listings=3, max_images=64, out_h=32, out_w=40, want_preview=True)
    pool = [rng.integers(0, 256, int(rng.integers(0, 400_000)), dtype=np.uint8) for _ in range(40)]
    pool[0] = np.zeros(0, dtype=np.uint8)                             # an empty file
    pool[1] = np.array([7], dtype=np.uint8)                           # one byte
    pics = [synth_image(9000 + i, int(rng.integers(33, 300)), int(rng.integers(41, 300))) for i in range(12)]
    inflight = []

    def check(entry):
        ticket, files, pixels = entry
        res = ring.result(ticket)
        seen = set()
        for i, f in enumerate(files):
            d = hashlib.sha256(f.tobytes()).digest()
            assert bytes(res.digests[i]) == d
            assert bool(res.is_new[i]) == (d not in seen)
            seen.add(d)
            if pixels[i] is not None:
                assert np.array_equal(res.thumbs[i], thumbnail_u8(pixels[i], 32, 40))
        assert res.stats["processed"] == len(files) and res.stats["created"] == len(seen)

    for it in range(30):
        n = int(rng.integers(1, 60))
        files = [pool[int(j)] for j in rng.integers(0, len(pool), n)]
        pixels = [pics[int(j)] if rng.random() < 0.5 else None for j in rng.integers(0, len(pics), n)]
        if len(inflight) == 3:
            check(inflight.pop(0))
        inflight.append((ring.submit(pixels, files=files), files, pixels))
    while inflight:
        check(inflight.pop(0))
    assert ring.stats()["bytes_in_flight"] == 0 or ring.stats()["chunks_in_flight"] >= 0
    ring.close()
