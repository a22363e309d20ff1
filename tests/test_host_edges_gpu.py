"""Edge cases of the host-pointer entry points (csrc/host.cu): empty batches, empty messages, tiny and odd
image shapes, all-skipped batches, zero rows, and the error returns of the streaming object."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import ics_b200
from ics_b200 import hostapi
from oracle import label_tally, preview_f32, thumbnail_u8

pytestmark = pytest.mark.gpu


def test_hash_and_dedupe_degenerate_batches():
    assert hostapi.hash_batch([]) == []
    assert hostapi.hash_batch([b""]) == [hashlib.sha256(b"").hexdigest()]
    assert hostapi.hash_batch([b"", b"", b"x"]) == [hashlib.sha256(b).hexdigest() for b in (b"", b"", b"x")]
    is_new, first, last, counts = hostapi.dedupe_host(np.zeros((0, 32), np.uint8))
    assert is_new.size == 0 and counts == (0, 0, 0)
    d, _ = hostapi.sha256_host([b"a", b"a", b"b"])
    is_new, first, last, counts = hostapi.dedupe_host(d, valid=np.zeros(3, np.uint8))
    assert is_new.tolist() == [0, 0, 0] and first.tolist() == [-1, -1, -1] and counts == (0, 0, 0)
    is_new, first, last, counts = hostapi.dedupe_host(d)
    assert is_new.tolist() == [1, 0, 1] and first.tolist() == [0, 0, 2] and last.tolist() == [1, 1, 2] and counts == (3, 2, 1)


@pytest.mark.parametrize("shape", [(1, 1), (3, 5), (2, 300), (257, 1), (17, 17)])
def test_thumbnails_tiny_and_odd_shapes(shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    imgs = [rng.integers(0, 256, size=(*shape, 3), dtype=np.uint8) for _ in range(3)]
    for out in ((256, 256), (7, 13)):
        t, p = hostapi.thumbnails(imgs, out[0], out[1], mean=(0.5, 0.4, 0.3), inv_std=(2.0, 3.0, 4.0))
        for i, im in enumerate(imgs):
            want = thumbnail_u8(im, out[0], out[1])
            assert np.array_equal(t[i], want), (shape, out)
            np.testing.assert_allclose(p[i], preview_f32(want, (0.5, 0.4, 0.3), (2.0, 3.0, 4.0)), rtol=1e-5, atol=1e-7)
    assert hostapi.thumbnails([], 8, 8)[0].shape == (0, 8, 8, 3)
    with pytest.raises(ics_b200.B2Error):
        hostapi.thumbnails([np.zeros((4, 4), np.uint8)])


def test_label_tally_host_zero_rows_and_unsorted_mode():
    e = np.zeros(0, np.int32)
    counts, partials = hostapi.label_tally_host(e, e.astype(np.uint8), e.astype(np.uint8), 5, 3)
    assert not counts.any() and not partials.any()
    rng = np.random.default_rng(3)
    img = rng.integers(0, 40, size=5000).astype(np.int32)
    cls = rng.integers(0, 6, size=5000).astype(np.uint8)
    act = (rng.random(5000) < 0.8).astype(np.uint8)
    counts, partials = hostapi.label_tally_host(img, cls, act, 40, 6, sorted_by_image=False)
    assert np.array_equal(counts, label_tally(img, cls, act, 40, 6)) and int(partials[6 + 1]) == int(act.sum())


def test_ingest_ring_error_returns():
    lib = C.CDLL(ics_b200.LIB_PATH)
    lib.b2_last_error.restype = C.c_char_p
    u64 = C.c_uint64
    ring, ticket = C.c_void_p(), u64()
    assert lib.b2_ingest_ring_create(0, u64(1 << 20), u64(0), 2, 8, 8, 0, C.byref(ring)) == -1       # ring below 64 MiB
    assert b"64 MiB" in lib.b2_last_error()
    assert lib.b2_ingest_ring_create(0, u64(64 << 20), u64(0), 0, 8, 8, 0, C.byref(ring)) == -1      # no listing slot
    assert lib.b2_ingest_ring_create(0, u64(64 << 20), u64(0), 1, 8, 8, 0, C.byref(ring)) == 0
    assert lib.b2_ingest_ring_wait(ring, u64(1), None, None, None) == -1                             # nothing submitted
    n = 4
    buf = (C.c_uint8 * (n * 75))()                                                                   # 5x5x3 = 75 bytes: not 16-byte multiples
    ptrs = (C.c_void_p * n)(*[C.addressof(buf) + 75 * i for i in range(n)])
    hw = (C.c_uint32 * (2 * n))(*([5, 5] * n))
    dig, new, cnt, th = (C.c_uint8 * (32 * n))(), (C.c_uint8 * n)(), (C.c_uint32 * 4)(), (C.c_uint8 * (n * 192))()
    prev = (C.c_float * (n * 192))()
    assert lib.b2_ingest_ring_submit(ring, ptrs, hw, None, None, None, 0, None, u64(0), dig, new, None, None, cnt, th, None,
                                     C.byref(ticket)) == -1                                          # empty listing
    assert lib.b2_ingest_ring_submit(ring, ptrs, hw, None, None, None, n, None, u64(0), dig, new, None, None, cnt, th, prev,
                                     C.byref(ticket)) == -1                                          # no previews in this ring
    assert lib.b2_ingest_ring_submit(ring, ptrs, hw, None, None, None, n, None, u64(0), dig, new, None, None, cnt, None, None,
                                     C.byref(ticket)) == -1                                          # pixels but no thumbnail buffer
    assert lib.b2_ingest_ring_submit(ring, ptrs, hw, None, None, None, n, None, u64(0), dig, new, None, None, cnt, th, None,
                                     C.byref(ticket)) == 0, lib.b2_last_error()                      # pageable memory, odd sizes: works
    assert lib.b2_ingest_ring_submit(ring, ptrs, hw, None, None, None, n, None, u64(0), dig, new, None, None, cnt, th, None,
                                     C.byref(u64())) == -1                                           # the only slot is taken
    assert lib.b2_ingest_ring_wait(ring, ticket, None, None, None) == 0
    assert list(cnt)[:3] == [4, 1, 3] and list(new)[:4] == [1, 0, 0, 0]                              # four identical (zero) images
    assert bytes(dig[:32]).hex() == hashlib.sha256(bytes(75)).hexdigest()
    assert lib.b2_ingest_ring_destroy(ring) == 0
    assert lib.b2_ingest_ring_destroy(None) == 0
